// lpc_util.cuh -- small device/host helpers shared by the lpc kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

namespace lpc {

template <typename T> struct Cx;
template <> struct Cx<float> { using type = float2; };
template <> struct Cx<double> { using type = double2; };

// 16-byte vector of T (float4 / double2): the unit of every HBM access in the hot kernels.
template <typename T> struct Vec16;
template <> struct alignas(16) Vec16<float> { float v[4]; };
template <> struct alignas(16) Vec16<double> { double v[2]; };

template <typename T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
    return *reinterpret_cast<const Vec16<T>*>(p);
}
// read-only, non-coherent path (LDG.NC) for arrays the kernel never writes
__device__ __forceinline__ Vec16<float> ld16_nc(const float* p) {
    float4 r = __ldg(reinterpret_cast<const float4*>(p));
    Vec16<float> o; o.v[0] = r.x; o.v[1] = r.y; o.v[2] = r.z; o.v[3] = r.w; return o;
}
__device__ __forceinline__ Vec16<double> ld16_nc(const double* p) {
    double2 r = __ldg(reinterpret_cast<const double2*>(p));
    Vec16<double> o; o.v[0] = r.x; o.v[1] = r.y; return o;
}
template <typename T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& x) {
    *reinterpret_cast<Vec16<T>*>(p) = x;
}

// 32-byte (256-bit) global accesses, sm_100+ (SASS LDG.E.256 / STG.E.256): one request per 32-byte
// chunk of the T-/G-layout instead of two 16-byte ones -- the LSU data pipe, not HBM, bounds the
// fused row kernels (profiles/r01_ncu_full_v2_three_kernels_raw.csv: 77-90 % wavefront utilisation).
struct alignas(32) Vec32f { float v[8]; };
__device__ __forceinline__ Vec32f ld32_nc(const float* p) {
    Vec32f o;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3]), "=f"(o.v[4]), "=f"(o.v[5]), "=f"(o.v[6]), "=f"(o.v[7])
        : "l"(p));
    return o;
}
__device__ __forceinline__ void st32(float* p, const Vec32f& x) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]),
                 "f"(x.v[3]), "f"(x.v[4]), "f"(x.v[5]), "f"(x.v[6]), "f"(x.v[7])
                 : "memory");
}

// ---- L2 eviction-priority hints (sm_80+ `.L2::cache_hint` with a 64-bit policy; the constants are the fixed
// fractional-1.0 policies CUTLASS ships as TMA::CacheHintSm90).  The 126 MB L2 holds about half of one 199 MB
// kernel-to-kernel hand-off of the fused ADMM engine, and consecutive launches walk in opposite directions
// (FusedK::rev) so a consumer starts on what its producer wrote last.  MEASURED at 3x1080x1920 (recon/s, one run
// each): no hints 20.65; bit 0 (hand-off writes evict-last) 20.76; + bit 1 (last-use reads of hand-offs
// evict-first) 20.55; + bit 2 (duals / measurement evict-first) 20.51; + bit 3 (evict-first on the last sweep over
// a halo row of V / eta') 19.84 -- CTAs run out of order, so "last" readers evict rows their neighbours still
// need.  Default: bit 0 only.
#ifndef LPC_L2_HINTS
#define LPC_L2_HINTS 1
#endif
constexpr unsigned long long LPC_L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr unsigned long long LPC_L2_EF_ = 0x12F0000000000000ull, LPC_L2_EL_ = 0x14F0000000000000ull;
constexpr unsigned long long LPC_L2_EVICT_LAST = (LPC_L2_HINTS & 1) ? LPC_L2_EL_ : LPC_L2_EVICT_NORMAL;   // hand-off writes
constexpr unsigned long long LPC_L2_EVICT_FIRST = (LPC_L2_HINTS & 2) ? LPC_L2_EF_ : LPC_L2_EVICT_NORMAL;  // last-use reads of hand-offs
constexpr unsigned long long LPC_L2_STREAM = (LPC_L2_HINTS & 4) ? LPC_L2_EF_ : LPC_L2_EVICT_NORMAL;       // duals, measurement
constexpr unsigned long long LPC_L2_HALO = (LPC_L2_HINTS & 8) ? LPC_L2_EF_ : LPC_L2_EVICT_NORMAL;         // last sweep over a halo row
__device__ __forceinline__ Vec16<float> ld16_nc_hint(const float* p, unsigned long long pol) {
    Vec16<float> o;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3])
                 : "l"(p), "l"(pol));
    return o;
}
__device__ __forceinline__ Vec16<float> ld16_hint(const float* p, unsigned long long pol) {
    Vec16<float> o;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3])
                 : "l"(p), "l"(pol)
                 : "memory");
    return o;
}
// ---- the same accesses WITHOUT allocating the line in L1 (.L1::no_allocate): streams that are touched once must not
// evict what the kernel does re-read through L1 (twiddle tables, the +-1 rows of V).  Measured on B1: its V / HV stores
// allocated lines and pushed the L1 hit rate of the twiddle loads down to 72 % -- no-allocate stores: 88.2 -> 83.3 us.
__device__ __forceinline__ Vec16<float> ld16_nc_hint_na(const float* p, unsigned long long pol) {
    Vec16<float> o;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3])
                 : "l"(p), "l"(pol));
    return o;
}
__device__ __forceinline__ Vec16<float> ld16_hint_na(const float* p, unsigned long long pol) {
    Vec16<float> o;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3])
                 : "l"(p), "l"(pol)
                 : "memory");
    return o;
}
__device__ __forceinline__ void st16_hint_na(float* p, const Vec16<float>& x, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]),
                 "f"(x.v[2]), "f"(x.v[3]), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st16_na(float* p, const Vec16<float>& x) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3]) : "memory");
}
__device__ __forceinline__ void st4_na(float* p, float x) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
}
__device__ __forceinline__ float ld4_na(const float* p) {
    float o;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(o) : "l"(p) : "memory");
    return o;
}
__device__ __forceinline__ float ld4_nc_na(const float* p) {
    float o;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(o) : "l"(p));
    return o;
}
__device__ __forceinline__ Vec16<float> ld16_nc_na(const float* p) {
    Vec16<float> o;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3]) : "l"(p));
    return o;
}
__device__ __forceinline__ void st16_hint(float* p, const Vec16<float>& x, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]),
                 "f"(x.v[3]), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st4_hint_na(float* p, float x, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(x), "l"(pol) : "memory");
}
__device__ __forceinline__ void st4_hint(float* p, float x, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(x), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 ld8_nc_hint(const float2* p, unsigned long long pol) {
    float2 o;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(o.x), "=f"(o.y) : "l"(p), "l"(pol));
    return o;
}

template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }

// complex helpers on float2/double2
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }

// CAS-loop atomic max / min for float and double (setup-time reductions only)
__device__ __forceinline__ void atomic_max(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) < v) {
        int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_min(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) > v) {
        int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_min(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = tmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = tmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace lpc
