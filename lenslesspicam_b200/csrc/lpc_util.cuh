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
