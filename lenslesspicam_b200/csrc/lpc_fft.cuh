// lpc_fft.cuh -- hand-written mixed-radix FFT core (5-smooth lengths), the building block of
// the fused engine-v2 kernels (lpc_fused.cuh).
//
//  * In-register DFTs of size 2,3,4,5 (hand-coded) and 6,8,9,10,12,15,16 (straight-line
//    Cooley-Tukey compositions with literal twiddles, generated: lpc_fft_radix.inc).
//  * One Stockham autosort pass = load R points (stride N/R) -> twiddle -> DFT_R -> store
//    (stride Ns): natural order in, natural order out, no bit reversal.  Passes run in
//    shared memory, in place, one butterfly per thread-iteration, the read and the write
//    phase of a pass separated by a CTA barrier.
//  * Per-pass twiddle tables laid out [m-1][k] so that the lanes of a warp (consecutive k)
//    read consecutive entries (coalesced, L1-resident).
//
// Everything that does index arithmetic or floating point is __host__ __device__ so the
// host harness tests/csrc/test_fft_host.cpp can run exactly this code on the CPU (there is
// no GPU in the build container).
//
// The reference has no FFT of its own: it calls scipy.fft / torch.fft
// (rfft_convolve.py:121-131,146-170,191-216; admm.py:286-293).  Sizes are always 5-smooth
// (next_fast_len, rfft_convolve.py:112).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LPC_HD
#define LPC_HD __host__ __device__ __forceinline__
#endif

namespace lpc {

template <typename T> struct CxT;
template <> struct CxT<float> { using type = float2; };
template <> struct CxT<double> { using type = double2; };

// Blackwell packed fp32x2 arithmetic (FADD2 / FFMA2, sm_100+): one instruction per complex add / subtract
// instead of two, bit-identical to the scalar form (a - b == fma(b, -1, a) exactly).  MEASURED on B200
// (profiles/README.md, round 1 session 2): 23 % fewer floating-point instructions, but no faster -- column
// kernel 174.6 vs 172.4 us, B1 114 vs 111 us, FISTA-300 96.9 vs 93.3 ms -- the FFT kernels are bound by
// latency and the LSU pipe, not by instruction issue.  Off by default; -DLPC_PACKED2=1 builds it.
#ifndef LPC_PACKED2
#define LPC_PACKED2 0
#endif
#ifndef LPC_TW_PRODUCTS
#define LPC_TW_PRODUCTS 1
#endif
template <typename C> LPC_HD C cx_add(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> LPC_HD C cx_sub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && LPC_PACKED2
template <> __device__ __forceinline__ float2 cx_add<float2>(float2 a, float2 b) { return __fadd2_rn(a, b); }
template <> __device__ __forceinline__ float2 cx_sub<float2>(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
#endif
template <typename C> LPC_HD C cx_mul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR, typename C> LPC_HD C cx_rot(C a) {
    C r;
    if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
    return r;
}

// ----------------------------------------------------------------------------------------
// in-register DFT of size R:  v[k] <- sum_n v[n] * exp(DIR * 2*pi*i*n*k/R), natural order
// ----------------------------------------------------------------------------------------
template <typename T, int R, int DIR> struct Dft;

template <typename T, int DIR> struct Dft<T, 2, DIR> {
    using C = typename CxT<T>::type;
    static LPC_HD void run(C* v) {
        C a = v[0], b = v[1];
        v[0] = cx_add(a, b);
        v[1] = cx_sub(a, b);
    }
};

template <typename T, int DIR> struct Dft<T, 3, DIR> {
    using C = typename CxT<T>::type;
    static LPC_HD void run(C* v) {
        const T s3 = (T)0.86602540378443864676;
        C t = cx_add(v[1], v[2]);
        C d = cx_sub(v[1], v[2]);
        C m; m.x = v[0].x - (T)0.5 * t.x; m.y = v[0].y - (T)0.5 * t.y;
        C r = cx_rot<DIR>(d); r.x *= s3; r.y *= s3;
        v[0] = cx_add(v[0], t);
        v[1] = cx_add(m, r);
        v[2] = cx_sub(m, r);
    }
};

template <typename T, int DIR> struct Dft<T, 4, DIR> {
    using C = typename CxT<T>::type;
    static LPC_HD void run(C* v) {
        C t0 = cx_add(v[0], v[2]), t1 = cx_sub(v[0], v[2]);
        C t2 = cx_add(v[1], v[3]), t3 = cx_rot<DIR>(cx_sub(v[1], v[3]));
        v[0] = cx_add(t0, t2);
        v[2] = cx_sub(t0, t2);
        v[1] = cx_add(t1, t3);
        v[3] = cx_sub(t1, t3);
    }
};

template <typename T, int DIR> struct Dft<T, 5, DIR> {
    using C = typename CxT<T>::type;
    static LPC_HD void run(C* v) {
        const T c1 = (T)0.30901699437494742410;   // cos(2pi/5)
        const T c2 = (T)-0.80901699437494742410;  // cos(4pi/5)
        const T s1 = (T)0.95105651629515357212;   // sin(2pi/5)
        const T s2 = (T)0.58778525229247312917;   // sin(4pi/5)
        C a1 = cx_add(v[1], v[4]), a2 = cx_add(v[2], v[3]);
        C b1 = cx_sub(v[1], v[4]), b2 = cx_sub(v[2], v[3]);
        C p1, p2, q1, q2;
        p1.x = v[0].x + c1 * a1.x + c2 * a2.x; p1.y = v[0].y + c1 * a1.y + c2 * a2.y;
        p2.x = v[0].x + c2 * a1.x + c1 * a2.x; p2.y = v[0].y + c2 * a1.y + c1 * a2.y;
        q1.x = s1 * b1.x + s2 * b2.x; q1.y = s1 * b1.y + s2 * b2.y;
        q2.x = s2 * b1.x - s1 * b2.x; q2.y = s2 * b1.y - s1 * b2.y;
        q1 = cx_rot<DIR>(q1);
        q2 = cx_rot<DIR>(q2);
        v[0].x = v[0].x + a1.x + a2.x; v[0].y = v[0].y + a1.y + a2.y;
        v[1] = cx_add(p1, q1);
        v[4] = cx_sub(p1, q1);
        v[2] = cx_add(p2, q2);
        v[3] = cx_sub(p2, q2);
    }
};

// composite radices 6, 8, 9, 10, 12, 15, 16: straight-line code generated by gen/gen_fft_radix.py
#include "lpc_fft_radix.inc"

// ----------------------------------------------------------------------------------------
// plan: radix sequence + per-pass twiddle tables
// ----------------------------------------------------------------------------------------
constexpr int LPC_FFT_MAX_PASSES = 12;
struct FftPlan {
    int N;
    int npass;
    int radix[LPC_FFT_MAX_PASSES];
    int Ns[LPC_FFT_MAX_PASSES];      // product of the radices of the previous passes
    int tw_off[LPC_FFT_MAX_PASSES];  // offset (complex entries) of the pass's table [m-1][k], k < Ns
    int tw_len;                      // total entries
    int padsh;                       // shared-memory padding: phys(i) = i + (i >> padsh); 4 or 30 (= none)
};

// Shared-memory padding.  The first Stockham pass stores with element stride R: conflict-free
// across a half-warp of 8-byte accesses iff the stride is odd.  The planner puts an odd radix
// first when N has one (padsh = 30: no padding); for power-of-two-only first radices it pads
// one complex slot every 16 (padsh = 4) which turns stride 16/8/4/2 into 17 / 8.5 / 4.25 / 2.125.
LPC_HD int fft_pad(int i, int padsh) { return padsh >= 30 ? i : i + (i >> padsh); }  // folds away when padsh is a constant 30
LPC_HD int fft_padded_len(int n, int padsh) { return n + (n >> padsh) + 1; }

// ----------------------------------------------------------------------------------------
// one butterfly of one Stockham pass, split in the three phases the CTA barrier separates.
// Element i of the transform lives at sm[fft_pad(base + i*es)].
// ----------------------------------------------------------------------------------------
template <typename T, int R>
LPC_HD void fft_bfly_load(const typename CxT<T>::type* sm, int base, int es, int padsh, int N, int j,
                          typename CxT<T>::type* v) {
    const int stride = N / R;
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = sm[fft_pad(base + (j + m * stride) * es, padsh)];
}

// twiddle (skipped when Ns == 1) + in-register DFT.  tw = this pass's table [m-1][k]
template <typename T, int R, int DIR>
LPC_HD void fft_bfly_compute(typename CxT<T>::type* v, const typename CxT<T>::type* tw, int Ns, int j) {
    using C = typename CxT<T>::type;
    if (Ns > 1) {
        const int k = j % Ns;
#if LPC_TW_PRODUCTS
        if constexpr (R >= 6) {
            // Only the twiddles w^1, w^2, w^4, w^8 of the butterfly come from the table; the others are products of those
            // (at most three deep: 7 = (1*2)*4, 15 = 7*8).  The row kernels are bound by the LSU data pipe, where the 15
            // table loads per butterfly and pass were a fifth of the wavefronts, while their FMA pipe is a third busy
            // (timing experiment without any twiddle load: B1 83 -> 74 us, GD rows 62 / 65 -> 57 / 59 us).  Accuracy: the host
            // harness (tests/csrc/test_fft_host.cpp) measures the transforms against a long-double DFT with this scheme.
            C w[R];
#pragma unroll
            for (int m = 1; m < R; m *= 2) {
#if defined(__CUDA_ARCH__)
                w[m] = __ldg(tw + (m - 1) * Ns + k);
#else
                w[m] = tw[(m - 1) * Ns + k];
#endif
                if (DIR > 0) w[m].y = -w[m].y;
            }
#pragma unroll
            for (int m = 3; m < R && m < 8; ++m) {
                const int hb = m >= 4 ? 4 : 2;  // highest power of two in m
                if (m != hb) w[m] = cx_mul(w[m - hb], w[hb]);
            }
#pragma unroll
            for (int m = 1; m < R && m < 8; ++m) v[m] = cx_mul(v[m], w[m]);
            if constexpr (R > 8) {
                v[8] = cx_mul(v[8], w[8]);
#pragma unroll
                for (int m = 9; m < R; ++m) v[m] = cx_mul(v[m], cx_mul(w[m - 8], w[8]));  // (w^9 .. w^15 are never all live)
            }
        } else
#endif
        {
#pragma unroll
            for (int m = 1; m < R; ++m) {
#if defined(__CUDA_ARCH__) && defined(LPC_EXP_NO_TW)
                C w; w.x = (T)1 - (T)1e-9 * (T)k; w.y = (T)1e-9 * (T)m;   // timing experiment only: no table loads
#elif defined(__CUDA_ARCH__)
                C w = __ldg(tw + (m - 1) * Ns + k);
#else
                C w = tw[(m - 1) * Ns + k];
#endif
                if (DIR > 0) w.y = -w.y;
                v[m] = cx_mul(v[m], w);
            }
        }
    }
    Dft<T, R, DIR>::run(v);
}

template <typename T, int R>
LPC_HD void fft_bfly_store(typename CxT<T>::type* sm, int base, int es, int padsh, int Ns, int j,
                           const typename CxT<T>::type* v) {
    const int k = j % Ns;
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int m = 0; m < R; ++m) sm[fft_pad(base + (j0 + m * Ns) * es, padsh)] = v[m];
}

// ----------------------------------------------------------------------------------------
// In-place decimation passes (no autosort), used where the data meets an element-wise stage
// between a forward and an inverse transform (the column kernel of lpc_fused_tma.cuh):
//   forward = decimation in frequency: natural order in, mixed-radix digit-reversed order out;
//   inverse = its mirror image, decimation in time: digit-reversed in, natural out.
// A butterfly reads and writes the SAME R addresses, so a pass needs ONE barrier (not the two of
// a Stockham pass), nothing is staged in registers across the barrier, and any number of threads
// can walk the butterflies of a pass.  The element-wise stage works in the permuted order (its
// tables are permuted once per PSF).
//
// Pass over blocks of size M (M = N / product of the radices of the earlier DIF passes), radix R,
// stride S = M/R: butterfly j -> block j / S, offset n2 = j % S, elements blk*M + n2 + m*S.
//   DIF: v <- DFT_R(v);  v[k1] *= w_M^(n2*k1)          DIT: v[k1] *= w_M^(n2*k1);  v <- DFT_R(v)
// tw = this pass's table [(k1-1)*S + n2] = exp(-2*pi*i*n2*k1/M) (conjugated for DIR > 0); a pass
// with S == 1 has no twiddles.  After all DIF passes position p holds frequency fft_ip_freq_of_pos(p).
// ----------------------------------------------------------------------------------------
// WIN (outermost pass of a zero-padded transform whose non-zero elements are lo <= index < hi): a DIF butterfly takes
// the elements outside the window as zeros WITHOUT reading them (the buffer need not be zero-filled), a DIT butterfly
// does not store the outputs outside the window (nobody reads them).
template <typename T, int R, int DIR, bool DIF, bool WIN = false>
LPC_HD void fft_ip_bfly(typename CxT<T>::type* sm, int es, int M, int j, const typename CxT<T>::type* tw, int lo = 0, int hi = 0) {
    using C = typename CxT<T>::type;
    const int S = M / R;
    const int blk = j / S, n2 = j - blk * S;
    const int e0 = blk * M + n2;
    C* p = sm + e0 * es;
    C v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        if (WIN && DIF) {
            const int e = e0 + m * S;
            v[m].x = v[m].y = T(0);
            if (e >= lo && e < hi) v[m] = p[m * S * es];
        } else {
            v[m] = p[m * S * es];
        }
    }
    if (DIF) Dft<T, R, DIR>::run(v);
    if (S > 1) {
#pragma unroll
        for (int m = 1; m < R; ++m) {
            C w = tw[(m - 1) * S + n2];  // plain load: the table may live in shared memory
            if (DIR > 0) w.y = -w.y;
            v[m] = cx_mul(v[m], w);
        }
    }
    if (!DIF) Dft<T, R, DIR>::run(v);
#pragma unroll
    for (int m = 0; m < R; ++m) {
        if (WIN && !DIF) {
            const int e = e0 + m * S;
            if (e >= lo && e < hi) p[m * S * es] = v[m];
        } else {
            p[m * S * es] = v[m];
        }
    }
}

// frequency index held at position p after the forward (DIF) passes with radices r[0..npass)
LPC_HD int fft_ip_freq_of_pos(int N, const int* r, int npass, int p) {
    int k = 0, mult = 1, M = N;
    for (int i = 0; i < npass; ++i) {
        const int S = M / r[i];
        const int a = p / S;
        p -= a * S;
        k += a * mult;
        mult *= r[i];
        M = S;
    }
    return k;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------
// CTA-cooperative in-place FFT with a COMPILE-TIME radix sequence.
//
//   FftSeq<N, R0, R1, ...>      N = R0*R1*...; an odd R0 means no shared-memory padding
//   cta_fft_static<T, SEQ, DIR, NB, NTH>(sm, tw, tid)
//
// NB transforms are interleaved in shared memory: element i of transform b lives at
// sm[fft_pad(i*NB + b)].  Everything (butterfly counts, strides, Ns, twiddle offsets, the number
// of butterflies a thread stages in registers) is a constant expression, so each pass is
// straight-line code: no calls, no local memory, j % Ns strength-reduced by the compiler.
// (A first version dispatched on run-time radices: with 11 radices x 2 directions inlined nvcc
// kept the staging arrays in local memory, and as __noinline__ functions the ABI register
// saves cost 0.8 GB of extra DRAM writes per launch -- profiles/r01_ncu_v2a_*.)
// All NTH threads of the CTA must call it (it contains __syncthreads).
// ----------------------------------------------------------------------------------------
template <int N_, int... Rs> struct FftSeq {
    static constexpr int N = N_;
    static constexpr int npass = sizeof...(Rs);
    static constexpr int radix[sizeof...(Rs)] = {Rs...};
    static constexpr int padsh = (radix[0] & 1) ? 30 : 4;
};

template <typename T, int R, int DIR, int NB, int NTH, int PADSH, int N, int NS>
__device__ __forceinline__ void cta_fft_static_pass(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
    using C = typename CxT<T>::type;
    constexpr int nb = N / R;
    constexpr int total = nb * NB;
    constexpr int MAXB = (total + NTH - 1) / NTH;
    constexpr bool exact = (MAXB * NTH == total);
    C v[MAXB][R];
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int g = tid + it * NTH;
        if (exact || g < total) {
            const int b = g % NB, j = g / NB;
            fft_bfly_load<T, R>(sm, b, NB, PADSH, N, j, v[it]);
            fft_bfly_compute<T, R, DIR>(v[it], tw, NS, j);
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int g = tid + it * NTH;
        if (exact || g < total) {
            const int b = g % NB, j = g / NB;
            fft_bfly_store<T, R>(sm, b, NB, PADSH, NS, j, v[it]);
        }
    }
    __syncthreads();
}

template <typename T, int DIR, int NB, int NTH, int PADSH, int N, int NS, int TWOFF, int R, int... Rest>
__device__ __forceinline__ void cta_fft_static_passes(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
    cta_fft_static_pass<T, R, DIR, NB, NTH, PADSH, N, NS>(sm, tw + TWOFF, tid);
    if constexpr (sizeof...(Rest) > 0)
        cta_fft_static_passes<T, DIR, NB, NTH, PADSH, N, NS * R, TWOFF + (NS == 1 ? 0 : NS * (R - 1)), Rest...>(sm, tw, tid);
}

// ---- pieces of a sequence, for kernels that keep the data in registers across the boundary between
//      two transforms: "all passes but the last" of one, "all passes but the first" of the next.
//      When the next transform uses the REVERSED radix sequence, its first pass (radix RL, Ns = 1,
//      input stride N/RL) reads exactly the indices j + m*(N/RL) that the last pass of the previous
//      transform (radix RL, Ns = N/RL) has just produced in the registers of butterfly j.
template <typename T, int DIR, int NB, int NTH, int PADSH, int N, int NS, int TWOFF, int R, int... Rest>
__device__ __forceinline__ void cta_fft_static_butlast(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
    if constexpr (sizeof...(Rest) > 0) {
        cta_fft_static_pass<T, R, DIR, NB, NTH, PADSH, N, NS>(sm, tw + TWOFF, tid);
        cta_fft_static_butlast<T, DIR, NB, NTH, PADSH, N, NS * R, TWOFF + (NS == 1 ? 0 : NS * (R - 1)), Rest...>(sm, tw, tid);
    }
}
template <typename T, int DIR, int NB, int NTH, int PADSH, int N, int R0, int... Rest>
__device__ __forceinline__ void cta_fft_static_butfirst(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
    if constexpr (sizeof...(Rest) > 0) cta_fft_static_passes<T, DIR, NB, NTH, PADSH, N, R0, 0, Rest...>(sm, tw, tid);
}

// ---- the FIRST Stockham pass (Ns = 1, no twiddles) with its input produced by a functor ld(i) -> element i instead of
//      read from the transform buffer: the caller's gather / expansion stage feeds the butterflies directly and one
//      shared-memory round trip of the row (and its barrier) disappears.  The outputs are written to sm right away: the
//      caller guarantees that nobody still reads sm (its previous use ended at a barrier).  Ends with a barrier.
template <typename T, int R, int DIR, int NTH, int PADSH, int N, typename LD>
__device__ __forceinline__ void cta_fft_static_first_pass_ld(typename CxT<T>::type* sm, int tid, LD ld) {
    using C = typename CxT<T>::type;
    constexpr int nb = N / R;
#pragma unroll 1
    for (int j = tid; j < nb; j += NTH) {
        C v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = ld(j + m * nb);
        Dft<T, R, DIR>::run(v);
        fft_bfly_store<T, R>(sm, 0, 1, PADSH, 1, j, v);
    }
    __syncthreads();
}
// passes 2 .. npass-1 of a sequence (everything between the first and the last)
template <typename T, int DIR, int NB, int NTH, int PADSH, int N, int R0, int... Rest>
__device__ __forceinline__ void cta_fft_static_mid(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
    if constexpr (sizeof...(Rest) > 1) cta_fft_static_butlast<T, DIR, NB, NTH, PADSH, N, R0, 0, Rest...>(sm, tw, tid);
}

template <int... Rs> struct IntList {};
template <typename Acc, int... Rs> struct RevList;
template <int... Acc> struct RevList<IntList<Acc...>> { using type = IntList<Acc...>; };
template <int... Acc, int R, int... Rs> struct RevList<IntList<Acc...>, R, Rs...> : RevList<IntList<R, Acc...>, Rs...> {};
template <int N, typename L> struct SeqFromList;
template <int N, int... Rs> struct SeqFromList<N, IntList<Rs...>> { using type = FftSeq<N, Rs...>; };

template <typename SEQ> struct SeqTraits;
template <int N_, int... Rs> struct SeqTraits<FftSeq<N_, Rs...>> {
    using Seq = FftSeq<N_, Rs...>;
    using Rev = typename SeqFromList<N_, typename RevList<IntList<>, Rs...>::type>::type;  // reversed radix order
    static constexpr int N = N_;
    static constexpr int npass = sizeof...(Rs);
    static constexpr int last = Seq::radix[npass - 1];
    static constexpr int ns_last = N_ / last;  // Ns of the last pass
    static constexpr int compute_tw_off_last() {  // offset of the last pass's table in the forward-ordered tables
        int off = 0, ns = 1;
        for (int p = 0; p + 1 < npass; ++p) { off += (ns == 1) ? 0 : ns * (Seq::radix[p] - 1); ns *= Seq::radix[p]; }
        return off;
    }
    static constexpr int tw_off_last = compute_tw_off_last();
    template <typename T, int DIR, int NB, int NTH>
    static __device__ __forceinline__ void run_butlast(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
        cta_fft_static_butlast<T, DIR, NB, NTH, Seq::padsh, N_, 1, 0, Rs...>(sm, tw, tid);
    }
    static constexpr int first = Seq::radix[0];
    template <typename T, int DIR, int NTH, typename LD>
    static __device__ __forceinline__ void run_first_ld(typename CxT<T>::type* sm, int tid, LD ld) {
        cta_fft_static_first_pass_ld<T, Seq::radix[0], DIR, NTH, Seq::padsh, N_, LD>(sm, tid, ld);
    }
    template <typename T, int DIR, int NB, int NTH>
    static __device__ __forceinline__ void run_mid(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
        cta_fft_static_mid<T, DIR, NB, NTH, Seq::padsh, N_, Rs...>(sm, tw, tid);
    }
    template <typename T, int DIR, int NB, int NTH>
    static __device__ __forceinline__ void run_butfirst(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
        cta_fft_static_butfirst<T, DIR, NB, NTH, Seq::padsh, N_, Rs...>(sm, tw, tid);
    }
};

template <typename T, typename SEQ, int DIR, int NB, int NTH> struct CtaFftStatic;
template <typename T, int N, int... Rs, int DIR, int NB, int NTH> struct CtaFftStatic<T, FftSeq<N, Rs...>, DIR, NB, NTH> {
    static __device__ __forceinline__ void run(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid) {
        cta_fft_static_passes<T, DIR, NB, NTH, FftSeq<N, Rs...>::padsh, N, 1, 0, Rs...>(sm, tw, tid);
    }
};
// ---- barrier policies: the whole CTA, or a named barrier shared by a group of COUNT threads
struct BarCta {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
template <int COUNT> struct BarNamed {
    int id;  // 1..15 (0 is __syncthreads)
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }
};

// ---- one in-place pass walked by NTH threads over NB interleaved transforms of length N
// (Measured and dropped, round 2: a butterfly that takes TWO adjacent transforms with 16-byte accesses and shared twiddle
// loads -- half the shared-memory instructions of the outer passes -- made kernel A slower, 140.6 -> 143.8 us without
// spills at 2 x 192 threads and 149-155 us with the spills of 2 x 288: the kernel wants more independent warps, not
// fewer instructions.  profiles/README.md.)
template <typename T, int R, int DIR, bool DIF, int NB, int NTH, int N, int M, bool WIN = false>
__device__ __forceinline__ void cta_fft_ip_pass(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, int lo = 0, int hi = 0) {
    constexpr int total = (N / R) * NB;
#pragma unroll 1
    for (int g = tid; g < total; g += NTH) {
        const int b = g % NB, j = g / NB;
        fft_ip_bfly<T, R, DIR, DIF, WIN>(sm + b, NB, M, j, tw, lo, hi);
    }
}
// forward: all DIF passes but the last (radix R_last, S = 1, no twiddles -- the caller fuses it with its
// element-wise stage); a barrier separates consecutive passes, NONE follows the last one: the caller issues
// whatever it wants in flight across that barrier (prefetches), then syncs itself
template <typename T, int DIR, int NB, int NTH, int N, int M, int TWOFF, bool FIRST, bool WIN, typename BAR, int R, int... Rest>
__device__ __forceinline__ void cta_fft_ip_dif_butlast(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo = 0, int hi = 0) {
    if constexpr (sizeof...(Rest) > 0) {
        if constexpr (!FIRST) bar.sync();
        cta_fft_ip_pass<T, R, DIR, true, NB, NTH, N, M, (FIRST && WIN)>(sm, tw + TWOFF, tid, lo, hi);
        cta_fft_ip_dif_butlast<T, DIR, NB, NTH, N, M / R, TWOFF + (R - 1) * (M / R), false, false, BAR, Rest...>(sm, tw, tid, bar);
    }
}
// inverse: all DIT passes but the innermost (the caller has done it), innermost first; a barrier
// PRECEDES every pass (so the first one also orders the caller's stage before the first pass)
template <typename T, int DIR, int NB, int NTH, int N, int M, int TWOFF, bool WIN, typename BAR, int R, int... Rest>
__device__ __forceinline__ void cta_fft_ip_dit_butinner(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo = 0, int hi = 0) {
    if constexpr (sizeof...(Rest) > 0) {
        cta_fft_ip_dit_butinner<T, DIR, NB, NTH, N, M / R, TWOFF + (R - 1) * (M / R), false, BAR, Rest...>(sm, tw, tid, bar);
        bar.sync();
        cta_fft_ip_pass<T, R, DIR, false, NB, NTH, N, M, WIN>(sm, tw + TWOFF, tid, lo, hi);
    }
}

// ---- the two OUTERMOST in-place passes without a group barrier between them.
// DIF pass 1 (radix R1, stride S1 = N/R1) butterfly n2 writes the elements n2 + m*S1; DIF pass 2 (radix R2 inside the
// blocks of M1 = N/R1, stride S2 = M1/R2) butterfly (blk, n2') reads blk*M1 + n2' + m*S2 -- all produced by the pass-1
// butterflies of the residue CLASS n2 = n2' (mod S2).  A team of TS lanes of ONE warp therefore takes a class through
// both passes (R2 butterflies of pass 1, __syncwarp, R1 butterflies of pass 2, per interleaved transform), and the S2
// classes are independent of each other: the 9-warp compute group of the column kernels no longer meets at a named
// barrier between these passes (ncu, round 2: 18 % of kernel A's stall samples sat on those barriers); the DIT
// inverse mirrors it.  Lanes of a team: b = lane % NB (transform), idx = lane / NB (+ TS/NB per round).  Conflict-free
// shared-memory accesses need the lane strides S2 and M1 (in 8*NB-byte granules) to be odd, which holds when R1
// carries every factor of two of N (2160 = 16 * 135): only such sequences take this path (ip_warplocal_ok).
#ifndef LPC_IP_WARPLOCAL
#define LPC_IP_WARPLOCAL 1
#endif
template <int N, int NB, int NTH, int R1, int R2> struct IpWl {
    static constexpr int M1 = N / R1, S2 = M1 / R2;
    static constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
    static constexpr int cost(int ts) {  // butterfly rounds of a warp for both passes with teams of ts lanes
        return (ts % NB != 0 || ts < NB) ? (1 << 30) : cdiv(S2, (NTH / 32) * (32 / ts)) * (cdiv(R2, ts / NB) + cdiv(R1, ts / NB));
    }
    static constexpr int TS = cost(32) <= cost(16) ? (cost(32) <= cost(8) ? 32 : 8) : (cost(16) <= cost(8) ? 16 : 8);
    static constexpr int plain = cdiv((N / R1) * NB, NTH) + cdiv((N / R2) * NB, NTH);  // rounds of the barrier-separated passes
    static constexpr bool ok = LPC_IP_WARPLOCAL && (32 % NB == 0) && (NTH % 32 == 0) && (M1 % 2 == 1) && (S2 % 2 == 1) &&
                               cost(TS) <= plain;
};
// one PHASE (0 = the pass executed first, 1 = the second) of class c for lane (b, li) of a team with LPB idx-lanes;
// host-callable: tests/csrc/test_fft_host.cpp walks the classes with exactly this function
template <typename T, int DIR, bool DIF, int NB, int N, int R1, int R2, bool WIN>
LPC_HD void fft_ip_outer2_phase(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int phase, int c, int b, int li, int LPB,
                                int lo, int hi) {
    constexpr int M1 = N / R1, S2 = M1 / R2, TWOFF2 = (R1 - 1) * (N / R1);
    if ((phase == 0) == DIF) {  // the radix-R1 pass over the whole transform: butterflies n2 = c + S2*idx, idx < R2
#pragma unroll 1
        for (int idx = li; idx < R2; idx += LPB) fft_ip_bfly<T, R1, DIR, DIF, WIN>(sm + b, NB, N, c + S2 * idx, tw, lo, hi);
    } else {                    // the radix-R2 pass inside the R1 blocks of M1: butterflies (blk = idx, n2' = c), idx < R1
#pragma unroll 1
        for (int idx = li; idx < R1; idx += LPB) fft_ip_bfly<T, R2, DIR, DIF>(sm + b, NB, M1, idx * S2 + c, tw + TWOFF2);
    }
}
template <typename T, int DIR, bool DIF, int NB, int NTH, int N, int R1, int R2, bool WIN>
__device__ __forceinline__ void cta_fft_ip_outer2(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, int lo, int hi) {
    using W = IpWl<N, NB, NTH, R1, R2>;
    constexpr int TS = W::TS, LPB = TS / NB, S2 = W::S2, NTEAM = (NTH / 32) * (32 / TS);
    const int team = tid / TS, tl = tid % TS, b = tl % NB, li = tl / NB;
    // the trip count is uniform over the CTA, so every lane of a warp reaches the __syncwarp (teams smaller than a warp
    // may run out of classes before their neighbours)
#pragma unroll 1
    for (int c0 = 0; c0 < S2; c0 += NTEAM) {
        const int c = c0 + team;
        if (c < S2) fft_ip_outer2_phase<T, DIR, DIF, NB, N, R1, R2, WIN>(sm, tw, 0, c, b, li, LPB, lo, hi);
        __syncwarp();
        if (c < S2) fft_ip_outer2_phase<T, DIR, DIF, NB, N, R1, R2, WIN>(sm, tw, 1, c, b, li, LPB, lo, hi);
    }
}
template <typename T, int DIR, int NB, int NTH, int N, bool WIN, typename BAR, int R1, int R2, int... Rest>
__device__ __forceinline__ void cta_fft_ip_dif_butlast_wl(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo, int hi) {
    cta_fft_ip_outer2<T, DIR, true, NB, NTH, N, R1, R2, WIN>(sm, tw, tid, lo, hi);
    constexpr int M2 = N / R1 / R2;
    cta_fft_ip_dif_butlast<T, DIR, NB, NTH, N, M2, (R1 - 1) * (N / R1) + (R2 - 1) * M2, false, false, BAR, Rest...>(sm, tw, tid, bar);
}
template <typename T, int DIR, int NB, int NTH, int N, bool WIN, typename BAR, int R1, int R2, int... Rest>
__device__ __forceinline__ void cta_fft_ip_dit_butinner_wl(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo, int hi) {
    constexpr int M2 = N / R1 / R2;
    cta_fft_ip_dit_butinner<T, DIR, NB, NTH, N, M2, (R1 - 1) * (N / R1) + (R2 - 1) * M2, false, BAR, Rest...>(sm, tw, tid, bar);
    bar.sync();
    cta_fft_ip_outer2<T, DIR, false, NB, NTH, N, R1, R2, WIN>(sm, tw, tid, lo, hi);
}
template <int N, int NB, int NTH, int... Rs> struct IpWlSel { static constexpr bool ok = false; };
template <int N, int NB, int NTH, int R1, int R2, int R3, int... Rest> struct IpWlSel<N, NB, NTH, R1, R2, R3, Rest...> {
    static constexpr bool ok = IpWl<N, NB, NTH, R1, R2>::ok;
};

template <typename SEQ> struct IpTraits;
template <int N_, int... Rs> struct IpTraits<FftSeq<N_, Rs...>> {
    using Seq = FftSeq<N_, Rs...>;
    static constexpr int N = N_;
    static constexpr int npass = sizeof...(Rs);
    static constexpr int last = Seq::radix[npass - 1];
    static constexpr int nb_last = N_ / last;  // butterflies of the innermost pass (blocks of `last` consecutive elements)
    static constexpr int compute_tw_len() {    // entries of the in-place twiddle tables (fill_fft_ip_twiddles)
        int len = 0, M = N_;
        for (int i = 0; i < npass; ++i) { len += (Seq::radix[i] - 1) * (M / Seq::radix[i]); M /= Seq::radix[i]; }
        return len;
    }
    static constexpr int tw_len = compute_tw_len();
    // WIN: the transform's input is zero outside [lo, hi) and only that range of the inverse's output is wanted
    template <typename T, int DIR, int NB, int NTH, typename BAR, bool WIN = false>
    static __device__ __forceinline__ void dif_butlast(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo = 0, int hi = 0) {
        if constexpr (IpWlSel<N_, NB, NTH, Rs...>::ok) cta_fft_ip_dif_butlast_wl<T, DIR, NB, NTH, N_, WIN, BAR, Rs...>(sm, tw, tid, bar, lo, hi);
        else cta_fft_ip_dif_butlast<T, DIR, NB, NTH, N_, N_, 0, true, WIN, BAR, Rs...>(sm, tw, tid, bar, lo, hi);
    }
    template <typename T, int DIR, int NB, int NTH, typename BAR, bool WIN = false>
    static __device__ __forceinline__ void dit_butinner(typename CxT<T>::type* sm, const typename CxT<T>::type* tw, int tid, BAR bar, int lo = 0, int hi = 0) {
        if constexpr (IpWlSel<N_, NB, NTH, Rs...>::ok) cta_fft_ip_dit_butinner_wl<T, DIR, NB, NTH, N_, WIN, BAR, Rs...>(sm, tw, tid, bar, lo, hi);
        else cta_fft_ip_dit_butinner<T, DIR, NB, NTH, N_, N_, 0, WIN, BAR, Rs...>(sm, tw, tid, bar, lo, hi);
    }
};
#endif  // __CUDACC__

}  // namespace lpc
