// lpc_fused_gd.cuh -- engine v2 for the gradient-descent family (gd.py:62-241): one iteration
//
//     g = deconvolve(convolve(x) - y)                    (gd.py:128-130; 4 FFT pairs in the reference)
//     x <- update(x, alpha*g)                            (gd.py:132-134, 183-188, 235-241)
//
// in FOUR kernels, every transform hand-written (lpc_fft.cuh) and fused with its neighbours:
//
//   k_gd_cols (forward) : column FFT  ->  * S*H/(2N)        -> column IFFT        [one tile / CTA]
//   k_gd_rows RESID     : row IFFT -> crop, - y, zero-pad   -> row FFT            [two rows / CTA]
//   k_gd_cols (adjoint) : column FFT  ->  * S*conj(H)/(2N)  -> column IFFT
//   k_gd_rows UPDATE    : row IFFT -> crop, step + projection + momentum -> zero-pad -> row FFT
//
// (the row FFT of the new estimate is the first half of the next iteration's convolve; an INIT
// launch does it once after reset).  Zero padding is exploited on both sides: only the h window
// rows are ever transformed along rows, stored, or loaded; the column kernel fills the other
// Hp - h rows of its shared-memory tile with zeros and stores only the window rows.
//
// Two real rows per complex row transform: rows 2r and 2r+1 of the window are packed as
// real + i*imag, so h/2 complex FFTs of length Wp serve h rows (any parity of Wp, odd h handled).
//
// "G-layout" of the single spectral buffer Tg (updated in place by both kernels):
//     complex index ((p*NTg + kx/4)*h + r)*4 + kx%4,   r = window row, NTg = ceil(Nh/4)
// i.e. 32-byte chunks per (tile,row) for the row kernel, one contiguous h*32-byte block per tile for
// the column kernel.  HBM traffic per iteration ~ 2*(H table) + 6*(h/Hp)*Sp + 5d = 0.55 GB at
// 3x1080x1920, against 10d + 6Sp = 0.85 GB canonical and ~2 GB for cuFFT + kernels.
#pragma once
#include "lpc_fft.cuh"
#include "lpc_fused.cuh"
#include "lpc_setup.cuh"
#include "lpc_util.cuh"

namespace lpc {

constexpr int LPC_GNC = 4;  // columns per tile of the G-layout
// LPC_GD_HINTS = 1: the column kernel reads its filter table with an evict-first L2 policy.  The two tables (H and conj H,
// 2 x 100 MB at 3x1080x1920) are streamed once per iteration each and used to flush the 125 MB of Tg / x / momentum / y
// that ARE reused out of the 126 MB L2: FISTA-300 13.45 -> 13.92 recon/s (five A/B runs, profiles/README.md).  Further
// hints on top (evict-first on the filter's L2 prefetch, on y, on the momentum; evict-last on Tg) changed nothing or cost.
#ifndef LPC_GD_HINTS
#define LPC_GD_HINTS 1
#endif
// LPC_GD_NA: the row kernel's single-use streams bypass L1 allocation (bit 0: stores of x / momentum / Tg -- measured slower, more spills in the update variants --, bit 1: loads of
// Tg / x / momentum / y), so that its two twiddle tables stay L1-resident (lpc_util.cuh)
#ifndef LPC_GD_NA
#define LPC_GD_NA 2
#endif

struct GdFusedK {
    int Hp, Wp, Nh, NTg;
    int h, w, sy, sx;
    int n_planes, DC, C, D, Dd;   // DC = number of transfer-function planes (D*C, or B*D*C with one PSF per sample)
    int plane0;  // first plane of this launch
    int NC, PS;  // columns of the constants table, 1 = one PSF per sample (lpc_const_col, lpc_setup.cuh)
};

enum { LPC_GD_INIT = 0, LPC_GD_RESID = 1, LPC_GD_UPD_GD = 2, LPC_GD_UPD_NESTEROV = 3, LPC_GD_UPD_FISTA = 4 };

__device__ __forceinline__ int gdf_data_plane(int p, const GdFusedK& k) {
    const int dc = k.D * k.C;
    const int b = p / dc, rem = p - b * dc;
    const int dep = rem / k.C, c = rem - dep * k.C;
    return (b * k.Dd + (k.Dd == 1 ? 0 : dep)) * k.C + c;
}

// ---- the element-wise stage of the row kernels, on the RL pixels j + m*NSL that butterfly j of the last inverse pass holds
//      (v[m].x: row ra, v[m].y: row rb); leaves the packed input of the forward transform in v
template <typename T, int RL, int NSL, int MODE>
__device__ __forceinline__ void gd_rows_elementwise(typename CxT<T>::type* v, int j, int sx, int w, bool has_b, int64_t ia, int64_t ib,
                                                    int ra, int rb, const T* __restrict__ dpl, T* __restrict__ x, T* __restrict__ mom,
                                                    T alpha, T scalar, int first) {
    using C = typename CxT<T>::type;
    if constexpr (MODE == LPC_GD_INIT || MODE == LPC_GD_RESID) {
#pragma unroll
        for (int m = 0; m < RL; ++m) {
            const int col = j + m * NSL - sx;  // window column of padded pixel j + m*NSL
            C z; z.x = T(0); z.y = T(0);       // everything outside the window is the zero padding
            if (col >= 0 && col < w) {
                if (MODE == LPC_GD_INIT) {
                    z.x = x[ia + col];
                    if (has_b) z.y = x[ib + col];
                } else {
                    if constexpr ((LPC_GD_NA & 2) != 0 && sizeof(T) == 4) {
                        z.x = v[m].x - ld4_nc_na(dpl + (int64_t)ra * w + col);  // crop(Hx) - y   (gd.py:129)
                        if (has_b) z.y = v[m].y - ld4_nc_na(dpl + (int64_t)rb * w + col);
                    } else {
                        z.x = v[m].x - dpl[(int64_t)ra * w + col];
                        if (has_b) z.y = v[m].y - dpl[(int64_t)rb * w + col];
                    }
                }
            }
            v[m] = z;
        }
    } else {
        // update modes read AND write x / mom: a load that follows a store to the same array cannot be hoisted
        // above it, so an element-by-element loop exposes one global-load latency per element (ncu, round 2: 48 %
        // of the kernel's samples on the long scoreboard).  Elements are therefore handled in chunks: all loads of
        // a chunk, then its arithmetic and stores -- gd_rows(update) 83 -> 66 us, FISTA-300 12.5 -> 13.4 recon/s.
        // (The same batching of SHARED-memory reads ahead of stores in the gather / separation loops was measured
        // slower -- B2 205 -> 226 us -- and is not used: there the interleaving feeds the LSU more evenly.)
        constexpr int CH = 4;
#pragma unroll
        for (int m0 = 0; m0 < RL; m0 += CH) {
            T xv[CH][2], pv[CH][2];
#pragma unroll
            for (int mm = 0; mm < CH; ++mm) {
                const int m = m0 + mm;
                xv[mm][0] = xv[mm][1] = pv[mm][0] = pv[mm][1] = T(0);
                if (m < RL) {
                    const int col = j + m * NSL - sx;
                    if (col >= 0 && col < w) {
                        if constexpr ((LPC_GD_NA & 2) != 0 && sizeof(T) == 4) {
                            xv[mm][0] = ld4_na(x + ia + col);
                            if (has_b) xv[mm][1] = ld4_na(x + ib + col);
                            if (MODE == LPC_GD_UPD_NESTEROV || (MODE == LPC_GD_UPD_FISTA && !first)) {
                                pv[mm][0] = ld4_na(mom + ia + col);
                                if (has_b) pv[mm][1] = ld4_na(mom + ib + col);
                            }
                        } else {
                            xv[mm][0] = x[ia + col];
                            if (has_b) xv[mm][1] = x[ib + col];
                            if (MODE == LPC_GD_UPD_NESTEROV || (MODE == LPC_GD_UPD_FISTA && !first)) {
                                pv[mm][0] = mom[ia + col];
                                if (has_b) pv[mm][1] = mom[ib + col];
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int mm = 0; mm < CH; ++mm) {
                const int m = m0 + mm;
                if (m < RL) {
                    const int col = j + m * NSL - sx;
                    C z; z.x = T(0); z.y = T(0);
                    if (col >= 0 && col < w) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            if (s == 1 && !has_b) break;
                            const int64_t o = (s == 0 ? ia : ib) + col;
                            const T gs = (s == 0) ? v[m].x : v[m].y;
                            const T xo = xv[mm][s];
                            T xn;
                            if (MODE == LPC_GD_UPD_GD) {  // gd.py:132-134
                                const T u = xo - alpha * gs;
                                xn = u > T(0) ? u : T(0);
                            } else if (MODE == LPC_GD_UPD_NESTEROV) {  // gd.py:183-188
                                const T p_prev = pv[mm][s];
                                const T p_new = scalar * p_prev - alpha * gs;
                                if constexpr ((LPC_GD_NA & 1) != 0 && sizeof(T) == 4) st4_na(mom + o, p_new); else mom[o] = p_new;
                                const T u = xo + (-scalar * p_prev + (T(1) + scalar) * p_new);
                                xn = u > T(0) ? u : T(0);
                            } else {  // FISTA, gd.py:235-241
                                const T zz = xo - alpha * gs;
                                const T xk = zz > T(0) ? zz : T(0);
                                const T prev = first ? zz : pv[mm][s];
                                if constexpr ((LPC_GD_NA & 1) != 0 && sizeof(T) == 4) st4_na(mom + o, xk); else mom[o] = xk;
                                xn = xk + scalar * (xk - prev);
                            }
                            if constexpr ((LPC_GD_NA & 1) != 0 && sizeof(T) == 4) st4_na(x + o, xn); else x[o] = xn;
                            if (s == 0) z.x = xn; else z.y = xn;
                        }
                    }
                    v[m] = z;
                }
            }
        }
    }
}

// ---- separate the packed pair held in W (natural order) and store the two half-spectrum rows ra, ra+1 (no 1/2: folded
//      into the filter); two lanes per 32-byte chunk, 16-byte stores
template <typename T, int Wp, int PF>
__device__ __forceinline__ void gd_rows_separate_store(const typename CxT<T>::type* W, typename CxT<T>::type* Tp, int64_t tile_stride,
                                                       int ra, bool has_b, int tid) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS, NC = LPC_GNC, Nh = Wp / 2 + 1, NTg = (Nh + NC - 1) / NC;
    for (int g = tid; g < NTg * 2; g += NTH) {
        const int t = g >> 1, hf = g & 1;
        Vec16<T> oa, ob;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int kx = t * NC + hf * 2 + c;
            oa.v[2 * c] = oa.v[2 * c + 1] = ob.v[2 * c] = ob.v[2 * c + 1] = T(0);
            if (kx < Nh) {
                const C zk = W[fft_pad(kx, PF)];
                const C zm = W[fft_pad(kx == 0 ? 0 : Wp - kx, PF)];
                oa.v[2 * c] = zk.x + zm.x; oa.v[2 * c + 1] = zk.y - zm.y;   // Z[k] + conj(Z[-k])
                ob.v[2 * c] = zk.y + zm.y; ob.v[2 * c + 1] = zm.x - zk.x;   // (Z[k] - conj(Z[-k])) / i
            }
        }
        C* da = Tp + (int64_t)t * tile_stride + (int64_t)ra * NC + hf * 2;
        if constexpr ((LPC_GD_NA & 1) != 0 && sizeof(T) == 4) {
            st16_na(reinterpret_cast<T*>(da), oa);
            if (has_b) st16_na(reinterpret_cast<T*>(da + NC), ob);
        } else {
            st16(reinterpret_cast<T*>(da), oa);
            if (has_b) st16(reinterpret_cast<T*>(da + NC), ob);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rows: one CTA per pair of window rows (ra = 2*pair, rb = ra + 1) of one plane.   grid (ceil(h/2), planes)
//   x, mom, data are planar UNPADDED [planes][h][w]; consts[3C + c] = alpha_c (gd.py:108-112)
//   scalar: Nesterov mu (gd.py:183-188) or FISTA beta = (t_k - 1)/t_{k+1} (gd.py:238-239)
//   first : FISTA only -- right after reset() the reference's previous iterate aliases the stepped
//           estimate (gd.py:233,236), mirrored
// The inverse transform runs SEQ, the forward one the REVERSED sequence, so that the last inverse
// pass, the element-wise stage and the first forward pass share registers: butterfly j holds pixels
// j + m*(Wp/RL), m < RL (coalesced 4-byte accesses to x / mom / y per m).
// twI: tables of SEQ (inverse), twF: tables of the reversed sequence (forward).
// ---------------------------------------------------------------------------------------------
template <typename T, typename SEQ, int MODE>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 4)
k_gd_rows(GdFusedK f, const typename CxT<T>::type* __restrict__ twI, const typename CxT<T>::type* __restrict__ twF,
          typename CxT<T>::type* __restrict__ Tg, const T* __restrict__ data, T* __restrict__ x, T* __restrict__ mom,
          const T* __restrict__ consts, T scalar, int first) {
    using C = typename CxT<T>::type;
    using TR = SeqTraits<SEQ>;
    using REV = typename TR::Rev;
    constexpr int NTH = LPC_ROWS_THREADS, NC = LPC_GNC, Wp = SEQ::N, Nh = Wp / 2 + 1;
    constexpr int PI = SEQ::padsh, PF = REV::padsh, RL = TR::last, NSL = TR::ns_last;
    constexpr int NTg = (Nh + NC - 1) / NC;
    static_assert(sizeof(T) == 4 && NC == 4, "the fused engine is float-only with 32-byte chunks");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* W = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int p = f.plane0 + blockIdx.y;
    const int ra = 2 * blockIdx.x, rb = ra + 1;
    const int h = f.h, w = f.w, sx = f.sx;
    const bool has_b = rb < h;
    const int64_t tile_stride = (int64_t)h * NC;  // complex elements per tile
    C* Tp = Tg + (int64_t)p * NTg * tile_stride;

    if (MODE != LPC_GD_INIT) {
        // gather the two half-spectrum rows, build the full-length packed row Y_a + i*Y_b.  Two lanes per 32-byte
        // chunk, each taking two columns (16 bytes) of row a and of row b: a warp instruction touches 16 lines
        // instead of the 32 of a chunk-per-lane load (~1.6 LSU wavefronts per lane under ncu)
        for (int g = tid; g < NTg * 2; g += NTH) {
            const int t = g >> 1, hf = g & 1;
            const C* sa = Tp + (int64_t)t * tile_stride + (int64_t)ra * NC + hf * 2;
            const Vec16<T> va = (LPC_GD_NA & 2) ? ld16_nc_na(reinterpret_cast<const T*>(sa)) : ld16_nc(reinterpret_cast<const T*>(sa));
            Vec16<T> vb;
            vb.v[0] = vb.v[1] = vb.v[2] = vb.v[3] = T(0);
            if (has_b) vb = (LPC_GD_NA & 2) ? ld16_nc_na(reinterpret_cast<const T*>(sa + NC)) : ld16_nc(reinterpret_cast<const T*>(sa + NC));
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int kx = t * NC + hf * 2 + c;
                const T yax = va.v[2 * c], yay = va.v[2 * c + 1], ybx = vb.v[2 * c], yby = vb.v[2 * c + 1];
                if (kx < Nh) {
                    C y; y.x = yax - yby; y.y = yay + ybx;
                    W[fft_pad(kx, PI)] = y;
                    const int m = Wp - kx;
                    if (kx >= 1 && m >= Nh) {
                        C ym; ym.x = yax + yby; ym.y = ybx - yay;
                        W[fft_pad(m, PI)] = ym;
                    }
                }
            }
        }
        __syncthreads();
        TR::template run_butlast<T, +1, 1, NTH>(W, twI, tid);  // inverse FFT, all passes but the last
    }

    // ---- last inverse pass -> element-wise stage -> first forward pass, in registers
    const int64_t ia = ((int64_t)p * h + ra) * w, ib = ia + w;
    const T alpha = (MODE >= LPC_GD_UPD_GD) ? consts[3 * f.NC + lpc_const_col(p, f.D * f.C, f.C, f.PS)] : T(0);
    const T* dpl = data + (int64_t)gdf_data_plane(p, f) * h * w;
    constexpr int MAXB = (NSL + NTH - 1) / NTH;
    C v[MAXB][RL];
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int j = tid + it * NTH;
        if (MAXB * NTH == NSL || j < NSL) {
            if (MODE != LPC_GD_INIT) {
                fft_bfly_load<T, RL>(W, 0, 1, PI, Wp, j, v[it]);
                fft_bfly_compute<T, RL, +1>(v[it], twI + TR::tw_off_last, NSL, j);
            }
            gd_rows_elementwise<T, RL, NSL, MODE>(v[it], j, sx, w, has_b, ia, ib, ra, rb, dpl, x, mom, alpha, scalar, first);
            Dft<T, RL, -1>::run(v[it]);  // first pass of the reversed sequence: Ns = 1, no twiddles
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int j = tid + it * NTH;
        if (MAXB * NTH == NSL || j < NSL) fft_bfly_store<T, RL>(W, 0, 1, PF, 1, j, v[it]);
    }
    __syncthreads();
    SeqTraits<REV>::template run_butfirst<T, -1, 1, NTH>(W, twF, tid);

    gd_rows_separate_store<T, Wp, PF>(W, Tp, tile_stride, ra, has_b, tid);
}

// ---------------------------------------------------------------------------------------------
// columns: one CTA per tile of LPC_GNC columns, in place.   grid (NTg, planes)
//   F: filter in tile layout [DC][NTg][Hp][4] = S*H/(2*Hp*Wp) (convolve) or S*conj(H)/(2*Hp*Wp)
//   (deconvolve; the same shift S, rfft_convolve.py:195)
//   forward FFT with SEQ, inverse with the reversed sequence; the last forward pass, the filter
//   multiply and the first inverse pass share registers (see k_admm_cols).
// ---------------------------------------------------------------------------------------------
template <typename T, typename SEQ, int NTH, int MINB>
__global__ void __launch_bounds__(NTH, MINB)
k_gd_cols(GdFusedK f, const typename CxT<T>::type* __restrict__ twF, const typename CxT<T>::type* __restrict__ twI,
          typename CxT<T>::type* __restrict__ Tg, const typename CxT<T>::type* __restrict__ F) {
    using C = typename CxT<T>::type;
    using TR = SeqTraits<SEQ>;
    using REV = typename TR::Rev;
    constexpr int NC = LPC_GNC, Hp = SEQ::N, PF = SEQ::padsh, PI = REV::padsh, RL = TR::last, NSL = TR::ns_last;
    constexpr int n = Hp * NC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, p = f.plane0 + blockIdx.y;
    const int q = p % f.DC;
    const int h = f.h, sy = f.sy;
    C* blk = Tg + ((int64_t)p * f.NTg + tile) * (int64_t)h * NC;
    const C* Fq = F + ((int64_t)q * f.NTg + tile) * (int64_t)n;

    for (int e = tid * 16; e < n; e += NTH * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(Fq + e));
    // window rows from HBM (16-byte loads), zeros elsewhere
    for (int v = tid; v < n / 2; v += NTH) {
        const int row = v / (NC / 2);
        C a, b;
        a.x = a.y = b.x = b.y = T(0);
        const int r = row - sy;
        if (r >= 0 && r < h) {
            const Vec16<T> d = ld16_nc(reinterpret_cast<const T*>(blk) + 4 * (v - sy * (NC / 2)));
            a.x = d.v[0]; a.y = d.v[1]; b.x = d.v[2]; b.y = d.v[3];
        }
        sm[fft_pad(2 * v, PF)] = a;
        sm[fft_pad(2 * v + 1, PF)] = b;
    }
    __syncthreads();
    TR::template run_butlast<T, -1, NC, NTH>(sm, twF, tid);

    constexpr int total = NSL * NC;
    constexpr int MAXB = (total + NTH - 1) / NTH;
    C v[MAXB][RL];
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int g = tid + it * NTH;
        if (MAXB * NTH == total || g < total) {
            const int c = g % NC, j = g / NC;
            fft_bfly_load<T, RL>(sm, c, NC, PF, Hp, j, v[it]);
            fft_bfly_compute<T, RL, -1>(v[it], twF + TR::tw_off_last, NSL, j);
#pragma unroll
            for (int m = 0; m < RL; ++m) v[it][m] = cx_mul(v[it][m], __ldg(Fq + (j + m * NSL) * NC + c));
            Dft<T, RL, +1>::run(v[it]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MAXB; ++it) {
        const int g = tid + it * NTH;
        if (MAXB * NTH == total || g < total) fft_bfly_store<T, RL>(sm, g % NC, NC, PI, 1, g / NC, v[it]);
    }
    __syncthreads();
    SeqTraits<REV>::template run_butfirst<T, +1, NC, NTH>(sm, twI, tid);
    for (int v2 = tid; v2 < h * (NC / 2); v2 += NTH) {
        const int e = 2 * (v2 + sy * (NC / 2));
        const C a = sm[fft_pad(e, PI)], b = sm[fft_pad(e + 1, PI)];
        Vec16<T> o;
        o.v[0] = a.x; o.v[1] = a.y; o.v[2] = b.x; o.v[3] = b.y;
        st16(reinterpret_cast<T*>(blk) + 4 * v2, o);
    }
}

// filter tables in tile layout: F[adj][q][tile][ky][c] = S(ky,kx) * (adj ? conj(H) : H) / (2*Hp*Wp)
template <typename T>
__global__ void k_gd_tile_filter(const typename CxT<T>::type* __restrict__ H, const typename CxT<T>::type* __restrict__ Sy,
                                 const typename CxT<T>::type* __restrict__ Sx, typename CxT<T>::type* __restrict__ F,
                                 int DC, int Hp, int Wp, int Nh, int NTg) {
    using C = typename CxT<T>::type;
    const int64_t n = (int64_t)DC * NTg * Hp * LPC_GNC;
    const T scale = (T)(0.5 / ((double)Hp * (double)Wp));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % LPC_GNC);
        int64_t r = i / LPC_GNC;
        const int ky = (int)(r % Hp); r /= Hp;
        const int tile = (int)(r % NTg);
        const int q = (int)(r / NTg);
        const int kx = tile * LPC_GNC + c;
        C fwd, adj;
        fwd.x = fwd.y = adj.x = adj.y = T(0);
        if (kx < Nh) {
            C hq = H[((int64_t)q * Hp + ky) * Nh + kx];
            const C S = cx_mul(Sy[ky], Sx[kx]);
            fwd = cx_mul(S, hq);
            hq.y = -hq.y;
            adj = cx_mul(S, hq);
            fwd.x *= scale; fwd.y *= scale; adj.x *= scale; adj.y *= scale;
        }
        F[i] = fwd;
        F[n + i] = adj;
    }
}

}  // namespace lpc
