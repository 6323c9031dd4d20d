// lpc_fused.cuh -- engine v2: the whole ADMM iteration in TWO kernels, transforms included.
//
//   k_admm_rows  (B): row IFFT -> K1 spatial primal/dual update -> row FFT      [per band of rows]
//   k_admm_cols  (A): column FFT -> K2 spectral update -> column IFFT           [per tile of columns]
//
// V and HV are never stored: they only exist inside B, between its inverse and forward row
// transforms.  What persists between iterations is
//   duals  xi', rho', eta' (spatial, planar)                                     4P
//   Tin    row spectra of (V, HV)/Wp  after A's column IFFT                       2Sp
//   Tout   row spectra of 2*(a, xi')  after B's row FFT                           2Sp
// so one iteration moves  B: 2Sp + 4P + P(eta'0 halo, L2) + d in, 4P + 2Sp out;  A: 2Sp + H + R in,
// 2Sp out  =  8P + 9.5Sp + d  ~ 1.77 GB at 3x1080x1920 (v1 with cuFFT: ~3.3 GB; SURVEY's
// canonical figure 16P + 9.5Sp + d = 2.56 GB counts a, V, HV materialised).
//
// Two real rows per complex transform: B packs z = a + i*xi' (forward) and Y = Vhat + i*HVhat
// (inverse) so every row transform is ONE complex FFT of length Wp (any parity of Wp), with
// the separation  A_row[k] = (Z[k] + conj(Z[-k]))/2,  Xi_row[k] = (Z[k] - conj(Z[-k]))/(2i)
// done in shared memory.  Only the half spectrum k in [0, Wp/2] is stored.
//
// "T-layout" of Tin / Tout (tiled-transposed half spectrum), chosen so that BOTH kernels see
// whole 32-byte sectors:  complex index
//     ((p*NT + kx/NCT)*Hp + ky)*(2*NCT) + slot*NCT + kx%NCT,     NT = ceil(Nh/NCT), NCT = 2
// A reads/writes one tile (all ky of NCT columns, both slots) as ONE contiguous 2*NCT*Hp*8-byte
// block; B touches a 2*NCT*8 = 32-byte chunk per (tile, row).
//
// Reference semantics reproduced: admm.py:232-329 (see lpc_admm.cuh for the per-pixel algebra,
// which is shared), rfft_convolve.py:133-223 (H, conj(H), ifftshift as the phase S).
#pragma once
#include "lpc_admm.cuh"
#include "lpc_fft.cuh"
#include "lpc_util.cuh"

namespace lpc {

constexpr int LPC_NCT = 2;           // columns per tile and slot
constexpr int LPC_ROWS_THREADS = 256;
constexpr int LPC_COLS_THREADS = 512;

struct FusedK {
    int Hp, Wp, Nh, NT;      // NT tiles of LPC_NCT columns cover Nh = Wp/2+1
    int n_planes, DC;
    int rows_per_band, bands_per_plane;
    int padsh_row, padsh_col;
};

enum { LPC_ROWS_FULL = 0, LPC_ROWS_IFFT_ONLY = 1, LPC_ROWS_FFT_ONLY = 2 };

// ---------------------------------------------------------------------------------------------
// B: one CTA marches down a band of rows of one plane.
// shared memory:  W  complex[padded Wp]  (row transform, in place)
//                 Vr real[3][Wp]         rolling window of V rows i-1, i, i+1
//                 HVr real[2][Wp]        HV rows i, i+1
// MODE FULL      : the ADMM row step.
// MODE IFFT_ONLY : only materialise V, HV (planar spatial, VH[2][planes][Hp][Wp]) -- _form_image / debug.
// MODE FFT_ONLY  : only transform given V, HV into Tin (warm start, admm.py:154-155,172-176).
// ---------------------------------------------------------------------------------------------
template <typename T, int PPT, int MODE>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 2)
k_admm_rows(FusedK f, AdmmK<T> k, FftPlan plan, const typename CxT<T>::type* __restrict__ tw,
            const typename CxT<T>::type* __restrict__ Tin, typename CxT<T>::type* __restrict__ Tout,
            const T* __restrict__ data, T* __restrict__ xi, T* __restrict__ rho, const T* __restrict__ eta_in,
            T* __restrict__ eta_out, T* __restrict__ VH, T sep_scale) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS;
    constexpr int NCT = LPC_NCT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int Wp = f.Wp, Hp = f.Hp, Nh = f.Nh, NT = f.NT;
    const int padsh = plan.padsh;
    const int wlen = fft_padded_len(Wp, padsh);
    C* W = reinterpret_cast<C*>(smem_raw);
    const int wp4 = (Wp + 3) & ~3;  // keep the real rows 16-byte aligned
    T* Vr = reinterpret_cast<T*>(W + ((wlen + 1) & ~1));
    T* HVr = Vr + 3 * wp4;

    const int p = blockIdx.x / f.bands_per_plane;
    const int band = blockIdx.x - p * f.bands_per_plane;
    const int i0 = band * f.rows_per_band;
    if (i0 >= Hp) return;
    const int i1 = (i0 + f.rows_per_band < Hp) ? i0 + f.rows_per_band : Hp;

    const int64_t plane = (int64_t)Hp * Wp;
    const int64_t base = (int64_t)p * plane;
    const int64_t tile_stride = (int64_t)Hp * (2 * NCT);  // complex elements per tile
    const C* TinP = Tin + (int64_t)p * NT * tile_stride;
    C* ToutP = Tout + (int64_t)p * NT * tile_stride;

    // ---- gather one row of Tin into W as the full-length complex row  Y = Vhat + i*HVhat
    //      (Hermitian extension for kx > Wp/2), inverse FFT, scatter to the real rows
    auto load_ifft_row = [&](int row, T* Vdst, T* HVdst) {
        for (int t = tid; t < NT; t += NTH) {
            const C* src = TinP + (int64_t)t * tile_stride + (int64_t)row * (2 * NCT);
            C vh[NCT], hh[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) { vh[c] = src[c]; hh[c] = src[NCT + c]; }
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                const int kx = t * NCT + c;
                if (kx < Nh) {
                    C y; y.x = vh[c].x - hh[c].y; y.y = vh[c].y + hh[c].x;
                    W[fft_pad(kx, padsh)] = y;
                    const int m = Wp - kx;
                    if (kx >= 1 && m >= Nh) {
                        C ym; ym.x = vh[c].x + hh[c].y; ym.y = hh[c].x - vh[c].y;
                        W[fft_pad(m, padsh)] = ym;
                    }
                }
            }
        }
        __syncthreads();
        cta_fft<T, +1, PPT>(W, plan, tw, 1, 0, 1, tid, NTH);
        for (int j = tid; j < Wp; j += NTH) {
            const C c = W[fft_pad(j, padsh)];
            Vdst[j] = c.x;
            if (HVdst) HVdst[j] = c.y;
        }
        __syncthreads();
    };

    // ---- forward FFT of the row held in W (z = a + i*xi'), separate the two real spectra and
    //      store the half spectrum to row `row` of Tdst (T-layout)
    auto fft_store_row = [&](int row, C* Tdst) {
        cta_fft<T, -1, PPT>(W, plan, tw, 1, 0, 1, tid, NTH);
        for (int t = tid; t < NT; t += NTH) {
            C a[NCT], x[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                const int kx = t * NCT + c;
                a[c].x = a[c].y = x[c].x = x[c].y = T(0);
                if (kx < Nh) {
                    const C zk = W[fft_pad(kx, padsh)];
                    const C zm = W[fft_pad(kx == 0 ? 0 : Wp - kx, padsh)];
                    a[c].x = (zk.x + zm.x) * sep_scale;   // Z[k] + conj(Z[-k])
                    a[c].y = (zk.y - zm.y) * sep_scale;
                    x[c].x = (zk.y + zm.y) * sep_scale;   // (Z[k] - conj(Z[-k])) / i
                    x[c].y = (zm.x - zk.x) * sep_scale;
                }
            }
            C* dst = Tdst + (int64_t)t * tile_stride + (int64_t)row * (2 * NCT);
#pragma unroll
            for (int c = 0; c < NCT; ++c) { dst[c] = a[c]; dst[NCT + c] = x[c]; }
        }
        __syncthreads();
    };

    if (MODE == LPC_ROWS_IFFT_ONLY) {
        for (int i = i0; i < i1; ++i) {
            load_ifft_row(i, Vr, HVr);
            T* Vout = VH + base + (int64_t)i * Wp;
            T* HVout = VH + (int64_t)f.n_planes * plane + base + (int64_t)i * Wp;
            for (int j = tid; j < Wp; j += NTH) { Vout[j] = Vr[j]; HVout[j] = HVr[j]; }
            __syncthreads();
        }
        return;
    }
    if (MODE == LPC_ROWS_FFT_ONLY) {
        for (int i = i0; i < i1; ++i) {
            const T* Vin = VH + base + (int64_t)i * Wp;
            const T* HVin = VH + (int64_t)f.n_planes * plane + base + (int64_t)i * Wp;
            for (int j = tid; j < Wp; j += NTH) { C z; z.x = Vin[j]; z.y = HVin[j]; W[fft_pad(j, padsh)] = z; }
            __syncthreads();
            fft_store_row(i, const_cast<C*>(TinP));
        }
        return;
    }

    // ------------------------------------------------------------------ FULL: the ADMM row step
    const int64_t eta1_off = (int64_t)k.n_planes * plane;
    const T* dp = data + (int64_t)data_plane_of(p, k) * k.dplane;
    int sp = 0, sc = 1, sn = 2;  // ring slots of rows i-1, i, i+1
    load_ifft_row(i0 == 0 ? Hp - 1 : i0 - 1, Vr + sp * wp4, nullptr);
    load_ifft_row(i0, Vr + sc * wp4, HVr);
    int hc = 0;  // HVr slot of row i
    for (int i = i0; i < i1; ++i) {
        const int id = (i == Hp - 1) ? 0 : i + 1;
        load_ifft_row(id, Vr + sn * wp4, HVr + (hc ^ 1) * wp4);
        const T* Vp = Vr + sp * wp4;
        const T* Vc = Vr + sc * wp4;
        const T* Vn = Vr + sn * wp4;
        const T* HVc = HVr + hc * wp4;
        const int64_t ro = base + (int64_t)i * Wp;
        const int64_t rd = base + (int64_t)id * Wp;
        const bool row_in = (i >= k.sy) && (i < k.sy + k.h);
        for (int j = tid; j < Wp; j += NTH) {
            const int jl = (j == 0) ? Wp - 1 : j - 1;
            const int jr = (j == Wp - 1) ? 0 : j + 1;
            const T v = Vc[j];
            T e0n, e1n, dummy;
            const T t0 = tv_component(Vp[j] - v, eta_in[ro + j], k, e0n);
            const T t1 = tv_component(Vc[jl] - v, eta_in[eta1_off + ro + j], k, e1n);
            const T t0d = tv_component(v - Vn[j], eta_in[rd + j], k, dummy);
            const T t1r = tv_component(v - Vc[jr], eta_in[eta1_off + ro + jr], k, dummy);
            const bool inw = row_in && (j >= k.sx) && (j < k.sx + k.w);
            T y = T(0);
            if (inw) y = dp[(int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0)];
            T xin, rhon;
            point_update(v, HVc[j], y, inw, xi[ro + j], rho[ro + j], k, xin, rhon);
            xi[ro + j] = xin;
            rho[ro + j] = rhon;
            eta_out[ro + j] = e0n;
            eta_out[eta1_off + ro + j] = e1n;
            C z;
            z.x = -rhon + ((t0d - t0) + (t1r - t1));
            z.y = xin;
            W[fft_pad(j, padsh)] = z;
        }
        __syncthreads();
        fft_store_row(i, ToutP);
        const int tmp = sp; sp = sc; sc = sn; sn = tmp;
        hc ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// A: one CTA owns one tile: LPC_NCT columns x 2 slots = 4 interleaved column transforms of
// length Hp, resident in shared memory from the load to the store.
//     forward column FFT of (2A_row, 2Xi_row)
//     Vhat  = Rn * (A - S*conj(H)*Xi)        HVhat = S*H*Vhat        (Rn carries 1/(2*Hp*Wp))
//     inverse column FFT -> row spectra of (V, HV)/Wp
// Ht: [DC][NT][Hp][NCT] complex, Rt: same shape real (tile layout, built once per PSF).
// grid (NT, planes)
// ---------------------------------------------------------------------------------------------
template <typename T, int PPT>
__global__ void __launch_bounds__(LPC_COLS_THREADS)
k_admm_cols(FusedK f, FftPlan plan, const typename CxT<T>::type* __restrict__ tw,
            const typename CxT<T>::type* __restrict__ Tsrc, typename CxT<T>::type* __restrict__ Tdst,
            const typename CxT<T>::type* __restrict__ Ht, const T* __restrict__ Rt,
            const typename CxT<T>::type* __restrict__ Sy, const typename CxT<T>::type* __restrict__ Sx) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_COLS_THREADS;
    constexpr int NCT = LPC_NCT;
    constexpr int NB = 2 * NCT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int Hp = f.Hp, NT = f.NT;
    const int padsh = plan.padsh;
    const int tile = blockIdx.x;
    const int p = blockIdx.y;
    const int q = p % f.DC;
    const int n = Hp * NB;
    const C* src = Tsrc + ((int64_t)p * NT + tile) * n;
    C* dst = Tdst + ((int64_t)p * NT + tile) * n;

    for (int e = tid; e < n; e += NTH) sm[fft_pad(e, padsh)] = src[e];
    __syncthreads();
    cta_fft<T, -1, PPT>(sm, plan, tw, NB, 1, NB, tid, NTH);

    const C* Hq = Ht + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT;
    const T* Rq = Rt + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT;
    for (int e = tid; e < Hp * NCT; e += NTH) {
        const int ky = e / NCT, c = e - ky * NCT;
        const int kx = tile * NCT + c;
        const int ia = fft_pad(ky * NB + c, padsh), ix = fft_pad(ky * NB + NCT + c, padsh);
        const C A = sm[ia];
        const C Xi = sm[ix];
        const C Hh = Hq[e];
        const T Rn = Rq[e];
        const C S = cx_mul(Sy[ky], Sx[kx < f.Nh ? kx : 0]);
        const C Hs = cx_mul(S, Hh);
        C Hc = Hh; Hc.y = -Hc.y;
        const C Hsa = cx_mul(S, Hc);
        const C m = cx_mul(Hsa, Xi);
        C Vh; Vh.x = Rn * (A.x - m.x); Vh.y = Rn * (A.y - m.y);
        sm[ia] = Vh;
        sm[ix] = cx_mul(Hs, Vh);
    }
    __syncthreads();
    cta_fft<T, +1, PPT>(sm, plan, tw, NB, 1, NB, tid, NTH);
    for (int e = tid; e < n; e += NTH) dst[e] = sm[fft_pad(e, padsh)];
}

// H, R (planar [DC][Hp][Nh]) -> tile layout [DC][NT][Hp][NCT]; Rn gets the 1/2 of the row separation
template <typename T>
__global__ void k_tile_HR(const typename CxT<T>::type* __restrict__ H, const T* __restrict__ R,
                          typename CxT<T>::type* __restrict__ Ht, T* __restrict__ Rt, int DC, int Hp, int Nh, int NT) {
    using C = typename CxT<T>::type;
    const int64_t n = (int64_t)DC * NT * Hp * LPC_NCT;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % LPC_NCT);
        int64_t r = i / LPC_NCT;
        const int ky = (int)(r % Hp); r /= Hp;
        const int tile = (int)(r % NT);
        const int q = (int)(r / NT);
        const int kx = tile * LPC_NCT + c;
        C h; h.x = h.y = T(0);
        T rr = T(0);
        if (kx < Nh) {
            h = H[((int64_t)q * Hp + ky) * Nh + kx];
            rr = R[((int64_t)q * Hp + ky) * Nh + kx] * T(0.5);
        }
        Ht[i] = h;
        Rt[i] = rr;
    }
}

}  // namespace lpc
