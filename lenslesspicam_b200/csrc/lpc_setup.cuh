// lpc_setup.cuh -- layout conversion (channels-last <-> planar) and the once-per-PSF
// kernels: H scaling, shift tables, ADMM R_divmat, GD step size / initial value.
//
// Planar layout (DESIGN.md "Data layout in HBM"): plane index p = (b*D + dep)*C + c,
// each plane a contiguous [rows][pitch] array.  The reference's arrays are
// channels-last [B, D, h, w, C] (recon.py:223-224).
#pragma once
#include "lpc_util.cuh"

namespace lpc {

// element strides of a channels-last caller array for (B, D, h, w, C); stride 0 broadcasts
struct Strides5 { int64_t b, d, y, x, c; };

// ---------------------------------------------------------------------------
// K5a: channels-last -> planar window.  dst[p][y+oy][x+ox] = src[b,dep,y,x,c] (- bg, clamp >= 0)
// Replaces RealFFTConvolve2D._pad (rfft_convolve.py:84-100) as a layout step; the
// optional background subtract + clamp is recon.py:553-555.
// grid: x over (x*C + c), y over rows, z over (b*Dn + dep)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_pack(const T* __restrict__ src, Strides5 ss, const T* __restrict__ bg, Strides5 bs,
                       T* __restrict__ dst, T* __restrict__ echo, int Dn, int C, int h, int w,
                       int64_t plane_stride, int pitch, int oy, int ox) {
    const int xc = blockIdx.x * blockDim.x + threadIdx.x;
    if (xc >= w * C) return;
    const int y = blockIdx.y;
    const int bd = blockIdx.z;
    const int b = bd / Dn, dep = bd - b * Dn;
    const int x = xc / C, c = xc - x * C;
    T v = src[b * ss.b + dep * ss.d + y * ss.y + x * ss.x + c * ss.c];
    if (bg != nullptr) {
        v -= bg[b * bs.b + dep * bs.d + y * bs.y + x * bs.x + c * bs.c];
        if (v < T(0)) v = T(0);
    }
    if (echo != nullptr) echo[((int64_t)bd * h + y) * ((int64_t)w * C) + xc] = v;
    dst[((int64_t)bd * C + c) * plane_stride + (int64_t)(y + oy) * pitch + (x + ox)] = v;
}

// ---------------------------------------------------------------------------
// K5b: planar window -> channels-last contiguous [B*Dn, h, w, C], optional clamp >= 0.
// Replaces RealFFTConvolve2D._crop (rfft_convolve.py:79-82) + ADMM._form_image's clamp
// (admm.py:331-338) / gd.non_neg (gd.py:41-59).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_unpack(const T* __restrict__ src, int64_t plane_stride, int pitch, int oy, int ox,
                         T* __restrict__ dst, int C, int h, int w, int clamp) {
    const int xc = blockIdx.x * blockDim.x + threadIdx.x;
    if (xc >= w * C) return;
    const int y = blockIdx.y;
    const int bd = blockIdx.z;
    const int x = xc / C, c = xc - x * C;
    T v = src[((int64_t)bd * C + c) * plane_stride + (int64_t)(y + oy) * pitch + (x + ox)];
    if (clamp && v < T(0)) v = T(0);
    dst[((int64_t)bd * h + y) * ((int64_t)w * C) + xc] = v;
}

// complex planar [DC][Hp][Nh] -> reference layout [D, Hp, Nh, C] (interleaved re,im): debug / parity of _H
template <typename T>
__global__ void k_unpack_complex(const typename Cx<T>::type* __restrict__ src, typename Cx<T>::type* __restrict__ dst,
                                 int D, int C, int Hp, int Nh) {
    const int64_t n = (int64_t)D * Hp * Nh * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int64_t r = i / C;
        int kx = (int)(r % Nh); r /= Nh;
        int ky = (int)(r % Hp);
        int dep = (int)(r / Hp);
        dst[i] = src[(((int64_t)dep * C + c) * Hp + ky) * Nh + kx];
    }
}

// ---------------------------------------------------------------------------
// K4a: H *= scale (rfft2(..., norm=ortho|forward), rfft_convolve.py:121-123)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_scale_complex(typename Cx<T>::type* __restrict__ H, int64_t n, T scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        typename Cx<T>::type v = H[i];
        v.x *= scale; v.y *= scale;
        H[i] = v;
    }
}

// ---------------------------------------------------------------------------
// K4b: ifftshift as a spectral phase:
//   S(ky,kx) = exp(+2*pi*i*(ky*floor(Hp/2)/Hp + kx*floor(Wp/2)/Wp)) = Sy[ky]*Sx[kx]
// (= (-1)^(ky+kx) for even sizes).  Mirrors fft.ifftshift at rfft_convolve.py:150,195.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_shift_table(typename Cx<T>::type* __restrict__ tab, int n_entries, int N) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_entries) return;
    long long m = ((long long)k * (N / 2)) % N;  // exact phase reduction in integers
    double s, c;
    sincospi(2.0 * (double)m / (double)N, &s, &c);
    typename Cx<T>::type v; v.x = (T)c; v.y = (T)s;
    tab[k] = v;
}

// ---------------------------------------------------------------------------
// K4c: ADMM  R = 1 / (mu1*|H|^2 + mu2*|PsiT Psi| + mu3)   (admm.py:186-190), with
//   |PsiT Psi|(ky,kx) = 4 - 2cos(2*pi*ky/Hp) - 2cos(2*pi*kx/Wp)  -- closed form of
//   |rfft2(finite_diff_gram)| (admm.py:373-397; equal to 4e-16)
// and cuFFT's unnormalised inverse folded in: Rn = R / (Hp*Wp).  Stored REAL.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_admm_R(const typename Cx<T>::type* __restrict__ H, T* __restrict__ R, int DC, int Hp, int Wp,
                         int Nh, double mu1, double mu2, double mu3) {
    const int64_t plane = (int64_t)Hp * Nh;
    const int64_t n = plane * DC;
    const double inv_n = 1.0 / ((double)Hp * (double)Wp);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i % plane;
        int ky = (int)(r / Nh), kx = (int)(r - (int64_t)ky * Nh);
        typename Cx<T>::type h = H[i];
        // |conj(H)*H| evaluated in the working precision like the reference, the rest in double
        T hh = h.x * h.x + h.y * h.y;
        double g = 4.0 - 2.0 * cospi(2.0 * (double)ky / (double)Hp) - 2.0 * cospi(2.0 * (double)kx / (double)Wp);
        double den = mu1 * (double)hh + mu2 * fabs(g) + mu3;
        R[i] = (T)(inv_n / den);
    }
}

// ---------------------------------------------------------------------------
// Per-PSF constants (device, T[4*NC]): [0,NC) hmax  [NC,2NC) psf min  [2NC,3NC) psf max  [3NC,4NC) alpha.
// NC = C when the plan holds ONE PSF (every sample shares it), NC = B*C when it holds one PSF per sample
// (LPC_FLAG_PSF_PER_SAMPLE: multi-mask batches, eval/benchmark.py:186-190,255-259).  Column of plane p:
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int lpc_const_col(int p, int DC, int C, int per_sample) {
    return (per_sample ? (p / DC) * C : 0) + p % C;
}

// K4d: GD step size (gd.py:108-112): hmax[col] = max over (dep,ky,kx) |conj(H)*H| of that PSF's channel
template <typename T>
__global__ void k_hh_max(const typename Cx<T>::type* __restrict__ H, T* __restrict__ hmax, int C, int DC, int per_sample,
                         int64_t plane) {
    // grid: y = transfer-function plane q = (s*D + dep)*C + c
    const int q = blockIdx.y;
    const int col = lpc_const_col(q, DC, C, per_sample);
    const typename Cx<T>::type* Hq = H + (int64_t)q * plane;
    T m = T(0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        typename Cx<T>::type h = Hq[i];
        m = tmax(m, h.x * h.x + h.y * h.y);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomic_max(&hmax[col], m);
}

// PSF per-channel min / max over (D,h,w) on the caller's channels-last PSF (gd.py:100-103); grid (blocks, C, n_psf)
template <typename T>
__global__ void k_psf_minmax(const T* __restrict__ psf, int64_t npix, int C, T* __restrict__ vmin, T* __restrict__ vmax) {
    const int c = blockIdx.y, s = blockIdx.z;
    psf += (int64_t)s * npix * C;
    T lo = psf[c], hi = psf[c];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        T v = psf[i * C + c];
        lo = tmin(lo, v); hi = tmax(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { atomic_min(&vmin[s * C + c], lo); atomic_max(&vmax[s * C + c], hi); }
}

// keep_x0: leave the psf min / max columns alone (UnrolledFISTA keeps the constructor PSF's initial value when a new
// PSF arrives through forward(psfs=...), trainable_recon.py:347-350 vs unrolled_fista.py:55-59)
template <typename T>
__global__ void k_init_consts(T* consts, const T* __restrict__ psf, int C, int NC, int64_t npixC, int keep_x0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NC) return;
    const int s = i / C, c = i - s * C;
    consts[i] = T(0);
    consts[3 * NC + i] = T(0);
    if (!keep_x0) { const T v = psf[(int64_t)s * npixC + c]; consts[NC + i] = v; consts[2 * NC + i] = v; }
}
template <typename T>
__global__ void k_alpha(T* consts, int NC, T lip_fact) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NC) consts[3 * NC + i] = lip_fact / consts[i];
}

// power-of-two balance of the fused engine's (V, HV) row packing (FusedK::pack): {t, 1/t} with
// t = 2^clamp(round(log2(sqrt(max |H|^2))) - 3, 0, 24); fixed_exp >= 0 overrides the exponent (LPC_PACK_T)
template <typename T>
__global__ void k_pack_scale(const T* __restrict__ consts, int NC, float* __restrict__ pack, int fixed_exp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T m = T(0);
    for (int i = 0; i < NC; ++i) m = tmax(m, consts[i]);
    int e = 0;
    if (m > T(0)) e = (int)lrint(0.5 * log2((double)m)) - 3;
    if (e < 0) e = 0;
    if (e > 24) e = 24;
    if (fixed_exp > -1000) e = fixed_exp;
    pack[0] = ldexpf(1.0f, e);
    pack[1] = ldexpf(1.0f, -e);
}

// fill the window of planar planes with a per-channel constant: GD initial estimate
// x0 = (max_c + min_c)/2  (gd.py:99-105)
template <typename T>
__global__ void k_fill_x0(T* __restrict__ dst, const T* __restrict__ consts, int C, int DC, int NC, int per_sample, int h,
                          int w, int64_t plane_stride, int pitch, int oy, int ox) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const int y = blockIdx.y;
    const int p = blockIdx.z;
    const int col = lpc_const_col(p, DC, C, per_sample);
    const T v = (consts[2 * NC + col] + consts[NC + col]) / T(2);
    dst[(int64_t)p * plane_stride + (int64_t)(y + oy) * pitch + (x + ox)] = v;
}

}  // namespace lpc
