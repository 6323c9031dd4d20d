// lpc_gd.cuh -- kernels of the gradient-descent family (gd.py:62-241) and of the bare
// RealFFTConvolve2D operator (rfft_convolve.py:133-223).
//
// One GD / Nesterov / FISTA iteration (gd.py:128-134, 183-188, 235-241):
//     g = deconvolve(convolve(x) - y)         [4 FFTs: pad -> rfft2 -> *H -> irfft2 -> ifftshift -> crop, twice]
//     x <- update(x, alpha_c * g)
// The image estimate x lives INSIDE the window of a persistent zero-bordered padded
// buffer (xpad), so "pad" costs nothing: R2C reads xpad directly and the update kernel
// writes the new estimate straight back into the window.  Likewise the residual
// crop(Hx) - y is written into the window of a second zero-bordered buffer (dpad).
#pragma once
#include "lpc_setup.cuh"
#include "lpc_util.cuh"

namespace lpc {

struct GdK {
    int Hp, Wp, Nh;
    int h, w, sy, sx;
    int C, D, Dd;
    int dpitch;        // data planes are [h][dpitch]
    int64_t plane;     // Hp*Wp
    int64_t dplane;    // h*dpitch
    int n_planes;
    int NC, PS;        // columns of the constants table, 1 = one PSF per sample (lpc_const_col, lpc_setup.cuh)
};

__device__ __forceinline__ int gd_data_plane(int p, const GdK& k) {
    const int dc = k.D * k.C;
    const int b = p / dc, rem = p - b * dc;
    const int dep = rem / k.C, c = rem - dep * k.C;
    return (b * k.Dd + (k.Dd == 1 ? 0 : dep)) * k.C + c;
}

// ---------------------------------------------------------------------------
// spectral multiply, in place:  spec *= scale * S * (adjoint ? conj(H) : H)
// (rfft_convolve.py:146,191 `* self._H` / `* self._Hadj` and the ifftshift at :150,:195
// folded in as the phase S; scale = 1/(Hp*Wp) for cuFFT's unnormalised C2R.)
// grid (ceil(plane/256), planes)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_spec_mul(typename Cx<T>::type* __restrict__ spec, const typename Cx<T>::type* __restrict__ H,
           const typename Cx<T>::type* __restrict__ Sy, const typename Cx<T>::type* __restrict__ Sx,
           int DC, int Hp, int Nh, int adjoint, T scale) {
    using C = typename Cx<T>::type;
    const int64_t plane = (int64_t)Hp * Nh;
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= plane) return;
    const int p = blockIdx.y;
    const int q = p % DC;
    const int ky = (int)(r / Nh), kx = (int)(r - (int64_t)ky * Nh);
    C Hq = H[(int64_t)q * plane + r];
    if (adjoint) Hq.y = -Hq.y;
    C f = cmul(cmul(Sy[ky], Sx[kx]), Hq);
    f.x *= scale; f.y *= scale;
    const int64_t i = (int64_t)p * plane + r;
    spec[i] = cmul(spec[i], f);
}

// ---------------------------------------------------------------------------
// residual:  dpad[window] = crop(Hx) - y        (gd.py:129; borders of dpad stay zero)
// grid (ceil(w/128), h, planes)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
k_gd_residual(GdK k, const T* __restrict__ hx, const T* __restrict__ data, T* __restrict__ dpad) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int y = blockIdx.y;
    const int p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(y + k.sy) * k.Wp + (x + k.sx);
    dpad[o] = hx[o] - data[(int64_t)gd_data_plane(p, k) * k.dplane + (int64_t)y * k.dpitch + x];
}

// ---------------------------------------------------------------------------
// updates.  g = crop(deconvolve(...)) is read from the window of `grad`.
// consts[3C + c] = alpha_c (gd.py:108-112).
// ---------------------------------------------------------------------------
// GradientDescent._update (gd.py:132-134): x <- max(x - alpha*g, 0)
template <typename T>
__global__ void __launch_bounds__(128)
k_gd_update(GdK k, const T* __restrict__ grad, const T* __restrict__ consts, T* __restrict__ xpad) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(blockIdx.y + k.sy) * k.Wp + (x + k.sx);
    const T alpha = consts[3 * k.NC + lpc_const_col(p, k.D * k.C, k.C, k.PS)];
    T v = xpad[o] - alpha * grad[o];
    xpad[o] = v > T(0) ? v : T(0);
}

// NesterovGradientDescent._update (gd.py:183-188):
//   p_new = mu*p - alpha*g ; x <- max(x + (-mu*p_prev + (1+mu)*p_new), 0)
// mom is planar [planes][h][w]
template <typename T>
__global__ void __launch_bounds__(128)
k_nesterov_update(GdK k, const T* __restrict__ grad, const T* __restrict__ consts, T* __restrict__ xpad,
                  T* __restrict__ mom, T mu) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(blockIdx.y + k.sy) * k.Wp + (x + k.sx);
    const int64_t m = ((int64_t)p * k.h + blockIdx.y) * k.w + x;
    const T alpha = consts[3 * k.NC + lpc_const_col(p, k.D * k.C, k.C, k.PS)];
    const T p_prev = mom[m];
    const T p_new = mu * p_prev - alpha * grad[o];
    mom[m] = p_new;
    T v = xpad[o] + (-mu * p_prev + (T(1) + mu) * p_new);
    xpad[o] = v > T(0) ? v : T(0);
}

// FISTA._update (gd.py:235-241):
//   z = x - alpha*g ; xk = max(z,0) ; x <- xk + beta*(xk - xk_prev) ; xk_prev <- xk
// beta = (t_k - 1)/t_{k+1} comes from the host (t is a host double in the reference).
// `first`: right after reset() the reference's xk_prev ALIASES x (gd.py:233) and the
// step at gd.py:236 is in place, so the previous iterate of the first iteration is z
// itself (only visible for tk != 1; SURVEY 8b quirk iii) -- mirrored.
template <typename T>
__global__ void __launch_bounds__(128)
k_fista_update(GdK k, const T* __restrict__ grad, const T* __restrict__ consts, T* __restrict__ xpad,
               T* __restrict__ xk_prev, T beta, int first) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(blockIdx.y + k.sy) * k.Wp + (x + k.sx);
    const int64_t m = ((int64_t)p * k.h + blockIdx.y) * k.w + x;
    const T alpha = consts[3 * k.NC + lpc_const_col(p, k.D * k.C, k.C, k.PS)];
    const T z = xpad[o] - alpha * grad[o];
    const T xk = z > T(0) ? z : T(0);
    const T prev = first ? z : xk_prev[m];
    xk_prev[m] = xk;
    xpad[o] = xk + beta * (xk - prev);
}

}  // namespace lpc
