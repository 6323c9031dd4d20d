// lpc_pnp.cuh -- plug-and-play (PnP) iterations: the reference lets a denoiser replace the prox / projection
// (admm.py:126-133,235-243,266-275,302-329; gd.py:89-92,136-140).  The denoiser is a torch callable OUTSIDE the engine, so
// an iteration is split where the reference calls it:
//
//   ADMM   begin : denoiser input  = U + eta/mu2 (use_dual) | V                        [admm.py:236-243]
//          finish: U <- denoised;  X, W updates;  rk = (mu3 W - rho) + mu2 U - eta  (use_dual)
//                                                  | mu2 U + deconvolve(mu1 X - xi)   (not use_dual)
//                  -- the reference's conditional expression binds exactly like this (admm.py:266-275): with use_dual
//                  the deconvolve term is dropped, without it the (mu3 W - rho) term is -- mirrored --
//                  V = irfft2(R rfft2(rk)), HV = convolve(V), xi += mu1 (HV - X), eta += mu2 (V - U) [use_dual only],
//                  rho += mu3 (V - W)                                                 [admm.py:249-329]
//   GD     begin : z = x - alpha g   (Nesterov: p <- mu p - alpha g, z = x - mu p_prev + (1 + mu) p)   [gd.py:132-134,183-188,235-236]
//          finish: x <- denoised     (FISTA: xk = denoised, x = xk + beta (xk - xk_prev))              [gd.py:136-140,237-241]
//
// The duals are kept in their plain form here (no half-updated trick): a denoiser call of tens of milliseconds sits
// between the two halves, so these element-wise passes are not the bottleneck and follow the reference literally.
// Everything runs on the cuFFT engine (K2 spectral kernel, R2C / C2R); denoiser tensors are channels-last
// [B, D, Hp, Wp, C] (ADMM) or [B, D, h, w, C] (GD family).
#pragma once
#include "lpc_admm.cuh"
#include "lpc_gd.cuh"

namespace lpc {

// channels-last index of planar pixel (p, i, j) in a [B*D, H, W, C] array
__device__ __forceinline__ int64_t pnp_cl_index(int p, int i, int j, int C, int H, int W) {
    const int bd = p / C, c = p - bd * C;
    return (((int64_t)bd * H + i) * W + j) * C + c;
}

// grid (ceil(Wp/128), Hp, planes)
template <typename T>
__global__ void __launch_bounds__(128)
k_pnp_admm_pre(AdmmK<T> k, const T* __restrict__ V, const T* __restrict__ U, const T* __restrict__ eta, T* __restrict__ out_cl,
               int use_dual) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.Wp) return;
    const int i = blockIdx.y, p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)i * k.Wp + j;
    out_cl[pnp_cl_index(p, i, j, k.C, k.Hp, k.Wp)] = use_dual ? U[o] + eta[o] / k.mu2 : V[o];
}

template <typename T>
__global__ void __launch_bounds__(128)
k_pnp_admm_mid(AdmmK<T> k, const T* __restrict__ den_cl, const T* __restrict__ V, const T* __restrict__ HV,
               const T* __restrict__ data, const T* __restrict__ xi, const T* __restrict__ rho, const T* __restrict__ eta,
               T* __restrict__ U, T* __restrict__ X, T* __restrict__ W, T* __restrict__ a, T* __restrict__ nbq, int use_dual) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.Wp) return;
    const int i = blockIdx.y, p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)i * k.Wp + j;
    const T u = den_cl[pnp_cl_index(p, i, j, k.C, k.Hp, k.Wp)];
    const bool inw = (i >= k.sy) && (i < k.sy + k.h) && (j >= k.sx) && (j < k.sx + k.w);
    T y = T(0);
    if (inw) y = data[(int64_t)data_plane_of(p, k) * k.dplane + (int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0)];
    const T v = V[o];
    const T x = (inw ? k.xdiv_in : k.xdiv_out) * (xi[o] + k.mu1 * HV[o] + y);   // admm.py:252-254
    T w = rho[o] / k.mu3 + v;                                                     // admm.py:259-263
    w = w > T(0) ? w : T(0);
    U[o] = u; X[o] = x; W[o] = w;
    if (use_dual) {
        a[o] = ((k.mu3 * w - rho[o]) + k.mu2 * u) - eta[o];
        nbq[o] = T(0);
    } else {
        a[o] = k.mu2 * u;
        nbq[o] = xi[o] - k.mu1 * x;  // -(mu1 X - xi): the spectral kernel SUBTRACTS S conj(H) times this plane
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
k_pnp_admm_post(AdmmK<T> k, const T* __restrict__ V, const T* __restrict__ HV, const T* __restrict__ X, const T* __restrict__ W,
                const T* __restrict__ U, T* __restrict__ xi, T* __restrict__ rho, T* __restrict__ eta, int use_dual) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.Wp) return;
    const int i = blockIdx.y, p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)i * k.Wp + j;
    const T v = V[o];
    xi[o] += k.mu1 * (HV[o] - X[o]);                 // admm.py:298-300
    if (use_dual) eta[o] += k.mu2 * (v - U[o]);      // admm.py:302-306, :325-326
    rho[o] += k.mu3 * (v - W[o]);                    // admm.py:310-311
}

// ---- gradient-descent family.  grad = crop(deconvolve(convolve(x) - y)) in the window of `grad`; x in the window of xpad
// mode: LPC_ALGO_GD / NESTEROV / FISTA.  out_cl: [B*D, h, w, C].      grid (ceil(w/128), h, planes)
template <typename T>
__global__ void __launch_bounds__(128)
k_pnp_gd_step(GdK k, int algo, const T* __restrict__ grad, const T* __restrict__ consts, const T* __restrict__ xpad,
              T* __restrict__ mom, T mu, int fista_first, T* __restrict__ out_cl) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int y = blockIdx.y, p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(y + k.sy) * k.Wp + (x + k.sx);
    const int64_t m = ((int64_t)p * k.h + y) * k.w + x;
    const T alpha = consts[3 * k.NC + lpc_const_col(p, k.D * k.C, k.C, k.PS)];
    T z;
    if (algo == 3) {  // Nesterov, gd.py:183-188
        const T p_prev = mom[m];
        const T p_new = mu * p_prev - alpha * grad[o];
        mom[m] = p_new;
        z = xpad[o] + (-mu * p_prev + (T(1) + mu) * p_new);
    } else {
        z = xpad[o] - alpha * grad[o];
        if (algo == 4 && fista_first) mom[m] = z;  // the previous iterate aliases the stepped estimate (gd.py:233,236)
    }
    out_cl[pnp_cl_index(p, y, x, k.C, k.h, k.w)] = z;
}

template <typename T>
__global__ void __launch_bounds__(128)
k_pnp_gd_finish(GdK k, int algo, const T* __restrict__ den_cl, T* __restrict__ xpad, T* __restrict__ mom, T beta) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= k.w) return;
    const int y = blockIdx.y, p = blockIdx.z;
    const int64_t o = (int64_t)p * k.plane + (int64_t)(y + k.sy) * k.Wp + (x + k.sx);
    const T d = den_cl[pnp_cl_index(p, y, x, k.C, k.h, k.w)];
    if (algo == 4) {  // FISTA, gd.py:237-241
        const int64_t m = ((int64_t)p * k.h + y) * k.w + x;
        const T prev = mom[m];
        mom[m] = d;
        xpad[o] = d + beta * (d - prev);
    } else {
        xpad[o] = d;
    }
}

}  // namespace lpc
