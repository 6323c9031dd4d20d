// lpc_ops.cuh -- the module-level helpers of lensless/recon/admm.py as stand-alone kernels on the caller's
// channels-last arrays [..., H, W, C] (callers import them: unrolled_admm.py:10):
//
//   soft_thresh(x, t)      sign(x) * max(|x| - t, 0)                                  admm.py:341-346
//   finite_diff(x)         stack(roll(x,+1,H) - x, roll(x,+1,W) - x) on a new last axis   admm.py:349-359
//   finite_diff_adj(u)     (roll(u0,-1,H) - u0) + (roll(u1,-1,W) - u1)                admm.py:362-370
//   finite_diff_gram(shape) rfft2 of the stencil [4; -1 -1 -1 -1] (depth 1) or [6; ... and -1 at depth +-1]
//                          (depth > 1), in closed form                                admm.py:373-397
//
// Inside the solvers these operators never run as separate passes (they are fused into the spatial kernels,
// lpc_admm.cuh / lpc_fused.cuh); here they are memory-bound element-wise kernels, one thread per element, coalesced.
#pragma once
#include "lpc_admm.cuh"
#include "lpc_util.cuh"

namespace lpc {

template <typename T>
__global__ void k_op_soft_thresh(const T* __restrict__ x, T* __restrict__ out, int64_t n, T thr) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = soft_thresh(x[i], thr);
}

// x: [outer, H, W, C] -> out: [outer, H, W, C, 2]
template <typename T>
__global__ void k_op_finite_diff(const T* __restrict__ x, T* __restrict__ out, int64_t outer, int H, int W, int C) {
    const int64_t n = outer * H * W * C;
    const int64_t rowlen = (int64_t)W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t wc = i % rowlen;
        const int64_t r = i / rowlen;
        const int y = (int)(r % H);
        const int xw = (int)(wc / C);
        const T v = x[i];
        const T up = x[i + (y == 0 ? (int64_t)(H - 1) * rowlen : -rowlen)];       // roll(+1) along H: element y-1
        const T left = x[i + (xw == 0 ? (int64_t)(W - 1) * C : -(int64_t)C)];     // roll(+1) along W: element x-1
        out[2 * i] = up - v;
        out[2 * i + 1] = left - v;
    }
}

// u: [outer, H, W, C, 2] -> out: [outer, H, W, C]
template <typename T>
__global__ void k_op_finite_diff_adj(const T* __restrict__ u, T* __restrict__ out, int64_t outer, int H, int W, int C) {
    const int64_t n = outer * H * W * C;
    const int64_t rowlen = (int64_t)W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t wc = i % rowlen;
        const int64_t r = i / rowlen;
        const int y = (int)(r % H);
        const int xw = (int)(wc / C);
        const int64_t dn = i + (y == H - 1 ? -(int64_t)(H - 1) * rowlen : rowlen);    // roll(-1) along H: element y+1
        const int64_t rt = i + (xw == W - 1 ? -(int64_t)(W - 1) * C : (int64_t)C);    // roll(-1) along W: element x+1
        const T d1 = u[2 * dn] - u[2 * i];
        const T d2 = u[2 * rt + 1] - u[2 * i + 1];
        out[i] = d1 + d2;
    }
}

// out: complex [D, Hp, Nh, C] = rfft2(gram) over (-3, -2).  Depth 1: 4 - 2cos(ty) - 2cos(tx).  Depth > 1 (admm.py:388-392):
// slice 0 holds the centre 6 and the four in-plane -1 taps, slices 1 and D-1 a single -1 at the origin (a delta: a
// constant -1 spectrum; D == 2 writes that slice once), every other slice is zero.  rfft2 does not transform depth.
template <typename T>
__global__ void k_op_finite_diff_gram(typename Cx<T>::type* __restrict__ out, int D, int Hp, int Wp, int Nh, int C) {
    const int64_t n = (int64_t)D * Hp * Nh * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / C;
        const int kx = (int)(r % Nh); r /= Nh;
        const int ky = (int)(r % Hp);
        const int dep = (int)(r / Hp);
        double v = 0.0;
        // taps at (0,+-1) and (+-1,0) wrap: with Hp == 1 or Wp == 1 ... sizes here are >= 2 (a 1-pixel axis would alias)
        const double inplane = -2.0 * cospi(2.0 * (double)ky / (double)Hp) - 2.0 * cospi(2.0 * (double)kx / (double)Wp);
        if (D == 1) v = 4.0 + inplane;
        else if (dep == 0) v = 6.0 + inplane;
        else if (dep == 1 || dep == D - 1) v = -1.0;
        typename Cx<T>::type c;
        c.x = (T)v; c.y = T(0);
        out[i] = c;
    }
}

}  // namespace lpc
