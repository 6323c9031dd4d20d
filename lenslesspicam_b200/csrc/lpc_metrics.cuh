// lpc_metrics.cuh -- ReconstructionAlgorithm.reconstruction_error (recon.py:607-653) after the convolution:
//
//     Hx  = crop(convolve(pad(prediction)))                         [the convolver plan, lpc_api.cu]
//     Hx  = (Hx - min) / max(Hx - min)      per (batch, depth) over (h, w, C)      (normalize=True, recon.py:640-644)
//     err = sum_{D,h,w,C} (Hx - lensless)^2 / npix                  per batch entry (recon.py:646-650)
//
// as two passes over the window of the planar C2R output: a min/max reduction and a fused normalise + squared-error
// sum, both with warp-shuffle partials and one atomic per warp.  The sum is accumulated in double (the reference's
// torch.sum is a pairwise fp32 reduction; the two agree to ~1e-6 relative).
#pragma once
#include "lpc_setup.cuh"
#include "lpc_util.cuh"

namespace lpc {

// mm[2*g] = +inf, mm[2*g+1] = -inf, acc[b] = 0
template <typename T>
__global__ void k_err_init(T* __restrict__ mm, int n_groups, double* __restrict__ acc, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) { mm[2 * i] = (T)INFINITY; mm[2 * i + 1] = -(T)INFINITY; }
    if (i < B) acc[i] = 0.0;
}

// grid (blocks, B*D): min / max over the C planes' windows of group g = b*D + dep
template <typename T>
__global__ void __launch_bounds__(256)
k_err_minmax(const T* __restrict__ src, int64_t plane_stride, int pitch, int oy, int ox, int C, int h, int w, T* __restrict__ mm) {
    const int g = blockIdx.y;
    const int64_t n = (int64_t)C * h * w;
    T lo = (T)INFINITY, hi = -(T)INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int64_t r = i / w;
        const int y = (int)(r % h), c = (int)(r / h);
        const T v = src[((int64_t)g * C + c) * plane_stride + (int64_t)(y + oy) * pitch + (x + ox)];
        lo = tmin(lo, v); hi = tmax(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { atomic_min(&mm[2 * g], lo); atomic_max(&mm[2 * g + 1], hi); }
}

// grid (blocks, B*D): acc[b] += sum over the group's window of (norm(Hx) - lensless[b, dl, y, x, c])^2
// lensless: channels-last contiguous [B, Dl, h, w, C], Dl in {1, D} (depth-1 data broadcasts, recon.py:646)
template <typename T>
__global__ void __launch_bounds__(256)
k_err_sse(const T* __restrict__ src, int64_t plane_stride, int pitch, int oy, int ox, int C, int D, int Dl, int h, int w,
          const T* __restrict__ mm, int normalize, const T* __restrict__ lensless, double* __restrict__ acc) {
    const int g = blockIdx.y;
    const int b = g / D, dep = g - b * D;
    const T* lp = lensless + ((int64_t)b * Dl + (Dl == 1 ? 0 : dep)) * h * (int64_t)w * C;
    // recon.py:641-644 in the working precision: Hx - min, then divided by max(Hx - min) = fl(max - min)
    const T lo = normalize ? mm[2 * g] : T(0);
    const T span = normalize ? mm[2 * g + 1] - lo : T(1);
    const int64_t n = (int64_t)C * h * w;
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int64_t r = i / w;
        const int y = (int)(r % h), c = (int)(r / h);
        T v = src[((int64_t)g * C + c) * plane_stride + (int64_t)(y + oy) * pitch + (x + ox)];
        if (normalize) v = (v - lo) / span;
        const T dlt = v - lp[((int64_t)y * w + x) * C + c];
        s += (double)(dlt * dlt);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[b], s);
}

template <typename T>
__global__ void k_err_final(const double* __restrict__ acc, T* __restrict__ out, int B, double inv_npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) out[i] = (T)(acc[i] * inv_npix);
}

}  // namespace lpc
