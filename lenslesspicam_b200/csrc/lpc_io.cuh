// lpc_io.cuh -- the input side of the path on the GPU: what lensless/utils/io.py load_psf (:210-388) /
// load_image (:20-207) / load_data (:389-600) do to a decoded sensor frame, as kernels on the raw uint8 / uint16 array
// resident in device memory (the file decode itself -- PNG inflate -- stays on the host):
//
//   BGR -> RGB (cv2.cvtColor, io.py:152), flips (:157-162), cast to float32
//   PSF:  per-channel background = mean of the [bg_pix0, bg_pix1)^2 corner (:343-355), subtract, clip at 0 (:357)
//   data: subtract bg * max_val (:164-171; max_val = get_max_val(raw), image.py:251-280), clip at 0
//   resize: torchvision Resize(antialias=True) = separable triangle filter with support `scale` (image.py:28-67,
//           ATen UpSampleKernel.cpp _compute_indices_min_size_weights_aa), width pass then height pass, float32
//   normalise: PSF by its L2 norm (:375), data by its max (:196-197)
//   rgb2gray (image.py:158-206, numpy branch: weights 0.299 / 0.587 / 0.114 in double)
//
// Arrays are channels-last [D, H, W, C] as the reference holds them.  Reductions use warp-shuffle partials and one atomic
// per warp (double accumulators for sums).
#pragma once
#include "lpc_util.cuh"

namespace lpc {

// raw [H, W, C] (uint8 / uint16) -> float [H, W, C]: channel order reversed when `bgr` (C == 3), flips applied
template <typename RAW>
__global__ void k_io_raw_to_float(const RAW* __restrict__ raw, float* __restrict__ out, int H, int W, int C, int bgr,
                                  int flip_ud, int flip_lr, unsigned int* __restrict__ raw_max) {
    const int64_t n = (int64_t)H * W * C;
    unsigned int mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const int x = (int)(r % W), y = (int)(r / W);
        const int sy = flip_ud ? H - 1 - y : y, sx = flip_lr ? W - 1 - x : x;
        const int sc = (bgr && C == 3) ? 2 - c : c;
        const RAW v = raw[((int64_t)sy * W + sx) * C + sc];
        out[i] = (float)v;
        mx = max(mx, (unsigned int)v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && raw_max) atomicMax(raw_max, mx);
}

// get_max_val (image.py:251-280): nbits = ceil(log2(max)), rounded up to a supported depth {8,10,12,16}; 2^nbits - 1
__global__ void k_io_max_val(const unsigned int* __restrict__ raw_max, float* __restrict__ max_val) {
    if (threadIdx.x || blockIdx.x) return;
    const unsigned int m = *raw_max;
    int nbits = (m <= 1) ? 0 : (int)ceil(log2((double)m));
    const int depths[4] = {8, 10, 12, 16};
    bool ok = false;
    for (int i = 0; i < 4; ++i) ok |= (nbits == depths[i]);
    if (!ok) { int pick = 16; for (int i = 3; i >= 0; --i) if (nbits < depths[i]) pick = depths[i]; nbits = pick; }
    double mv = exp2((double)nbits) - 1.0;
    if ((double)m > mv) mv = exp2(ceil(log2((double)m))) - 1.0;
    *max_val = (float)mv;
}

// bg[c] = mean over rows [p0,p1) x cols [p0,p1) of all depths (io.py:343-355), float32 result
__global__ void k_io_region_mean(const float* __restrict__ img, int D, int H, int W, int C, int p0, int p1, float* __restrict__ bg) {
    const int c = blockIdx.x;
    const int side = p1 - p0;
    const int64_t n = (int64_t)D * side * side;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int x = (int)(i % side);
        const int64_t r = i / side;
        const int y = (int)(r % side), d = (int)(r / side);
        s += (double)img[(((int64_t)d * H + (p0 + y)) * W + (p0 + x)) * C + c];
    }
    __shared__ double part[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += part[w];
        bg[c] = (float)(t / (double)n);
    }
}

// img[.., c] = max(img[.., c] - bg[c] * (*scale or 1), 0)      (io.py:164-171 / :346-357)
__global__ void k_io_sub_clip(float* __restrict__ img, int64_t n, int C, const float* __restrict__ bg, const float* __restrict__ scale) {
    const float sc = scale ? *scale : 1.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = img[i] - bg[i % C] * sc;
        img[i] = v > 0.0f ? v : 0.0f;
    }
}
__global__ void k_io_scale_vec(float* __restrict__ v, int n, const float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i] / *denom;
}

// one separable antialiased pass along an axis of length n_in -> n_out with element stride `stride`;
// the array is [outer][n_axis][inner] with inner = stride.  idx[o] = {xmin, xsize}, wts[o * max_w + j].
__global__ void k_io_resize_axis(const float* __restrict__ in, float* __restrict__ out, int64_t outer, int n_in, int n_out,
                                 int64_t inner, const int2* __restrict__ idx, const float* __restrict__ wts, int max_w) {
    const int64_t n = outer * n_out * inner;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t in_i = i % inner;
        const int64_t r = i / inner;
        const int o = (int)(r % n_out);
        const int64_t ou = r / n_out;
        const int2 ix = idx[o];
        const float* w = wts + (int64_t)o * max_w;
        const float* src = in + (ou * n_in + ix.x) * inner + in_i;
        float t = src[0] * w[0];  // accumulation order and precision of ATen's basic_loop_aa_* (float, ascending j)
        for (int j = 1; j < ix.y; ++j) t += src[(int64_t)j * inner] * w[j];
        out[i] = t;
    }
}

// acc[0] += sum x^2 (double), acc_max = max x  (non-negative data: uint bit pattern order == float order)
__global__ void k_io_reduce(const float* __restrict__ x, int64_t n, double* __restrict__ sumsq, unsigned int* __restrict__ vmax_bits) {
    double s = 0.0;
    float m = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        s += (double)v * (double)v;
        m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sumsq, s); atomicMax(vmax_bits, __float_as_uint(m)); }
}
// mode 0: x /= sqrt(sumsq) as float32 (np.linalg.norm of a float32 array returns float32, io.py:375); mode 1: x /= max
__global__ void k_io_normalize(float* __restrict__ x, int64_t n, const double* __restrict__ sumsq, const unsigned int* __restrict__ vmax_bits, int mode) {
    const float d = mode == 0 ? (float)sqrt(*sumsq) : __uint_as_float(*vmax_bits);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = x[i] / d;
}

// rgb2gray, numpy branch (image.py:197-206): tensordot with [0.299, 0.587, 0.114] in double, cast to float32
__global__ void k_io_rgb2gray(const float* __restrict__ rgb, float* __restrict__ gray, int64_t npix) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)rgb[3 * i] * 0.299 + (double)rgb[3 * i + 1] * 0.587 + (double)rgb[3 * i + 2] * 0.114;
        gray[i] = (float)v;
    }
}
// single_psf (io.py:363-368): sum the channels
__global__ void k_io_sum_channels(const float* __restrict__ in, float* __restrict__ out, int64_t npix, int C) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        float s = in[i * C];
        for (int c = 1; c < C; ++c) s += in[i * C + c];
        out[i] = s;
    }
}

}  // namespace lpc
