// lpc_admm.cuh -- the per-iteration ADMM kernels (admm.py:232-329 fused).
//
// One ADMM iteration is  K1 (spatial, this file) -> R2C(a), R2C(xi') -> K2 (spectral,
// this file) -> C2R(V), C2R(HV).  U, X, W and Psi(V) are never materialised; the duals
// are carried in half-updated form:
//     xi'  = xi  - mu1*X      eta' = eta - mu2*U      rho' = rho - mu3*W
// and completed at the start of the next K1 pass with the new V / HV:
//     xi = xi' + mu1*HV       eta = eta' + mu2*Psi(V) rho = rho' + mu3*V
// which is algebraically the reference's  xi += mu1*(HV - X) etc. (admm.py:298-311).
//
// Per pixel (i,j) of every padded plane, with periodic neighbours (finite_diff uses
// roll, admm.py:349-370):
//     psi0 = V(i-1,j) - V(i,j)            psi1 = V(i,j-1) - V(i,j)          [Psi,  admm.py:349-359]
//     U_k  = soft(psi_k + eta_k/mu2, tau/mu2)                                 [admm.py:245-247, 341-346]
//     X    = Xdiv * (xi + mu1*HV + pad(y))                                    [admm.py:252-254]
//     W    = max(rho/mu3 + V, 0)                                              [admm.py:259-263]
//     t_k  = mu2*U_k - eta_k  (= -eta'_k)
//     a    = (mu3*W - rho) + (t0(i+1,j) - t0(i,j)) + (t1(i,j+1) - t1(i,j))    [admm.py:277-279, 362-370]
//     bq   = mu1*X - xi = -xi'   (not stored: K2 subtracts instead)
// so a pixel needs the 5-point star of V, eta'0 at rows i and i+1, eta'1 at columns j
// and j+1.  eta' is ping-ponged between two buffers because the +1 halo of eta' is read
// by a different thread than the one that rewrites it.
#pragma once
#include "lpc_util.cuh"

namespace lpc {

template <typename T>
struct AdmmK {
    T mu1, mu2, mu3, thr;     // thr = tau / mu2
    T mu1p, mu2p, mu3p;       // the PREVIOUS iteration's mu: it completes the half-updated duals (equal to mu1..3
                              // unless an unrolled per-iteration schedule is set, unrolled_admm.py:222-234)
    T inv_mu2, inv_mu3;       // reciprocals for the fused engine (FAST = true: x*(1/mu) instead of x/mu)
    T xdiv_in, xdiv_out;      // X_divmat inside / outside the sensor window (admm.py:193)
    int Hp, Wp;               // padded plane
    int h, w, sy, sx;         // sensor window  rows [sy,sy+h) cols [sx,sx+w)
    int C, D, Dd;             // channels, PSF depth, data depth (1 or D)
    int dcol0, dpitch;        // data planes: row pitch, and padded column of their element 0
    int64_t plane;            // Hp*Wp
    int64_t dplane;           // h*dpitch
    int n_planes;             // B*D*C
};

template <typename T> __device__ __forceinline__ T soft_thresh(T z, T thr) {
    // sign(z) * max(|z| - thr, 0)      (admm.py:341-346)
    T m = fabs(z) - thr;
    m = m > T(0) ? m : T(0);
    return copysign(m, z);
}

// one TV component: completes eta, applies the U prox, returns t = mu2*U - eta = -eta'_new
// FAST: multiply by the precomputed reciprocal instead of the IEEE division (<= 1 ulp apart)
template <typename T, bool FAST = false>
__device__ __forceinline__ T tv_component(T psi, T eta_half, const AdmmK<T>& k, T& eta_half_new) {
    const T eta = eta_half + k.mu2p * psi;
    const T u = soft_thresh(psi + (FAST ? eta * k.inv_mu2 : eta / k.mu2), k.thr);
    eta_half_new = eta - k.mu2 * u;
    return -eta_half_new;
}

template <typename T, bool FAST = false>
__device__ __forceinline__ void point_update(T v, T hv, T y, bool in_window, T xi_half, T rho_half,
                                             const AdmmK<T>& k, T& xi_half_new, T& rho_half_new) {
    const T xi = xi_half + k.mu1p * hv;
    const T x = (in_window ? k.xdiv_in : k.xdiv_out) * (xi + k.mu1 * hv + y);
    xi_half_new = xi - k.mu1 * x;
    const T rho = rho_half + k.mu3p * v;
    T w = (FAST ? rho * k.inv_mu3 : rho / k.mu3) + v;
    w = w > T(0) ? w : T(0);
    rho_half_new = rho - k.mu3 * w;
}

// image plane p -> data plane (depth-1 data is broadcast over PSF depth, gd.py:129)
template <typename T> __device__ __forceinline__ int data_plane_of(int p, const AdmmK<T>& k) {
    const int dc = k.D * k.C;
    const int b = p / dc, rem = p - b * dc;
    const int dep = rem / k.C, c = rem - dep * k.C;
    return (b * k.Dd + (k.Dd == 1 ? 0 : dep)) * k.C + c;
}

// ---------------------------------------------------------------------------
// K1 (generic): one thread per pixel, any Hp x Wp (odd sizes, Wp % 4 != 0).
// grid (ceil(Wp/128), Hp, planes)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
k_admm_spatial_generic(AdmmK<T> k, const T* __restrict__ V, const T* __restrict__ HV,
                       const T* __restrict__ data, T* __restrict__ xi, T* __restrict__ rho,
                       const T* __restrict__ eta_in, T* __restrict__ eta_out, T* __restrict__ a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.Wp) return;
    const int i = blockIdx.y;
    const int p = blockIdx.z;
    const int64_t base = (int64_t)p * k.plane;
    const int64_t eta1_off = (int64_t)k.n_planes * k.plane;  // eta[1] planes follow eta[0] planes
    const int iu = (i == 0) ? k.Hp - 1 : i - 1;
    const int id = (i == k.Hp - 1) ? 0 : i + 1;
    const int jl = (j == 0) ? k.Wp - 1 : j - 1;
    const int jr = (j == k.Wp - 1) ? 0 : j + 1;
    const int64_t o = base + (int64_t)i * k.Wp + j;
    const int64_t od = base + (int64_t)id * k.Wp + j;
    const int64_t orr = base + (int64_t)i * k.Wp + jr;

    const T v = V[o];
    const T vu = V[base + (int64_t)iu * k.Wp + j];
    const T vd = V[od];
    const T vl = V[base + (int64_t)i * k.Wp + jl];
    const T vr = V[orr];

    T e0n, e1n, dummy;
    const T t0 = tv_component(vu - v, eta_in[o], k, e0n);
    const T t1 = tv_component(vl - v, eta_in[eta1_off + o], k, e1n);
    const T t0d = tv_component(v - vd, eta_in[od], k, dummy);
    const T t1r = tv_component(v - vr, eta_in[eta1_off + orr], k, dummy);

    const bool inw = (i >= k.sy) && (i < k.sy + k.h) && (j >= k.sx) && (j < k.sx + k.w);
    T y = T(0);
    if (inw) y = data[(int64_t)data_plane_of(p, k) * k.dplane + (int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0)];
    T xin, rhon;
    point_update(v, HV[o], y, inw, xi[o], rho[o], k, xin, rhon);

    xi[o] = xin;
    rho[o] = rhon;
    eta_out[o] = e0n;
    eta_out[eta1_off + o] = e1n;
    a[o] = -rhon + ((t0d - t0) + (t1r - t1));
}

// ---------------------------------------------------------------------------
// K1 (vector): the fast path, Wp % VEC == 0 (VEC = 4 floats / 2 doubles = 16 B).
// Warp-centric: a warp owns a 32*VEC-pixel segment of RY consecutive rows; lanes hold
// 16-byte vectors; the +-1 column neighbours come from warp shuffles, the +-1 row
// neighbours from a rolling register window (each V / eta'0 row is loaded once per
// row group, plus one halo row).  All HBM accesses are 128-bit and fully coalesced
// (a warp reads/writes 512 contiguous bytes per array per row).
// grid: flat over warps (planes * row_groups * segments), 8 warps per CTA.
// ---------------------------------------------------------------------------
template <typename T, int RY>
__global__ void __launch_bounds__(256)
k_admm_spatial_vec(AdmmK<T> k, const T* __restrict__ V, const T* __restrict__ HV,
                   const T* __restrict__ data, T* __restrict__ xi, T* __restrict__ rho,
                   const T* __restrict__ eta_in, T* __restrict__ eta_out, T* __restrict__ a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int lane = threadIdx.x & 31;
    const int nvec = k.Wp / VEC;
    const int nseg = (nvec + 31) >> 5;
    const int nrg = (k.Hp + RY - 1) / RY;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp_id >= (int64_t)k.n_planes * nrg * nseg) return;  // whole warp exits together
    const int seg = (int)(warp_id % nseg);
    const int64_t t = warp_id / nseg;
    const int rg = (int)(t % nrg);
    const int p = (int)(t / nrg);

    int vx = seg * 32 + lane;
    const bool active = vx < nvec;
    if (!active) vx = nvec - 1;  // keep the lane in the shuffles, on a valid address; stores masked
    const int j = vx * VEC;
    const bool left_edge = (lane == 0) || (vx == 0);
    const bool right_edge = (lane == 31) || (vx == nvec - 1);
    const int jl = (j == 0) ? k.Wp - 1 : j - 1;
    const int jr = (j + VEC == k.Wp) ? 0 : j + VEC;

    const int64_t base = (int64_t)p * k.plane;
    const T* eta0_in = eta_in + base;
    const T* eta1_in = eta_in + (int64_t)k.n_planes * k.plane + base;
    T* eta0_out = eta_out + base;
    T* eta1_out = eta_out + (int64_t)k.n_planes * k.plane + base;
    const T* Vp = V + base;
    const T* HVp = HV + base;
    T* xip = xi + base;
    T* rhop = rho + base;
    T* ap = a + base;
    const T* dp = data + (int64_t)data_plane_of(p, k) * k.dplane;
    const bool col_has_data = (j >= k.dcol0) && (j < k.dcol0 + k.dpitch);

    const int i0 = rg * RY;
    const int nrows = (k.Hp - i0 < RY) ? (k.Hp - i0) : RY;

    // prologue: row i0 and its upper neighbour -> t0 / eta'0_new of row i0
    Vec16<T> vcur = ld16_nc(Vp + (int64_t)i0 * k.Wp + j);
    Vec16<T> t0cur, e0new_cur;
    {
        const int iu = (i0 == 0) ? k.Hp - 1 : i0 - 1;
        const Vec16<T> vup = ld16_nc(Vp + (int64_t)iu * k.Wp + j);
        const Vec16<T> e0 = ld16_nc(eta0_in + (int64_t)i0 * k.Wp + j);
#pragma unroll
        for (int e = 0; e < VEC; ++e) t0cur.v[e] = tv_component(vup.v[e] - vcur.v[e], e0.v[e], k, e0new_cur.v[e]);
    }

#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
        const int i = i0 + r;
        const int id = (i == k.Hp - 1) ? 0 : i + 1;
        const int64_t ro = (int64_t)i * k.Wp;
        const int64_t rd = (int64_t)id * k.Wp;
        // ---- all 16-byte loads of this row step, issued back to back
        const Vec16<T> vnext = ld16_nc(Vp + rd + j);
        const Vec16<T> e0next = ld16_nc(eta0_in + rd + j);
        const Vec16<T> e1 = ld16_nc(eta1_in + ro + j);
        const Vec16<T> hv = ld16_nc(HVp + ro + j);
        const Vec16<T> xih = ld16(xip + ro + j);
        const Vec16<T> rhoh = ld16(rhop + ro + j);
        const bool row_in = (i >= k.sy) && (i < k.sy + k.h);
        Vec16<T> y;
#pragma unroll
        for (int e = 0; e < VEC; ++e) y.v[e] = T(0);
        if (row_in && col_has_data) y = ld16_nc(dp + (int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0));

        // ---- column neighbours: shuffles inside the warp, scalar (wrapped) loads at its edges
        T vl = __shfl_up_sync(0xffffffffu, vcur.v[VEC - 1], 1);
        T vr = __shfl_down_sync(0xffffffffu, vcur.v[0], 1);
        T e1r = __shfl_down_sync(0xffffffffu, e1.v[0], 1);
        if (left_edge) vl = __ldg(Vp + ro + jl);
        if (right_edge) { vr = __ldg(Vp + ro + jr); e1r = __ldg(eta1_in + ro + jr); }

        // ---- TV components
        Vec16<T> t1, e1new, t0next, e0new_next;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T left = (e == 0) ? vl : vcur.v[e - 1];
            t1.v[e] = tv_component(left - vcur.v[e], e1.v[e], k, e1new.v[e]);
            t0next.v[e] = tv_component(vcur.v[e] - vnext.v[e], e0next.v[e], k, e0new_next.v[e]);
        }
        T dummy;
        const T t1r = tv_component(vcur.v[VEC - 1] - vr, e1r, k, dummy);

        // ---- pointwise primal / dual updates and the FFT operand a
        Vec16<T> xin, rhon, av;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int jj = j + e;
            const bool inw = row_in && (jj >= k.sx) && (jj < k.sx + k.w);
            point_update(vcur.v[e], hv.v[e], y.v[e], inw, xih.v[e], rhoh.v[e], k, xin.v[e], rhon.v[e]);
            const T t1right = (e == VEC - 1) ? t1r : t1.v[e + 1];
            av.v[e] = -rhon.v[e] + ((t0next.v[e] - t0cur.v[e]) + (t1right - t1.v[e]));
        }
        if (active) {
            st16(xip + ro + j, xin);
            st16(rhop + ro + j, rhon);
            st16(eta0_out + ro + j, e0new_cur);
            st16(eta1_out + ro + j, e1new);
            st16(ap + ro + j, av);
        }
        vcur = vnext;
        t0cur = t0next;
        e0new_cur = e0new_next;
    }
}

// ---------------------------------------------------------------------------
// K2: spectral step, one pass over the half spectrum (admm.py:277-294 + 320 merged):
//     Vhat  = Rn * ( A - (S*conj(H)) * Xi' )      [A = rfft2(a), Xi' = rfft2(xi') = -rfft2(bq)]
//     HVhat = (S*H) * Vhat
// written in place over A and Xi'.  Rn carries 1/(Hp*Wp).  S = Sy[ky]*Sx[kx].
// The adjoint uses S*conj(H), NOT conj(S*H): deconvolve applies the same ifftshift
// (rfft_convolve.py:195).
// spec holds [2][planes][Hp][Nh]; H, R hold [D*C][Hp][Nh]; one thread per bin.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_admm_spectral(typename Cx<T>::type* __restrict__ spec, const typename Cx<T>::type* __restrict__ H,
                const T* __restrict__ R, const typename Cx<T>::type* __restrict__ Sy,
                const typename Cx<T>::type* __restrict__ Sx, int n_planes, int DC, int Hp, int Nh) {
    using C = typename Cx<T>::type;
    const int64_t plane = (int64_t)Hp * Nh;
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // bin inside the plane
    if (r >= plane) return;
    const int p = blockIdx.y;
    const int q = p % DC;
    const int ky = (int)(r / Nh), kx = (int)(r - (int64_t)ky * Nh);
    const int64_t ia = (int64_t)p * plane + r;
    const int64_t ix = ia + (int64_t)n_planes * plane;
    const C A = spec[ia];
    const C Xi = spec[ix];
    const C Hq = H[(int64_t)q * plane + r];
    const T Rn = R[(int64_t)q * plane + r];
    const C S = cmul(Sy[ky], Sx[kx]);
    const C Hs = cmul(S, Hq);
    const C Hsa = cmul(S, cconj(Hq));
    const C m = cmul(Hsa, Xi);
    C Vh; Vh.x = Rn * (A.x - m.x); Vh.y = Rn * (A.y - m.y);
    spec[ia] = Vh;
    spec[ix] = cmul(Hs, Vh);
}

// ---------------------------------------------------------------------------
// warm start (admm.py:154-155,172-176): given V and HV = convolve(V), set the half
// duals so that the first K1 pass completes them to exactly-zero-equivalent duals:
//   xi' = -mu1*HV   rho' = -mu3*V   eta'_k = -mu2*Psi_k(V)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_admm_warm_start(AdmmK<T> k, const T* __restrict__ V, const T* __restrict__ HV,
                                  T* __restrict__ xi, T* __restrict__ rho, T* __restrict__ eta) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.Wp) return;
    const int i = blockIdx.y;
    const int p = blockIdx.z;
    const int64_t base = (int64_t)p * k.plane;
    const int iu = (i == 0) ? k.Hp - 1 : i - 1;
    const int jl = (j == 0) ? k.Wp - 1 : j - 1;
    const int64_t o = base + (int64_t)i * k.Wp + j;
    const T v = V[o];
    xi[o] = -(k.mu1 * HV[o]);
    rho[o] = -(k.mu3 * v);
    eta[o] = -(k.mu2 * (V[base + (int64_t)iu * k.Wp + j] - v));
    eta[(int64_t)k.n_planes * k.plane + o] = -(k.mu2 * (V[base + (int64_t)i * k.Wp + jl] - v));
}

}  // namespace lpc
