// lpc_fused_tma.cuh -- kernel A of the fused ADMM engine as a PERSISTENT, TMA-pipelined kernel.
//
//   k_admm_cols_tma: column FFT -> K2 spectral update -> column IFFT, one tile (LPC_NCT columns x 2 slots,
//   a contiguous Hp*32-byte block of the T-layout) at a time per compute group.
//
// Why: the non-persistent column kernel (k_admm_cols, lpc_fused.cuh) loads its tile, computes, stores, all
// in sequence, two CTAs per SM; ncu showed it latency/barrier-bound at 2.5 TB/s with the SM idle while the
// tile is in flight and HBM idle while the CTAs compute (profiles/r01_ncu_full_v2_final_raw.csv).  Here
//   * one CTA per SM stays resident and walks the tiles  item = blockIdx.x + n*gridDim.x;
//   * a producer warp moves whole tiles with 1-D bulk TMA (cp.async.bulk, SASS UBLKCP): global -> shared
//     completing on an mbarrier (complete_tx), shared -> global as a bulk group; three tile buffers rotate,
//     so the load of entry n+3 and the store of entry n overlap the arithmetic of entries n+1 / n+2;
//   * NG compute groups (named barriers) each own the tile of every NG-th entry, so one group's barrier
//     bubbles are filled by the other group's butterflies;
//   * the transforms are in-place decimation passes (lpc_fft.cuh: DIF forward, DIT inverse): one barrier
//     per pass, nothing staged across it, and the last forward pass + spectral update + first inverse pass
//     run in registers on the same addresses.  The spectral update therefore sees the column spectrum in
//     digit-reversed order: position j*RL + m holds ky = ky0(j) + m*(Hp/RL).  H is stored in that order
//     once per PSF (k_tile_H_mid, [m][j][column] so a warp reads consecutive entries) and is the only
//     per-element global read: it is issued BEFORE the barrier that precedes the stage;
//   * R = 1/(mu1|H|^2 + mu2|PsiTPsi| + mu3) is evaluated on the fly (it was 17 KB of HBM reads per tile):
//     |PsiTPsi| = (2 - 2cos(2 pi ky/Hp)) + (2 - 2cos(2 pi kx/Wp)) and cos(theta_j + 2 pi m/RL) splits into a
//     per-butterfly (cos, sin) pair from a small shared-memory table and RL compile-time-indexed constants;
//     the ifftshift phase Sy[ky] = Sy[ky0(j)] * Sy[m*Hp/RL] splits the same way.  mu1..mu3 are kernel
//     arguments, so an unrolled solver may change them every iteration (unrolled_admm.py:181-234).
//
// Reference semantics: admm.py:280-294,320 merged as in DESIGN.md section 3 (identical to k_admm_cols).
#pragma once
#include "lpc_fused.cuh"
#include "lpc_fused_gd.cuh"

namespace lpc {

// ---- PTX wrappers (sm_90+/sm_100a): mbarrier, bulk async copies, proxy fences
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared, completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
// the same two with an L2 eviction-priority policy (lpc_util.cuh)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src_smem, uint32_t bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem), "r"(bytes),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef LPC_GD_WIN
#define LPC_GD_WIN 1
#endif
constexpr int LPC_TMA_NBUF = 3;
constexpr int LPC_TMA_MAXR = 16;

// per-launch constants of the spectral update
struct ColsTmaK {
    float mu1, mu2, mu3;        // admm.py:186-190
    float scale;                // 0.5 / (Hp*Wp): unnormalised inverse transforms and the row separation's 1/2
    float cm[LPC_TMA_MAXR];     // cos(2 pi m / RL)
    float sm[LPC_TMA_MAXR];     // sin(2 pi m / RL)
    float2 sym[LPC_TMA_MAXR];   // Sy[m * Hp/RL]: ifftshift phase of the innermost digit
    int s_folded;               // 1: Hm holds S*H (both padded sizes even: the ifftshift phase S is +-1, real), so
                                // S*conj(H) = conj(S*H) and the stage needs neither S nor the two products with it
};

// shared memory: LPC_TMA_NBUF tile buffers | full[NBUF], done[NBUF] mbarriers (64 bytes) | twiddles of the in-place
// passes [tw_len] | Wj[Hp/RL] = (cos, sin)(2 pi ky0(j)/Hp) | Sy0[Hp/RL] = Sy[ky0(j)].
// (Next to 207 KB of tile buffers the L1 carve-out is gone -- ncu: 4 % L1 hits -- so every table that is
// reused across tiles lives in shared memory.)
inline size_t cols_tma_smem_bytes(int Hp, int RL, int tw_len) {
    return (size_t)LPC_TMA_NBUF * Hp * 2 * LPC_NCT * 8 + 64 + (size_t)tw_len * 8 + (size_t)2 * (Hp / RL) * 8;
}

// Hm: [DC][NT][RL][Hp/RL][NCT] (middle-stage order).  tabs: tw[tw_len] | Wj[Hp/RL] | Sy0[Hp/RL] (global copy of
// the shared-memory tables).  Gx[NT*NCT] = 2 - 2cos(2 pi kx/Wp), Sx[Nh].  Items: n_items = NT * planes of this
// launch, item -> plane f.plane0 + item / NT, tile item % NT.   grid = min(n_items, #SMs), block = NG*NTHG + 32
template <typename T, typename SEQ, int NG, int NTHG>
__global__ void __launch_bounds__(NG * NTHG + 32, 1)
k_admm_cols_tma(FusedK f, ColsTmaK ck, int n_items, const typename CxT<T>::type* __restrict__ tabs,
                const typename CxT<T>::type* Tsrc, typename CxT<T>::type* Tdst,
                const typename CxT<T>::type* __restrict__ Hm, const T* __restrict__ Gx,
                const typename CxT<T>::type* __restrict__ Sx) {
    using C = typename CxT<T>::type;
    using IP = IpTraits<SEQ>;
    static_assert(sizeof(T) == 4, "float only");
    constexpr int NCT = LPC_NCT, NB = 2 * NCT, NBUF = LPC_TMA_NBUF;
    constexpr int Hp = SEQ::N, RL = IP::last, NBL = IP::nb_last;
    static_assert(RL <= LPC_TMA_MAXR, "innermost radix too large");
    constexpr int n = Hp * NB;                               // complex elements per tile
    constexpr uint32_t TILE_BYTES = (uint32_t)n * sizeof(C);  // Hp * 32: a multiple of 16 as bulk copies need
    constexpr int NTAB = IP::tw_len + 2 * NBL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* bufs = reinterpret_cast<C*>(smem_raw);
    const uint32_t bars = smem_u32(smem_raw + (size_t)NBUF * TILE_BYTES);
    C* tw_s = reinterpret_cast<C*>(smem_raw + (size_t)NBUF * TILE_BYTES + 64);
    const C* Wj = tw_s + IP::tw_len;
    const C* Sy0 = Wj + NBL;
    auto full_bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    auto done_bar = [&](int b) { return bars + 8u * (uint32_t)(NBUF + b); };

    const int tid = threadIdx.x;
    const int NT = f.NT;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) { mbar_init(full_bar(b), 1); mbar_init(done_bar(b), NTHG); }
        mbar_fence_init();
    }
    for (int i = tid; i < NTAB; i += NG * NTHG + 32) tw_s[i] = __ldg(tabs + i);
    __syncthreads();
    // entries of this CTA: n = 0 .. count-1  <->  item blockIdx.x + n*gridDim.x; entry n uses buffer n % NBUF
    // for the (n / NBUF)-th time (mbarrier parity (n / NBUF) & 1) and is computed by group n % NG
    const int count = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid >= NG * NTHG) {
        // ------------------------------------------------------------------ producer warp (one lane)
        if (tid != NG * NTHG) return;
        auto issue_load = [&](int e) {
            const int lin = (int)blockIdx.x + e * (int)gridDim.x, item = f.rev ? n_items - 1 - lin : lin;
            const int p = f.plane0 + item / NT, tile = item % NT, q = p % f.DC;
            const int b = e % NBUF;
            mbar_arrive_expect_tx(full_bar(b), TILE_BYTES);
            bulk_g2s_hint(smem_u32(bufs + (size_t)b * n), Tsrc + ((int64_t)p * NT + tile) * n, TILE_BYTES, full_bar(b),
                          LPC_L2_EVICT_FIRST);  // last use of Tout
            // H of the tile is read by the compute group about a tile time later: pull it into L2 now
            bulk_prefetch_l2(Hm + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT, (uint32_t)(Hp * NCT * sizeof(C)));
        };
        for (int e = 0; e < NBUF && e < count; ++e) issue_load(e);
        for (int e = 0; e < count; ++e) {
            const int b = e % NBUF;
            mbar_wait(done_bar(b), (uint32_t)((e / NBUF) & 1));
            const int lin = (int)blockIdx.x + e * (int)gridDim.x, item = f.rev ? n_items - 1 - lin : lin;
            const int p = f.plane0 + item / NT, tile = item % NT;
            bulk_s2g_hint(Tdst + ((int64_t)p * NT + tile) * n, smem_u32(bufs + (size_t)b * n), TILE_BYTES, LPC_L2_EVICT_LAST);  // B1 reads it next
            bulk_commit();
            if (e + NBUF < count) {
                bulk_wait_read0();  // the store has read the buffer: it may be overwritten
                issue_load(e + NBUF);
            }
        }
        bulk_wait0();  // shared memory must outlive the last stores
        return;
    }

    // ---------------------------------------------------------------------- compute groups
    const int grp = tid / NTHG, gtid = tid - grp * NTHG;
    const BarNamed<NTHG> bar{1 + grp};
    constexpr int NPAIR = NBL * NCT;  // (butterfly, column) pairs of the middle stage
    for (int e = grp; e < count; e += NG) {
        const int b = e % NBUF;
        C* sm = bufs + (size_t)b * n;
        const int lin = (int)blockIdx.x + e * (int)gridDim.x, item = f.rev ? n_items - 1 - lin : lin;
        const int p = f.plane0 + item / NT, tile = item % NT, q = p % f.DC;
        const C* Hq = Hm + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT;
        mbar_wait(full_bar(b), (uint32_t)((e / NBUF) & 1));
        // forward column FFT of (2A_row, 2Xi_row): all DIF passes but the innermost
        IP::template dif_butlast<T, -1, NB, NTHG>(sm, tw_s, gtid, bar);
        // ---- innermost forward pass + spectral update + innermost inverse pass, in registers, in place.
        // Pair g = (column c, butterfly j): positions j*RL + m of transforms 2c (A) and 2c+1 (Xi), adjacent in
        // shared memory -> one 16-byte access per m.  H of the first pair is requested before the barrier.
        int g = gtid;
        C Hh[RL];
        C sxv; sxv.x = sxv.y = T(0);
        T gxv = T(0);
        if (g < NPAIR) {
            const int c = g % NCT, j = g / NCT;
#pragma unroll
            for (int m = 0; m < RL; ++m) Hh[m] = ld8_nc_hint(Hq + (m * NBL + j) * NCT + c, LPC_L2_EF_);  // streamed once per iteration (A 133.3 -> 131.0 us)
            const int kx = tile * NCT + c, kxc = kx < f.Nh ? kx : 0;
            sxv = __ldg(Sx + kxc);
            gxv = __ldg(Gx + kxc);
        }
        bar.sync();
#pragma unroll 1
        while (g < NPAIR) {
            const int c = g % NCT, j = g / NCT;
            float4* pp = reinterpret_cast<float4*>(sm + (size_t)(j * RL) * NB + 2 * c);
            C v0[RL], v1[RL];
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                const float4 t = pp[m * (NB / 2)];
                v0[m].x = t.x; v0[m].y = t.y; v1[m].x = t.z; v1[m].y = t.w;
            }
            Dft<T, RL, -1>::run(v0);
            Dft<T, RL, -1>::run(v1);
            const C wj = Wj[j];
            const T k0 = fmaf(ck.mu2, gxv, ck.mu3);
            if (ck.s_folded) {
#pragma unroll
                for (int m = 0; m < RL; ++m) {
                    const C Hs = Hh[m];  // S*H; |S*H|^2 = |H|^2
                    const T gy = T(2) - T(2) * (wj.x * ck.cm[m] - wj.y * ck.sm[m]);
                    const T den = fmaf(ck.mu1, fmaf(Hs.x, Hs.x, Hs.y * Hs.y), fmaf(ck.mu2, gy, k0));
                    const T Rn = __fdividef(ck.scale, den);
                    C Hsa = Hs; Hsa.y = -Hsa.y;  // S*conj(H) = conj(S*H) for real S
                    const C mm = cx_mul(Hsa, v1[m]);
                    C Vh; Vh.x = Rn * (v0[m].x - mm.x); Vh.y = Rn * (v0[m].y - mm.y);
                    v0[m] = Vh;
                    v1[m] = cx_mul(Hs, Vh);
                }
            } else {
                const C s0 = cx_mul(Sy0[j], sxv);
#pragma unroll
                for (int m = 0; m < RL; ++m) {
                    const C Hv = Hh[m];
                    // R_divmat (admm.py:186-190) with |PsiTPsi| in closed form, times scale
                    const T gy = T(2) - T(2) * (wj.x * ck.cm[m] - wj.y * ck.sm[m]);
                    const T den = fmaf(ck.mu1, fmaf(Hv.x, Hv.x, Hv.y * Hv.y), fmaf(ck.mu2, gy, k0));
                    const T Rn = __fdividef(ck.scale, den);  // MUFU.RCP: 1 ulp is far inside the parity budget
                    const C S = cx_mul(s0, ck.sym[m]);
                    const C Hs = cx_mul(S, Hv);
                    C Hc = Hv; Hc.y = -Hc.y;
                    const C Hsa = cx_mul(S, Hc);
                    const C mm = cx_mul(Hsa, v1[m]);
                    C Vh; Vh.x = Rn * (v0[m].x - mm.x); Vh.y = Rn * (v0[m].y - mm.y);
                    v0[m] = Vh;
                    v1[m] = cx_mul(Hs, Vh);
                }
            }
            // H of this thread's next pair: in flight during the inverse butterflies
            const int gn = g + NTHG;
            if (gn < NPAIR) {
                const int cn = gn % NCT, jn = gn / NCT;
#pragma unroll
                for (int m = 0; m < RL; ++m) Hh[m] = ld8_nc_hint(Hq + (m * NBL + jn) * NCT + cn, LPC_L2_EF_);
                if (NTHG % NCT != 0) {  // the column changes only when the stride is odd
                    const int kx = tile * NCT + cn, kxc = kx < f.Nh ? kx : 0;
                    sxv = __ldg(Sx + kxc);
                    gxv = __ldg(Gx + kxc);
                }
            }
            Dft<T, RL, +1>::run(v0);
            Dft<T, RL, +1>::run(v1);
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                float4 t; t.x = v0[m].x; t.y = v0[m].y; t.z = v1[m].x; t.w = v1[m].y;
                pp[m * (NB / 2)] = t;
            }
            g = gn;
        }
        // inverse column FFT: the remaining DIT passes (a barrier precedes each)
        IP::template dit_butinner<T, +1, NB, NTHG>(sm, tw_s, gtid, bar);
        // hand the tile to the producer: writes -> async proxy, then arrive (release)
        fence_proxy_async();
        mbar_arrive(done_bar(b));
    }
}

// H (planar [DC][Hp][Nh]) -> the middle-stage order of k_admm_cols_tma.
// kymap[pos] = frequency held at position pos after the forward passes (fft_ip_freq_of_pos), pos = j*RL + m.
template <typename T>
__global__ void k_tile_H_mid(const typename CxT<T>::type* __restrict__ H, const int* __restrict__ kymap,
                             typename CxT<T>::type* __restrict__ Hm, int DC, int Hp, int Nh, int NT, int RL,
                             const typename CxT<T>::type* __restrict__ Sy = nullptr, const typename CxT<T>::type* __restrict__ Sx = nullptr) {
    using C = typename CxT<T>::type;
    const int NBL = Hp / RL;
    const int64_t n = (int64_t)DC * NT * Hp * LPC_NCT;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % LPC_NCT);
        int64_t r = i / LPC_NCT;
        const int mj = (int)(r % Hp); r /= Hp;  // m*NBL + j
        const int tile = (int)(r % NT);
        const int q = (int)(r / NT);
        const int m = mj / NBL, j = mj - m * NBL;
        const int ky = kymap[j * RL + m];
        const int kx = tile * LPC_NCT + c;
        C h; h.x = h.y = T(0);
        if (kx < Nh) {
            h = H[((int64_t)q * Hp + ky) * Nh + kx];
            if (Sy != nullptr) h = cx_mul(cx_mul(Sy[ky], Sx[kx]), h);  // fold the (real, +-1) ifftshift phase: ColsTmaK::s_folded
        }
        Hm[i] = h;
    }
}

}  // namespace lpc

// =============================================================================================
// GD / Nesterov / FISTA: the column kernel of lpc_fused_gd.cuh (column FFT -> filter -> column IFFT on a
// G-layout tile of LPC_GNC columns) on the same persistent TMA pipeline.  Only the h window rows of a tile
// exist in HBM (a contiguous h*32-byte block): the producer bulk-copies them to rows sy..sy+h-1 of a tile
// buffer, the compute group zero-fills the other Hp-h rows (the zero padding, rfft_convolve.py:84-100),
// transforms in place, multiplies by the filter in registers between the innermost passes, and the
// producer bulk-stores the window rows back.  Fm: filter in middle-stage order [DC][NTg][RL][Hp/RL][4].
// =============================================================================================
namespace lpc {

template <typename T, typename SEQ, int NG, int NTHG>
__global__ void __launch_bounds__(NG * NTHG + 32, 1)
k_gd_cols_tma(GdFusedK f, int n_items, const typename CxT<T>::type* __restrict__ tabs, typename CxT<T>::type* Tg,
              const typename CxT<T>::type* __restrict__ Fm) {
    using C = typename CxT<T>::type;
    using IP = IpTraits<SEQ>;
    static_assert(sizeof(T) == 4 && LPC_GNC == 4, "float only, 32-byte chunks");
    constexpr int NC = LPC_GNC, NBUF = LPC_TMA_NBUF;
    constexpr int Hp = SEQ::N, RL = IP::last, NBL = IP::nb_last;
    constexpr int n = Hp * NC;
    constexpr uint32_t TILE_BYTES = (uint32_t)n * sizeof(C);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* bufs = reinterpret_cast<C*>(smem_raw);
    const uint32_t bars = smem_u32(smem_raw + (size_t)NBUF * TILE_BYTES);
    C* tw_s = reinterpret_cast<C*>(smem_raw + (size_t)NBUF * TILE_BYTES + 64);
    auto full_bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    auto done_bar = [&](int b) { return bars + 8u * (uint32_t)(NBUF + b); };

    const int tid = threadIdx.x;
    const int NTg = f.NTg, h = f.h, sy = f.sy;
    const uint32_t win_bytes = (uint32_t)h * NC * (uint32_t)sizeof(C);
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) { mbar_init(full_bar(b), 1); mbar_init(done_bar(b), NTHG); }
        mbar_fence_init();
    }
    for (int i = tid; i < IP::tw_len; i += NG * NTHG + 32) tw_s[i] = __ldg(tabs + i);
    __syncthreads();
    const int count = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid >= NG * NTHG) {
        if (tid != NG * NTHG) return;
        auto issue_load = [&](int e) {
            const int item = (int)blockIdx.x + e * (int)gridDim.x;
            const int p = f.plane0 + item / NTg, tile = item % NTg, q = p % f.DC;
            const int b = e % NBUF;
            mbar_arrive_expect_tx(full_bar(b), win_bytes);
            bulk_g2s(smem_u32(bufs + (size_t)b * n + (size_t)sy * NC), Tg + ((int64_t)p * NTg + tile) * (int64_t)h * NC, win_bytes,
                     full_bar(b));
            bulk_prefetch_l2(Fm + ((int64_t)q * NTg + tile) * (int64_t)n, TILE_BYTES);
        };
        for (int e = 0; e < NBUF && e < count; ++e) issue_load(e);
        for (int e = 0; e < count; ++e) {
            const int b = e % NBUF;
            mbar_wait(done_bar(b), (uint32_t)((e / NBUF) & 1));
            const int item = (int)blockIdx.x + e * (int)gridDim.x;
            const int p = f.plane0 + item / NTg, tile = item % NTg;
            bulk_s2g(Tg + ((int64_t)p * NTg + tile) * (int64_t)h * NC, smem_u32(bufs + (size_t)b * n + (size_t)sy * NC), win_bytes);
            bulk_commit();
            if (e + NBUF < count) {
                bulk_wait_read0();
                issue_load(e + NBUF);
            }
        }
        bulk_wait0();
        return;
    }

    const int grp = tid / NTHG, gtid = tid - grp * NTHG;
    const BarNamed<NTHG> bar{1 + grp};
    constexpr int NPAIR = NBL * NC;  // (butterfly, column) pairs of the middle stage
#if !LPC_GD_WIN
    const int top16 = sy * (NC / 2);                  // 16-byte vectors above the window
    const int bot0 = (sy + h) * (NC / 2);             // first 16-byte vector below it
#endif
    for (int e = grp; e < count; e += NG) {
        const int b = e % NBUF;
        C* sm = bufs + (size_t)b * n;
        const int item = (int)blockIdx.x + e * (int)gridDim.x;
        const int p = f.plane0 + item / NTg, tile = item % NTg, q = p % f.DC;
        const C* Fq = Fm + ((int64_t)q * NTg + tile) * (int64_t)n;
        mbar_wait(full_bar(b), (uint32_t)((e / NBUF) & 1));
#if LPC_GD_WIN
        // Zero padding around the window: the bulk copy only writes the window rows; the others hold the output of the
        // buffer's previous tile.  They are never read: the outermost forward pass takes every element outside the
        // window as zero (and then defines the whole buffer, in place), the outermost inverse pass stores the window rows
        // only -- no zero-fill sweep, no barrier for it, half the shared-memory loads / stores of those two passes.
        IP::template dif_butlast<T, -1, NC, NTHG, BarNamed<NTHG>, true>(sm, tw_s, gtid, bar, sy, sy + h);
#else
        // zero padding around the window: the bulk copy only writes the window rows, the others hold the output of
        // the buffer's previous tile.  Only now is the buffer ours -- its previous tile may belong to the OTHER
        // group, and the load completing is what proves that tile was computed and stored.
        {
            float4* z = reinterpret_cast<float4*>(sm);
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = gtid; i < top16; i += NTHG) z[i] = zero;
            for (int i = bot0 + gtid; i < n / 2; i += NTHG) z[i] = zero;
        }
        bar.sync();  // the zero rows of every thread are in place
        IP::template dif_butlast<T, -1, NC, NTHG>(sm, tw_s, gtid, bar);
#endif
        int g = gtid;
        C Fv[RL];
        if (g < NPAIR) {
            const int c = g % NC, j = g / NC;
#pragma unroll
            for (int m = 0; m < RL; ++m) Fv[m] = (LPC_GD_HINTS & 1) ? ld8_nc_hint(Fq + (m * NBL + j) * NC + c, LPC_L2_EF_) : __ldg(Fq + (m * NBL + j) * NC + c);
        }
        bar.sync();
#pragma unroll 1
        while (g < NPAIR) {
            const int c = g % NC, j = g / NC;
            C* pp = sm + (size_t)(j * RL) * NC + c;
            C v[RL];
#pragma unroll
            for (int m = 0; m < RL; ++m) v[m] = pp[m * NC];
            Dft<T, RL, -1>::run(v);
#pragma unroll
            for (int m = 0; m < RL; ++m) v[m] = cx_mul(v[m], Fv[m]);
            const int gn = g + NTHG;
            if (gn < NPAIR) {
                const int cn = gn % NC, jn = gn / NC;
#pragma unroll
                for (int m = 0; m < RL; ++m) Fv[m] = (LPC_GD_HINTS & 1) ? ld8_nc_hint(Fq + (m * NBL + jn) * NC + cn, LPC_L2_EF_) : __ldg(Fq + (m * NBL + jn) * NC + cn);
            }
            Dft<T, RL, +1>::run(v);
#pragma unroll
            for (int m = 0; m < RL; ++m) pp[m * NC] = v[m];
            g = gn;
        }
#if LPC_GD_WIN
        IP::template dit_butinner<T, +1, NC, NTHG, BarNamed<NTHG>, true>(sm, tw_s, gtid, bar, sy, sy + h);
#else
        IP::template dit_butinner<T, +1, NC, NTHG>(sm, tw_s, gtid, bar);
#endif
        fence_proxy_async();
        mbar_arrive(done_bar(b));
    }
}

// filter tables in middle-stage order: Fm[adj][q][tile][m][j][c] = S(ky,kx) * (adj ? conj(H) : H) / (2*Hp*Wp),
// ky = kymap[j*RL + m]
template <typename T>
__global__ void k_gd_tile_filter_mid(const typename CxT<T>::type* __restrict__ H, const typename CxT<T>::type* __restrict__ Sy,
                                     const typename CxT<T>::type* __restrict__ Sx, const int* __restrict__ kymap,
                                     typename CxT<T>::type* __restrict__ F, int DC, int Hp, int Wp, int Nh, int NTg, int RL) {
    using C = typename CxT<T>::type;
    const int NBL = Hp / RL;
    const int64_t n = (int64_t)DC * NTg * Hp * LPC_GNC;
    const T scale = (T)(0.5 / ((double)Hp * (double)Wp));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % LPC_GNC);
        int64_t r = i / LPC_GNC;
        const int mj = (int)(r % Hp); r /= Hp;
        const int tile = (int)(r % NTg);
        const int q = (int)(r / NTg);
        const int m = mj / NBL, j = mj - m * NBL;
        const int ky = kymap[j * RL + m];
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
