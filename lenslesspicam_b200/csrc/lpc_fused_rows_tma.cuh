// lpc_fused_rows_tma.cuh -- kernel B1 (row IFFT -> V, HV) with its gather done by the TMA unit.
//
// k_rows_ifft (lpc_fused.cuh) fetches the 961 scattered 32-byte chunks of a row of the T-layout with per-lane 16-byte
// loads: ncu (profiles/r01s2_ncu_full_admm_raw.csv) showed the kernel bound by the LSU data pipe (79 % busy), about a
// third of it the gather -- one L1 wavefront per 128-byte line touched, 16 lines per warp instruction.  Here the row
// is described ONCE as a 3-D tensor map over Tin
//       dim0 = the 8 floats of a chunk (32 B), dim1 = ky (stride 32 B), dim2 = plane*NT + tile (stride Hp*32 B)
// and a box {8, 1, BOXT} brings BOXT chunks of one row into a dense staging array with a single
// cp.async.bulk.tensor (SASS UTMALDG), NBOX = ceil(NT/256) operations per row, completing on an mbarrier; the LSU
// never sees the gather.  The kernel is persistent (gridDim.x CTAs walk the (plane, row) items with stride gridDim.x):
// as soon as a row has been moved from the staging array into the transform buffer (where the Hermitian half is
// expanded), the elected thread requests the NEXT row, so the copy engine works under the row's FFT passes.
// Boxes may run past the last tile of a plane: they then read the first tiles of the next plane (ignored) or, for the
// last plane, zero fill.
#pragma once
#include <cuda.h>

#include "lpc_fused.cuh"
#include "lpc_fused_tma.cuh"

#ifndef LPC_B1_DIRECT
#define LPC_B1_DIRECT 1
#endif
// V / HV are written once and read by the next kernel: no L1 allocation, the twiddle tables stay resident (lpc_util.cuh)
#ifndef LPC_B1_ST
#define LPC_B1_ST st4_hint_na
#endif

namespace lpc {

__host__ __device__ constexpr int rows_tma_wlen(int n, int padsh) { return padsh >= 30 ? n + 1 : n + (n >> padsh) + 1; }  // = fft_padded_len

template <int NT> struct RowBoxes {
    static constexpr int NBOX = (NT + 255) / 256;
    // tiles per box (<= 256); with several boxes a multiple of 4, so that every box lands on a 128-byte boundary of the
    // staging array (a TMA destination must be 128-byte aligned) and the staging array stays dense in kx
    static constexpr int BOXT = NBOX == 1 ? NT : (((NT + NBOX - 1) / NBOX + 3) & ~3);
    static constexpr uint32_t BYTES = (uint32_t)NBOX * BOXT * 32;  // staging bytes = bytes completed per row
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

// shared memory: staging [RowBoxes::BYTES, 128-byte aligned] | W [transform buffer] | mbarrier
template <typename SEQ> inline size_t rows_ifft_tma_smem_bytes() {
    constexpr int Wp = SEQ::N, NT = (Wp / 2 + 1 + LPC_NCT - 1) / LPC_NCT;
    const size_t stage = (RowBoxes<NT>::BYTES + 127) & ~(size_t)127;
    const size_t w = (size_t)((rows_tma_wlen(Wp, SEQ::padsh) + 1) & ~1) * 8;
    return stage + ((w + 15) & ~(size_t)15) + 16;
}

template <typename T, typename SEQ>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 3)
k_rows_ifft_tma(FusedK f, const __grid_constant__ CUtensorMap tmap, int n_items, const typename CxT<T>::type* __restrict__ tw,
                T* __restrict__ VH) {
    using C = typename CxT<T>::type;
    static_assert(sizeof(T) == 4 && LPC_NCT == 2, "float, 32-byte chunks");
    constexpr int NTH = LPC_ROWS_THREADS, NCT = LPC_NCT, Wp = SEQ::N, Nh = Wp / 2 + 1, PADSH = SEQ::padsh;
    constexpr int NT = (Nh + NCT - 1) / NCT;
    using RB = RowBoxes<NT>;
    constexpr size_t STAGE = (RB::BYTES + 127) & ~(size_t)127;
    constexpr size_t WBYTES = ((size_t)((rows_tma_wlen(Wp, PADSH) + 1) & ~1) * 8 + 15) & ~(size_t)15;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float4* S4 = reinterpret_cast<const float4*>(smem_raw);
    C* W = reinterpret_cast<C*>(smem_raw + STAGE);
    const uint32_t bar = smem_u32(smem_raw + STAGE + WBYTES);
    const uint32_t stage_u32 = smem_u32(smem_raw);
    const int tid = threadIdx.x;
    const int Hp = f.Hp;
    const float2 pk = __ldg(reinterpret_cast<const float2*>(f.pack));  // {pack_t, 1/pack_t}

    auto issue = [&](int item) {  // elected thread: request the row of `item`
        const int lin = f.rev ? n_items - 1 - item : item;
        const int p = f.plane0 + lin / Hp, row = lin - (lin / Hp) * Hp;
        mbar_arrive_expect_tx(bar, RB::BYTES);
#pragma unroll
        for (int b = 0; b < RB::NBOX; ++b)
            tma_load_3d(stage_u32 + (uint32_t)b * RB::BOXT * 32, &tmap, 0, row, p * NT + b * RB::BOXT, bar);
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    int item = blockIdx.x;
    if (tid == 0 && item < n_items) issue(item);
    uint32_t phase = 0;
    using TR = SeqTraits<SEQ>;
    constexpr int RL = TR::last, NSL = TR::ns_last;
    const int64_t plane = (int64_t)Hp * Wp;
    for (; item < n_items; item += gridDim.x) {
        const int lin = f.rev ? n_items - 1 - item : item;
        const int p = f.plane0 + lin / Hp, row = lin - (lin / Hp) * Hp;
        mbar_wait(bar, phase);
        phase ^= 1;
#if LPC_B1_DIRECT
        // The first pass takes its input straight from the staging array: element i of Y = t*Vhat + i*HVhat is the 16-byte
        // piece (Vhat[i], HVhat[i]) at byte i*16 for i <= Wp/2 and the conjugate-symmetric image of piece Wp-i above
        // (two real rows per complex transform, see FusedK::pack) -- no expansion sweep, no extra barrier.
        TR::template run_first_ld<T, +1, NTH>(W, tid, [&](int i) {
            const bool lo = i < Nh;
            const float4 c4 = S4[lo ? i : Wp - i];
            const T vx = c4.x * pk.x, vy = c4.y * pk.x;
            C y;
            y.x = lo ? vx - c4.w : vx + c4.w;
            y.y = lo ? vy + c4.z : c4.z - vy;
            return y;
        });
        // (barrier passed: staging consumed, W holds the first pass's output)
        if (tid == 0 && item + (int)gridDim.x < n_items) {
            fence_proxy_async();  // order the generic-proxy reads of the staging array before the async-proxy writes
            issue(item + (int)gridDim.x);
        }
        TR::template run_mid<T, +1, 1, NTH>(W, tw, tid);
#else
        // staging -> transform buffer: column kx = the 16-byte piece (Vhat[kx], HVhat[kx]) at byte kx*16; expand the
        // Hermitian half (two real rows per complex transform: y = t*Vhat + i*HVhat, see FusedK::pack)
        for (int kx = tid; kx < Nh; kx += NTH) {
            const float4 c4 = S4[kx];
            const T vx = c4.x * pk.x, vy = c4.y * pk.x, hx = c4.z, hy = c4.w;
            C y; y.x = vx - hy; y.y = vy + hx;
            W[fft_pad(kx, PADSH)] = y;
            const int m = Wp - kx;
            if (kx >= 1 && m >= Nh) {
                C ym; ym.x = vx + hy; ym.y = hx - vy;
                W[fft_pad(m, PADSH)] = ym;
            }
        }
        __syncthreads();  // staging consumed, W complete
        if (tid == 0 && item + (int)gridDim.x < n_items) {
            fence_proxy_async();  // order the generic-proxy reads of the staging array before the async-proxy writes
            issue(item + (int)gridDim.x);
        }
        TR::template run_butlast<T, +1, 1, NTH>(W, tw, tid);
#endif
        T* Vout = VH + (int64_t)p * plane + (int64_t)row * Wp;
        T* HVout = Vout + (int64_t)f.n_planes * plane;
        for (int j = tid; j < NSL; j += NTH) {
            C v[RL];
            fft_bfly_load<T, RL>(W, 0, 1, PADSH, Wp, j, v);
            fft_bfly_compute<T, RL, +1>(v, tw + TR::tw_off_last, NSL, j);
#pragma unroll
            for (int m = 0; m < RL; ++m) {  // B2 reads them next
                LPC_B1_ST(Vout + j + m * NSL, v[m].x * pk.y, LPC_L2_EVICT_LAST);
                LPC_B1_ST(HVout + j + m * NSL, v[m].y, LPC_L2_EVICT_LAST);
            }
        }
        __syncthreads();  // W is rewritten by the next item
    }
}

}  // namespace lpc
