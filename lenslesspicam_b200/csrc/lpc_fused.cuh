// lpc_fused.cuh -- engine v2: the ADMM iteration in THREE kernels, all transforms hand-written and fused with the
// element-wise work on either side of them:
//
//   k_rows_ifft        (B1): gather Tin -> row IFFT -> V, HV                      [one row per CTA]
//   k_rows_spatial_fft (B2): K1 spatial primal/dual update -> row FFT -> Tout     [RY rows per CTA]
//   k_admm_cols        (A) : column FFT -> K2 spectral update -> column IFFT      [one tile per CTA]
//
// HBM traffic per iteration: B1 2Sp in, 2P out; B2 2P + 4P + d in (V's +-1 rows and eta'0's +1 row
// come from L2), 4P + 2Sp out; A 2Sp + H + R in, 2Sp out  =  12P + 9.5Sp + d ~ 2.17 GB at
// 3x1080x1920 (v1 with cuFFT measured ~3.3 GB; `a`, `bq`, and the un-tangled real spectra never
// reach HBM).  A first version kept V/HV on chip in one fused row kernel (8P + 9.5Sp + d); its
// 107 KB of shared memory per CTA left 16 warps per SM and the kernel latency-bound at
// 0.88 ms (profiles/r01_ncu_full_v2_fused_rows_cols_raw.csv) -- splitting it buys occupancy.
//
// Two real rows per complex transform: B packs z = a + i*xi' (forward) and Y = Vhat + i*HVhat
// (inverse) so every row transform is ONE complex FFT of length Wp (any parity of Wp), with
// the separation  A_row[k] = (Z[k] + conj(Z[-k]))/2,  Xi_row[k] = (Z[k] - conj(Z[-k]))/(2i)
// done in shared memory.  Only the half spectrum k in [0, Wp/2] is stored.
//
// "T-layout" of Tin / Tout (tiled-transposed half spectrum), chosen so that BOTH kernels see
// whole 32-byte sectors:  complex index
//     ((p*NT + kx/NCT)*Hp + ky)*(2*NCT) + (kx%NCT)*2 + slot,     NT = ceil(Nh/NCT), NCT = 2
// (the two slots of a column are adjacent: the column kernel reads / writes the coupled pair as 16 bytes)
// A reads/writes one tile (all ky of NCT columns, both slots) as ONE contiguous 2*NCT*Hp*8-byte
// block; B touches a 2*NCT*8 = 32-byte chunk per (tile, row).
//
// The transforms use compile-time radix sequences (FftSeq, lpc_fft.cuh): the kernels are
// instantiated for the menu of padded sizes below; any other size runs engine v1 (K1/K2 + cuFFT).
//
// Reference semantics reproduced: admm.py:232-329 (see lpc_admm.cuh for the per-pixel algebra,
// which is shared), rfft_convolve.py:133-223 (H, conj(H), ifftshift as the phase S).
#pragma once
#include "lpc_admm.cuh"
#include "lpc_fft.cuh"
#include "lpc_util.cuh"

namespace lpc {

constexpr int LPC_NCT = 2;  // columns per tile and slot
constexpr int LPC_ROWS_THREADS = 256;
constexpr int LPC_COLS_THREADS = 1024;  // variant 0; variant 1 (LPC_COLS_VARIANT=1) uses 576 threads x 2 CTAs/SM

// padded sizes with a compiled fast path: X(N, radices...).  Rows = Wp, columns = Hp.
// 3840x2160: 1080x1920 (headline); 1920x1080 / 960x540: 540x960 and 270x480 (the DiffuserCam dataset's size);
// 2048x1536 / 1024x768 / 512x384: the bundled 3040x4056 data at downsample 4 / 8 / 16; the small ones are the shapes of tests/golden (odd, mixed, gray).
// Round 2 widened the menu to the padded sizes next_fast_len(2d-1) of the usual sensor / dataset dimensions d (64 ... 4056):
// rows and columns are instantiated independently, so any combination of a row size and a column size is served.
// On-demand specialisation (lenslesspicam_b200/_lib.py::load_for): a padded size outside the menu is compiled into its own
// small library by nvcc at first use -- only the two extra entries: -DLPC_MENU_ONLY_EXTRA -DLPC_HAVE_MENU_EXTRA_H with a
// generated lpc_menu_extra.h on the include path that defines LPC_ROW_MENU_EXTRA(X) / LPC_COL_MENU_EXTRA(X) (nvcc
// splits -D values at commas, so the entries cannot ride on the command line) -- and cached next to liblpc.so.
#ifdef LPC_HAVE_MENU_EXTRA_H
#include "lpc_menu_extra.h"
#endif
#ifndef LPC_ROW_MENU_EXTRA
#define LPC_ROW_MENU_EXTRA(X)
#endif
#ifndef LPC_COL_MENU_EXTRA
#define LPC_COL_MENU_EXTRA(X)
#endif
#ifdef LPC_MENU_ONLY_EXTRA
#define LPC_ROW_MENU(X) LPC_ROW_MENU_EXTRA(X)
#define LPC_COL_MENU(X) LPC_COL_MENU_EXTRA(X)
#else
#define LPC_ROW_MENU(X) \
    X(4096, 16, 16, 16) X(3840, 15, 16, 16) X(3200, 10, 16, 10, 2) \
    X(3072, 12, 16, 16) X(2880, 15, 16, 12) X(2560, 10, 16, 16) X(2400, 15, 16, 10) X(2160, 15, 16, 9) X(2048, 16, 16, 8) \
    X(1920, 15, 16, 8) X(1600, 10, 16, 10) X(1536, 6, 16, 16) X(1440, 15, 16, 6) X(1280, 5, 16, 16) X(1200, 15, 16, 5) \
    X(1080, 15, 9, 8) X(1024, 16, 16, 4) X(960, 15, 16, 4) X(800, 5, 16, 10) X(768, 3, 16, 16) X(720, 15, 16, 3) X(640, 5, 16, 8) \
    X(600, 15, 8, 5) X(540, 15, 9, 4) X(512, 16, 8, 4) X(480, 15, 16, 2) X(400, 5, 16, 5) X(384, 3, 16, 8) X(360, 15, 8, 3) \
    X(320, 5, 16, 4) X(288, 9, 16, 2) X(270, 15, 9, 2) X(256, 16, 16) X(240, 15, 16) X(200, 5, 8, 5) X(128, 16, 8) X(125, 5, 5, 5) \
    X(75, 15, 5) X(64, 16, 4) X(54, 9, 6) \
    LPC_ROW_MENU_EXTRA(X)
// (column sequences end in a small radix: that pass is fused with the spectral update in registers; where the first radix
//  can carry every factor of two of N -- 2160 = 16 * 135, 1200 = 16 * 75 ... -- it does: the two outer passes then run
//  warp-locally without a group barrier between them, lpc_fft.cuh: IpWl)
#define LPC_COL_MENU(X) \
    X(4096, 16, 16, 4, 4) X(3840, 16, 16, 5, 3) X(3200, 16, 10, 5, 4) \
    X(3072, 16, 16, 4, 3) X(2880, 15, 16, 4, 3) X(2560, 16, 16, 10) X(2400, 15, 16, 10) X(2160, 16, 15, 9) X(2048, 16, 16, 8) \
    X(1920, 15, 16, 8) X(1600, 16, 10, 10) X(1536, 12, 16, 8) X(1440, 15, 16, 6) X(1280, 16, 16, 5) X(1200, 16, 15, 5) \
    X(1080, 8, 15, 9) X(1024, 16, 16, 4) X(960, 15, 16, 4) X(800, 16, 10, 5) X(768, 16, 16, 3) X(720, 15, 16, 3) X(640, 16, 8, 5) \
    X(600, 8, 15, 5) X(540, 4, 15, 9) X(512, 16, 8, 4) X(480, 15, 8, 4) X(405, 15, 9, 3) X(400, 16, 5, 5) X(384, 16, 8, 3) \
    X(360, 15, 8, 3) X(320, 16, 5, 4) X(270, 15, 9, 2) X(256, 16, 16) X(240, 15, 16) X(200, 8, 5, 5) X(128, 16, 8) X(48, 16, 3) \
    X(45, 15, 3) X(40, 5, 8) X(27, 9, 3) \
    LPC_COL_MENU_EXTRA(X)
#endif

struct FusedK {
    int Hp, Wp, Nh, NT;  // NT tiles of LPC_NCT columns cover Nh = Wp/2+1
    int n_planes, DC;
    int plane0;  // first plane of this launch (the planes of a plan are split over concurrent streams)
    int pf;      // B2: L2-prefetch distance in rows (0 = off), see k_rows_spatial_fft
    int rev;     // 1: walk rows / tiles and planes in DESCENDING order.  Consecutive launches alternate, so a kernel
                 // starts on what its predecessor wrote last -- the part of the hand-off still in the 126 MB L2
    // Two real signals share one complex row transform (z = a + i*xi', Y = Vhat + i*HVhat), and a transform's round-off
    // lands on BOTH outputs in proportion to the LARGER input: H V carries the PSF's DC gain (sum(psf)/||psf||, a few
    // hundred for an L2-normalised PSF) over V, `a` is ~mu3/mu1 times xi'.  The smaller partner is therefore scaled up
    // by an exact power of two before packing and down after separation (no rounding either way), so each signal sees
    // round-off relative to its own magnitude.  Measured at 3x1080x1920 (rms HV / rms V = 126, DC gain 381), error of
    // ADMM-100 against the fp64 reference: pack_t = 1: 2.4e-4, 2^4: 7.8e-5, 2^6: 8.5e-5, 2^8: 1.0e-4 (the reference's own
    // fp32 arithmetic: 1.8e-4; the cuFFT engine, which transforms V and HV separately: 5.8e-5); `a` and xi' turned out
    // balanced already (rms 4.4e-7 / 3.2e-7), pack_s = 1.  profiles/README.md, "packing balance".
    float pack_s, inv_pack_s;  // B2: z = a + i*(pack_s*xi'); separation multiplies Xi by inv_pack_s
    const float* pack;         // device {pack_t, 1/pack_t}: B1 transforms Y = pack_t*Vhat + i*HVhat and stores V * (1/pack_t);
                               // chosen per PSF on the device (k_pack_scale): 2^(round(log2(sqrt(max|H|^2))) - 3), >= 1
};
// logical (row or tile, plane) of a CTA under the launch's direction
__device__ __forceinline__ int lpc_dir_x(const FusedK& f, int n) { return f.rev ? n - 1 - (int)blockIdx.x : (int)blockIdx.x; }
__device__ __forceinline__ int lpc_dir_y(const FusedK& f) { return f.plane0 + (f.rev ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y); }

// L1 allocation of B2's single-use streams (LPC_B2_NA: bit 0 the stores, bit 1 the single-use loads): see lpc_util.cuh
#ifndef LPC_B2_NA
#define LPC_B2_NA 3
#endif
#if LPC_B2_NA & 1
#define LPC_B2_ST st16_hint_na
#else
#define LPC_B2_ST st16_hint
#endif
#if LPC_B2_NA & 2
#define LPC_B2_LDNC ld16_nc_hint_na
#define LPC_B2_LD ld16_hint_na
#else
#define LPC_B2_LDNC ld16_nc_hint
#define LPC_B2_LD ld16_hint
#endif
#ifndef LPC_B2_PF_BULK
#define LPC_B2_PF_BULK 1
#endif
constexpr int LPC_ROWS_RY = 4;  // rows a B2 CTA processes back to back

// ---- B helper: forward FFT of the row held in W (z = a + i*xi'), separate the two real spectra
//      and store the half spectrum to row `row` of Tdst (T-layout)
template <typename T, typename SEQ>
__device__ __forceinline__ void rows_fft_store(typename CxT<T>::type* W, const typename CxT<T>::type* __restrict__ tw,
                                               typename CxT<T>::type* __restrict__ Tdst, int64_t tile_stride, int row,
                                               int NT, T sep_scale, int tid, T sep_scale_im) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS, NCT = LPC_NCT, Wp = SEQ::N, Nh = Wp / 2 + 1, PADSH = SEQ::padsh;
    CtaFftStatic<T, SEQ, -1, 1, NTH>::run(W, tw, tid);
    if constexpr (sizeof(T) == 4 && NCT == 2) {
        // one COLUMN per lane: the 16-byte piece (A[kx], Xi[kx]) of the 32-byte chunk.  Two adjacent lanes fill a
        // chunk (one sector), a warp instruction touches 16 lines instead of the 32 of a chunk-per-lane store
        // (ncu: ~1.6 LSU wavefronts per lane for those), and the shared-memory reads are unit-stride.
        for (int kx = tid; kx < NT * NCT; kx += NTH) {
            Vec16<T> o;
            o.v[0] = o.v[1] = o.v[2] = o.v[3] = T(0);
            if (kx < Nh) {
                const C zk = W[fft_pad(kx, PADSH)];
                const C zm = W[fft_pad(kx == 0 ? 0 : Wp - kx, PADSH)];
                o.v[0] = (zk.x + zm.x) * sep_scale;  // A  = Z[k] + conj(Z[-k])
                o.v[1] = (zk.y - zm.y) * sep_scale;
                o.v[2] = (zk.y + zm.y) * sep_scale_im;  // Xi = (Z[k] - conj(Z[-k])) / i
                o.v[3] = (zm.x - zk.x) * sep_scale_im;
            }
            LPC_B2_ST(reinterpret_cast<T*>(Tdst + (int64_t)(kx / NCT) * tile_stride + (int64_t)row * (2 * NCT) + (kx % NCT) * 2), o,
                      LPC_L2_EVICT_LAST);  // the column kernel reads it next
        }
    } else {
        for (int t = tid; t < NT; t += NTH) {
            C a[NCT], x[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                const int kx = t * NCT + c;
                a[c].x = a[c].y = x[c].x = x[c].y = T(0);
                if (kx < Nh) {
                    const C zk = W[fft_pad(kx, PADSH)];
                    const C zm = W[fft_pad(kx == 0 ? 0 : Wp - kx, PADSH)];
                    a[c].x = (zk.x + zm.x) * sep_scale;  // Z[k] + conj(Z[-k])
                    a[c].y = (zk.y - zm.y) * sep_scale;
                    x[c].x = (zk.y + zm.y) * sep_scale_im;  // (Z[k] - conj(Z[-k])) / i
                    x[c].y = (zm.x - zk.x) * sep_scale_im;
                }
            }
            C* dst = Tdst + (int64_t)t * tile_stride + (int64_t)row * (2 * NCT);
#pragma unroll
            for (int c = 0; c < NCT; ++c) { dst[2 * c] = a[c]; dst[2 * c + 1] = x[c]; }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// B1: row IFFT.  One CTA per row: gather the row of Tin, inverse FFT in shared memory, write the
// V row and the HV row (planar spatial, VH[2][planes][Hp][Wp]).  30 KB of shared memory per CTA at
// Wp = 3840, so several CTAs per SM hide the gather latency.        grid (Hp, planes)
// Also what _form_image / lpc_get_state use to materialise V and HV.
// ---------------------------------------------------------------------------------------------
template <typename T, typename SEQ>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 3)
k_rows_ifft(FusedK f, const typename CxT<T>::type* __restrict__ tw, const typename CxT<T>::type* __restrict__ Tin,
            T* __restrict__ VH) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS, NCT = LPC_NCT, Wp = SEQ::N, Nh = Wp / 2 + 1, PADSH = SEQ::padsh;
    constexpr int NT = (Nh + NCT - 1) / NCT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* W = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int Hp = f.Hp;
    const int row = lpc_dir_x(f, Hp), p = lpc_dir_y(f);
    const int64_t tile_stride = (int64_t)Hp * (2 * NCT);
    const C* src0 = Tin + (int64_t)p * NT * tile_stride + (int64_t)row * (2 * NCT);
    // one COLUMN per lane: the 16-byte piece (Vhat[kx], HVhat[kx]) of the 32-byte chunk (two adjacent lanes share a
    // sector: 16 lines per warp instruction, unit-stride shared-memory stores); all pieces in flight at once
    static_assert(sizeof(T) == 4 && NCT == 2, "the fused engine is float-only with 32-byte chunks");
    const float2 pk = __ldg(reinterpret_cast<const float2*>(f.pack));  // {pack_t, 1/pack_t}
    constexpr int NK = NT * NCT;
    constexpr int IT = (NK + NTH - 1) / NTH;
    Vec16<T> pc[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int kx = tid + it * NTH;
        if (kx < NK) pc[it] = ld16_nc_hint(reinterpret_cast<const T*>(src0 + (int64_t)(kx / NCT) * tile_stride + (kx % NCT) * 2),
                                           LPC_L2_EVICT_FIRST);  // last use of Tin
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int kx = tid + it * NTH;
        if (kx < Nh) {
            const T vx = pc[it].v[0] * pk.x, vy = pc[it].v[1] * pk.x, hx = pc[it].v[2], hy = pc[it].v[3];
            C y; y.x = vx - hy; y.y = vy + hx;
            W[fft_pad(kx, PADSH)] = y;
            const int m = Wp - kx;
            if (kx >= 1 && m >= Nh) {
                C ym; ym.x = vx + hy; ym.y = hx - vy;
                W[fft_pad(m, PADSH)] = ym;
            }
        }
    }
    __syncthreads();
    // all passes but the last in shared memory; the last pass leaves pixel j + m*(Wp/RL) of the row in
    // register m of butterfly j, which goes straight to HBM (coalesced 4-byte stores per m): no
    // shared-memory round trip and no barrier for the output
    using TR = SeqTraits<SEQ>;
    TR::template run_butlast<T, +1, 1, NTH>(W, tw, tid);
    constexpr int RL = TR::last, NSL = TR::ns_last;
    const int64_t plane = (int64_t)Hp * Wp;
    T* Vout = VH + (int64_t)p * plane + (int64_t)row * Wp;
    T* HVout = Vout + (int64_t)f.n_planes * plane;
    for (int j = tid; j < NSL; j += NTH) {
        C v[RL];
        fft_bfly_load<T, RL>(W, 0, 1, PADSH, Wp, j, v);
        fft_bfly_compute<T, RL, +1>(v, tw + TR::tw_off_last, NSL, j);
#pragma unroll
        for (int m = 0; m < RL; ++m) {  // B2 reads them next
            st4_hint(Vout + j + m * NSL, v[m].x * pk.y, LPC_L2_EVICT_LAST);
            st4_hint(HVout + j + m * NSL, v[m].y, LPC_L2_EVICT_LAST);
        }
    }
}

// warm start (admm.py:154-155,172-176): given V, HV (planar) write their row spectra / Wp into Tin
template <typename T, typename SEQ>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 3)
k_rows_fft(FusedK f, const typename CxT<T>::type* __restrict__ tw, const T* __restrict__ VH,
           typename CxT<T>::type* __restrict__ Tin, T sep_scale) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS, NCT = LPC_NCT, Wp = SEQ::N, Nh = Wp / 2 + 1, PADSH = SEQ::padsh;
    constexpr int NT = (Nh + NCT - 1) / NCT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* W = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int row = blockIdx.x, p = f.plane0 + blockIdx.y;
    const int64_t plane = (int64_t)f.Hp * Wp;
    const T* Vin = VH + (int64_t)p * plane + (int64_t)row * Wp;
    const T* HVin = Vin + (int64_t)f.n_planes * plane;
    const float2 pk = __ldg(reinterpret_cast<const float2*>(f.pack));  // balance the packing (see FusedK::pack)
    for (int j = tid; j < Wp; j += NTH) { C z; z.x = Vin[j] * pk.x; z.y = HVin[j]; W[fft_pad(j, PADSH)] = z; }
    __syncthreads();
    const int64_t tile_stride = (int64_t)f.Hp * (2 * NCT);
    rows_fft_store<T, SEQ>(W, tw, Tin + (int64_t)p * NT * tile_stride, tile_stride, row, NT, sep_scale * pk.y, tid, sep_scale);
}

// ---------------------------------------------------------------------------------------------
// B2: the K1 spatial primal/dual update fused with the forward row FFT.  A CTA takes RY consecutive
// rows of one plane, one after the other: per row it reads V (3 rows), HV, the half duals and the
// measurement with 128-bit loads, writes the new half duals, puts z = a + i*xi' into shared memory,
// transforms it and stores the separated half spectra to Tout.  `a` and `bq` never reach HBM.
// grid (ceil(Hp/RY), planes)
// ---------------------------------------------------------------------------------------------
template <typename T, typename SEQ, int RY>
__global__ void __launch_bounds__(LPC_ROWS_THREADS, 4)
k_rows_spatial_fft(FusedK f, AdmmK<T> k, const typename CxT<T>::type* __restrict__ tw, const T* __restrict__ VH,
                   typename CxT<T>::type* __restrict__ Tout, const T* __restrict__ data, T* __restrict__ xi,
                   T* __restrict__ rho, const T* __restrict__ eta_in, T* __restrict__ eta_out) {
    using C = typename CxT<T>::type;
    constexpr int NTH = LPC_ROWS_THREADS, NCT = LPC_NCT, Wp = SEQ::N, Nh = Wp / 2 + 1, PADSH = SEQ::padsh;
    constexpr int NT = (Nh + NCT - 1) / NCT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* W = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int p = lpc_dir_y(f);
    const int Hp = f.Hp;
    const int64_t plane = (int64_t)Hp * Wp;
    const int64_t base = (int64_t)p * plane;
    const int64_t eta1_off = (int64_t)k.n_planes * plane;
    const T* V = VH + base;
    const T* HV = VH + (int64_t)f.n_planes * plane + base;
    const T* eta0 = eta_in + base;
    const T* eta1 = eta_in + eta1_off + base;
    T* eta0_o = eta_out + base;
    T* eta1_o = eta_out + eta1_off + base;
    T* xip = xi + base;
    T* rhop = rho + base;
    const T* dp = data + (int64_t)data_plane_of(p, k) * k.dplane;
    const int64_t tile_stride = (int64_t)Hp * (2 * NCT);
    C* ToutP = Tout + (int64_t)p * NT * tile_stride;

    const int i0 = lpc_dir_x(f, (Hp + RY - 1) / RY) * RY;
    for (int r = 0; r < RY; ++r) {
        const int i = i0 + r;
        if (i >= Hp) break;
        const int iu = (i == 0) ? Hp - 1 : i - 1;
        const int id = (i == Hp - 1) ? 0 : i + 1;
        const int64_t ro = (int64_t)i * Wp, ru = (int64_t)iu * Wp, rd = (int64_t)id * Wp;
        const bool row_in = (i >= k.sy) && (i < k.sy + k.h);
        if constexpr (Wp % 4 == 0 && sizeof(T) == 4) {
            constexpr int NV = Wp / 4;
            constexpr int IT = (NV + NTH - 1) / NTH;
#pragma unroll 2
            for (int it = 0; it < IT; ++it) {
                const int q = tid + it * NTH;
                if (q < NV) {
                    const int j = 4 * q;
                    const int jr = (j + 4 == Wp) ? 0 : j + 4;
                    const int jl = (j == 0) ? Wp - 1 : j - 1;
                    // the CTA that walks LAST over a row of V / eta'0 (row i-1's successor in the sweep direction) frees it
                    const unsigned long long pol_up = f.rev ? LPC_L2_EVICT_NORMAL : LPC_L2_HALO;
                    const unsigned long long pol_dn = f.rev ? LPC_L2_HALO : LPC_L2_EVICT_NORMAL;
                    const Vec16<T> vc = ld16_nc(V + ro + j), vp = ld16_nc_hint(V + ru + j, pol_up), vn = ld16_nc_hint(V + rd + j, pol_dn);
                    const Vec16<T> hv = LPC_B2_LDNC(HV + ro + j, LPC_L2_EVICT_FIRST);
                    const Vec16<T> e0 = ld16_nc_hint(eta0 + ro + j, pol_up), e0d = ld16_nc_hint(eta0 + rd + j, pol_dn);
                    const Vec16<T> e1 = ld16_nc_hint(eta1 + ro + j, LPC_L2_STREAM);  // (its neighbour column e1r re-reads the line)
                    const Vec16<T> xih = LPC_B2_LD(xip + ro + j, LPC_L2_STREAM), rhoh = LPC_B2_LD(rhop + ro + j, LPC_L2_STREAM);
                    const T e1r = __ldg(eta1 + ro + jr);
                    const T vl = __ldg(V + ro + jl), vr = __ldg(V + ro + jr);
                    Vec16<T> y;
                    y.v[0] = y.v[1] = y.v[2] = y.v[3] = T(0);
                    if (row_in && j >= k.dcol0 && j < k.dcol0 + k.dpitch)
                        y = LPC_B2_LDNC(dp + (int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0), LPC_L2_STREAM);
                    Vec16<T> t1, e1new, e0new, xin, rhon;
                    T t0[4], t0d[4], dummy, av[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const T left = (e == 0) ? vl : vc.v[e - 1];
                        t1.v[e] = tv_component<T, true>(left - vc.v[e], e1.v[e], k, e1new.v[e]);
                        t0[e] = tv_component<T, true>(vp.v[e] - vc.v[e], e0.v[e], k, e0new.v[e]);
                        t0d[e] = tv_component<T, true>(vc.v[e] - vn.v[e], e0d.v[e], k, dummy);
                    }
                    const T t1r = tv_component<T, true>(vc.v[3] - vr, e1r, k, dummy);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int jj = j + e;
                        const bool inw = row_in && (jj >= k.sx) && (jj < k.sx + k.w);
                        point_update<T, true>(vc.v[e], hv.v[e], y.v[e], inw, xih.v[e], rhoh.v[e], k, xin.v[e], rhon.v[e]);
                        const T t1right = (e == 3) ? t1r : t1.v[e + 1];
                        av[e] = -rhon.v[e] + ((t0d[e] - t0[e]) + (t1right - t1.v[e]));
                    }
                    if constexpr (PADSH >= 30) {  // z = a + i*xi' for 4 pixels as two 16-byte shared stores
                        Vec16<T> lo, hi;
                        lo.v[0] = av[0]; lo.v[1] = xin.v[0] * f.pack_s; lo.v[2] = av[1]; lo.v[3] = xin.v[1] * f.pack_s;
                        hi.v[0] = av[2]; hi.v[1] = xin.v[2] * f.pack_s; hi.v[2] = av[3]; hi.v[3] = xin.v[3] * f.pack_s;
                        // a lane's two 16-byte pieces are 32 bytes apart from its neighbour's: written in the same order by
                        // all lanes, lanes q and q+4 of a quarter-warp hit the same banks (2-way conflict, measured 2x the ideal
                        // wavefronts).  Lanes 4..7 of every 8 store their pieces in the opposite order.
                        const bool sw = (q >> 2) & 1;
                        *reinterpret_cast<Vec16<T>*>(W + j + (sw ? 2 : 0)) = sw ? hi : lo;
                        *reinterpret_cast<Vec16<T>*>(W + j + (sw ? 0 : 2)) = sw ? lo : hi;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { C z; z.x = av[e]; z.y = xin.v[e] * f.pack_s; W[fft_pad(j + e, PADSH)] = z; }
                    }
                    LPC_B2_ST(xip + ro + j, xin, LPC_L2_STREAM);  // the duals are not touched again for a whole iteration
                    LPC_B2_ST(rhop + ro + j, rhon, LPC_L2_STREAM);
                    LPC_B2_ST(eta0_o + ro + j, e0new, LPC_L2_STREAM);
                    LPC_B2_ST(eta1_o + ro + j, e1new, LPC_L2_STREAM);
                }
            }
        } else {
            for (int j = tid; j < Wp; j += NTH) {
                const int jl = (j == 0) ? Wp - 1 : j - 1;
                const int jr = (j == Wp - 1) ? 0 : j + 1;
                const T v = V[ro + j];
                T e0n, e1n, dummy;
                const T t0 = tv_component<T, true>(V[ru + j] - v, eta0[ro + j], k, e0n);
                const T t1 = tv_component<T, true>(V[ro + jl] - v, eta1[ro + j], k, e1n);
                const T t0d = tv_component<T, true>(v - V[rd + j], eta0[rd + j], k, dummy);
                const T t1r = tv_component<T, true>(v - V[ro + jr], eta1[ro + jr], k, dummy);
                const bool inw = row_in && (j >= k.sx) && (j < k.sx + k.w);
                T y = T(0);
                if (inw) y = dp[(int64_t)(i - k.sy) * k.dpitch + (j - k.dcol0)];
                T xin, rhon;
                point_update<T, true>(v, HV[ro + j], y, inw, xip[ro + j], rhop[ro + j], k, xin, rhon);
                xip[ro + j] = xin;
                rhop[ro + j] = rhon;
                eta0_o[ro + j] = e0n;
                eta1_o[ro + j] = e1n;
                C z;
                z.x = -rhon + ((t0d - t0) + (t1r - t1));
                z.y = xin * f.pack_s;
                W[fft_pad(j, PADSH)] = z;
            }
        }
        __syncthreads();
        // While this CTA transforms its row (no HBM traffic for a while), pull the streams of the row that the CTA
        // REPLACING it on this SM will read -- f.pf rows further along the sweep, about the number of resident CTAs --
        // into L2.  The streaming half of B2 alone runs at 6.7 TB/s (151 us, timing experiment); with the transform
        // behind it HBM idles whenever too few of an SM's four CTAs are in their streaming phase.
        if (f.pf > 0) {
            const int64_t lin = (int64_t)(p - f.plane0) * Hp + i + (f.rev ? -f.pf : f.pf);
            if (lin >= 0 && lin < (int64_t)gridDim.y * Hp) {
                const int pp = f.plane0 + (int)(lin / Hp), ii = (int)(lin % Hp);
                const int64_t o = (int64_t)pp * plane + (int64_t)ii * Wp;
#if LPC_B2_PF_BULK
                if constexpr ((Wp * (int)sizeof(T)) % 16 == 0) {
                    // one bulk prefetch per stream (cp.async.bulk.prefetch.L2, issued by six lanes) instead of one
                    // prefetch instruction per 128-byte line: the per-line form cost the LSU data pipe ~720 wavefronts per row
                    if (tid < 6) {
                        const T* ptr = tid == 0 ? VH + o : tid == 1 ? VH + (int64_t)f.n_planes * plane + o : tid == 2 ? xi + o
                                     : tid == 3 ? rho + o : tid == 4 ? eta_in + o : eta_in + eta1_off + o;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(Wp * (int)sizeof(T)) : "memory");
                    }
                } else
#endif
                {
                    const T* rows6[6] = {VH + o, VH + (int64_t)f.n_planes * plane + o, xi + o, rho + o, eta_in + o, eta_in + eta1_off + o};
                    constexpr int LPR = (Wp * (int)sizeof(T) + 127) / 128;  // 128-byte lines per row
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        for (int l = tid; l < LPR; l += NTH) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(rows6[a]) + (size_t)l * 128));
                }
            }
        }
        rows_fft_store<T, SEQ>(W, tw, ToutP, tile_stride, i, NT, T(1), tid, f.inv_pack_s);
    }
}

// ---------------------------------------------------------------------------------------------
// A: one CTA owns one tile: LPC_NCT columns x 2 slots = 4 interleaved column transforms (index 2*column + slot) of
// length Hp, resident in shared memory from the load to the store.
//     forward column FFT of (2A_row, 2Xi_row)      radix sequence SEQ
//     Vhat  = Rn * (A - S*conj(H)*Xi)        HVhat = S*H*Vhat        (Rn carries 1/(2*Hp*Wp))
//     inverse column FFT -> row spectra of (V, HV)/Wp      radix sequence REVERSED
// The last forward pass, the spectral update and the first inverse pass happen in registers: with
// the reversed sequence the first inverse butterfly j reads exactly the bins j + m*(Hp/RL) that the
// last forward butterfly j produced, and a thread takes both slots of its column, so the coupled
// update is thread-local.  Saves 4 of the 16 shared-memory sweeps of the tile and 3 barriers.
// Ht: [DC][NT][Hp][NCT] complex, Rt: same shape real (tile layout, built once per PSF).
// twF / twI: twiddle tables of SEQ / of the reversed sequence.           grid (NT, planes)
// ---------------------------------------------------------------------------------------------
template <typename T, typename SEQ, int NTH, int MINB>
__global__ void __launch_bounds__(NTH, MINB)
k_admm_cols(FusedK f, const typename CxT<T>::type* __restrict__ twF, const typename CxT<T>::type* __restrict__ twI,
            const typename CxT<T>::type* __restrict__ Tsrc, typename CxT<T>::type* __restrict__ Tdst,
            const typename CxT<T>::type* __restrict__ Ht, const T* __restrict__ Rt,
            const typename CxT<T>::type* __restrict__ Sy, const typename CxT<T>::type* __restrict__ Sx) {
    using C = typename CxT<T>::type;
    using TR = SeqTraits<SEQ>;
    using REV = typename TR::Rev;
    constexpr int NCT = LPC_NCT, NB = 2 * NCT;
    constexpr int Hp = SEQ::N, PF = SEQ::padsh, PI = REV::padsh, RL = TR::last, NSL = TR::ns_last;
    constexpr int n = Hp * NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int NT = f.NT;
    const int tile = blockIdx.x;
    const int p = f.plane0 + blockIdx.y;
    const int q = p % f.DC;
    const C* src = Tsrc + ((int64_t)p * NT + tile) * n;
    C* dst = Tdst + ((int64_t)p * NT + tile) * n;
    const C* Hq = Ht + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT;
    const T* Rq = Rt + ((int64_t)q * NT + tile) * (int64_t)Hp * NCT;
    // the tile is one contiguous block: 16-byte loads, all of a thread's loads in flight before the
    // first shared-memory store; H and R (needed after the forward FFT) are pulled into L2 meanwhile
    {
        constexpr int NV = n / 2;  // 16-byte vectors (2 complex) in the tile
        constexpr int IT = (NV + NTH - 1) / NTH;
        Vec16<T> buf[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int v = tid + it * NTH;
            if (v < NV) buf[it] = ld16_nc(reinterpret_cast<const T*>(src) + 4 * v);
        }
        for (int e = tid * 16; e < Hp * NCT; e += NTH * 16) {  // one prefetch per 128-byte line of H, per 64 B of R
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Hq + e));
            if ((e & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(Rq + e));
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int v = tid + it * NTH;
            if (v < NV) {
                C a, b;
                a.x = buf[it].v[0]; a.y = buf[it].v[1]; b.x = buf[it].v[2]; b.y = buf[it].v[3];
                sm[fft_pad(2 * v, PF)] = a;
                sm[fft_pad(2 * v + 1, PF)] = b;
            }
        }
    }
    __syncthreads();
    TR::template run_butlast<T, -1, NB, NTH>(sm, twF, tid);

    // ---- last forward pass + spectral update + first inverse pass, in registers
    constexpr int totalp = NSL * NCT;  // (column, butterfly) pairs; a thread owns both slots of its column
    constexpr int MAXP = (totalp + NTH - 1) / NTH;
    C v0[MAXP][RL], v1[MAXP][RL];
#pragma unroll
    for (int it = 0; it < MAXP; ++it) {
        const int g = tid + it * NTH;
        if (MAXP * NTH == totalp || g < totalp) {
            const int c = g % NCT, j = g / NCT;
            fft_bfly_load<T, RL>(sm, 2 * c, NB, PF, Hp, j, v0[it]);
            fft_bfly_load<T, RL>(sm, 2 * c + 1, NB, PF, Hp, j, v1[it]);
            fft_bfly_compute<T, RL, -1>(v0[it], twF + TR::tw_off_last, NSL, j);
            fft_bfly_compute<T, RL, -1>(v1[it], twF + TR::tw_off_last, NSL, j);
            const int kx = tile * NCT + c;
            const C sxv = __ldg(Sx + (kx < f.Nh ? kx : 0));
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                const int ky = j + m * NSL;
                const int e = ky * NCT + c;
                const C Hh = __ldg(Hq + e);
                const T Rn = __ldg(Rq + e);
                const C S = cx_mul(__ldg(Sy + ky), sxv);
                const C Hs = cx_mul(S, Hh);
                C Hc = Hh; Hc.y = -Hc.y;
                const C Hsa = cx_mul(S, Hc);
                const C mm = cx_mul(Hsa, v1[it][m]);
                C Vh; Vh.x = Rn * (v0[it][m].x - mm.x); Vh.y = Rn * (v0[it][m].y - mm.y);
                v0[it][m] = Vh;
                v1[it][m] = cx_mul(Hs, Vh);
            }
            Dft<T, RL, +1>::run(v0[it]);  // first pass of the reversed sequence: Ns = 1, no twiddles
            Dft<T, RL, +1>::run(v1[it]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MAXP; ++it) {
        const int g = tid + it * NTH;
        if (MAXP * NTH == totalp || g < totalp) {
            const int c = g % NCT, j = g / NCT;
            fft_bfly_store<T, RL>(sm, 2 * c, NB, PI, 1, j, v0[it]);
            fft_bfly_store<T, RL>(sm, 2 * c + 1, NB, PI, 1, j, v1[it]);
        }
    }
    __syncthreads();
    SeqTraits<REV>::template run_butfirst<T, +1, NB, NTH>(sm, twI, tid);
    for (int v = tid; v < n / 2; v += NTH) {
        const C a = sm[fft_pad(2 * v, PI)], b = sm[fft_pad(2 * v + 1, PI)];
        Vec16<T> o;
        o.v[0] = a.x; o.v[1] = a.y; o.v[2] = b.x; o.v[3] = b.y;
        st16(reinterpret_cast<T*>(dst) + 4 * v, o);
    }
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
