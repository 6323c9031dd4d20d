// lpc_api.cu -- C-ABI (include/lpc.h) over the sm_100a kernels and cuFFT.
//
// Host-side only: owns the planar solver state in HBM, the cuFFT plans and the CUDA
// graph of the iteration loop.  No torch types, no exceptions across the boundary.
#include "../../include/lpc.h"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "lpc_admm.cuh"
#include "lpc_fft_plan.h"
#include "lpc_fused.cuh"
#include "lpc_fused_gd.cuh"
#include "lpc_fused_tma.cuh"
#include "lpc_fused_rows_tma.cuh"
#include "lpc_gd.cuh"
#include "lpc_metrics.cuh"
#include "lpc_io.cuh"
#include "lpc_ops.cuh"
#include "lpc_pnp.cuh"
#include "lpc_setup.cuh"
#include "lpc_util.cuh"

using namespace lpc;

namespace {

thread_local std::string g_create_error;

// Every entry point runs on the plan's device and RESTORES the caller's current device on return: torch reads its
// current device from the CUDA runtime, so a bare cudaSetDevice here would silently re-target the caller's later work.
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

int next_fast_len_impl(int n) {
    // smallest 2^a 3^b 5^c >= n  (scipy.fftpack.next_fast_len; rfft_convolve.py:16,112)
    if (n <= 6) return n < 0 ? 0 : n;
    long long best = -1;
    for (long long p5 = 1; p5 < 2LL * n; p5 *= 5) {
        for (long long p35 = p5; p35 < 2LL * n; p35 *= 3) {
            long long c = p35;
            while (c < n) c *= 2;
            if (best < 0 || c < best) best = c;
        }
    }
    return (int)best;
}

template <typename T> struct FftTraits;
template <> struct FftTraits<float> {
    static constexpr cufftType R2C = CUFFT_R2C, C2R = CUFFT_C2R;
    static cufftResult fwd(cufftHandle p, float* in, float2* out) { return cufftExecR2C(p, in, (cufftComplex*)out); }
    static cufftResult inv(cufftHandle p, float2* in, float* out) { return cufftExecC2R(p, (cufftComplex*)in, out); }
};
template <> struct FftTraits<double> {
    static constexpr cufftType R2C = CUFFT_D2Z, C2R = CUFFT_Z2D;
    static cufftResult fwd(cufftHandle p, double* in, double2* out) { return cufftExecD2Z(p, in, (cufftDoubleComplex*)out); }
    static cufftResult inv(cufftHandle p, double2* in, double* out) { return cufftExecZ2D(p, (cufftDoubleComplex*)in, out); }
};

}  // namespace

// ---- compiled menu of the fused engine (float only): padded size -> kernel instantiation
namespace {
using IfftKern = void (*)(FusedK, const float2*, const float2*, float*);
using IfftTmaKern = void (*)(FusedK, const CUtensorMap, int, const float2*, float*);
using FftKern = void (*)(FusedK, const float2*, const float*, float2*, float);
using SpatKern = void (*)(FusedK, AdmmK<float>, const float2*, const float*, float2*, const float*, float*, float*,
                          const float*, float*);
using ColsKern = void (*)(FusedK, const float2*, const float2*, const float2*, float2*, const float2*, const float*,
                          const float2*, const float2*);
using GdRowsKern = void (*)(GdFusedK, const float2*, const float2*, float2*, const float*, float*, float*, const float*,
                            float, int);
using GdColsKern = void (*)(GdFusedK, const float2*, const float2*, float2*, const float2*);
struct RowsEntry { int N, npass, radix[4]; IfftKern ifft; FftKern fft; SpatKern spatial[3]; GdRowsKern gd[5]; IfftTmaKern ifft_tma;
                   size_t smem_ifft_tma; int nbox, boxt; };
using ColsTmaKern = void (*)(FusedK, ColsTmaK, int, const float2*, const float2*, float2*, const float2*, const float*,
                             const float2*);
// threads per compute group of the persistent column kernel: two groups of 288 (measured at Hp = 2160, A per launch:
// 1x960 232 us, 2x480 175, 3x288 without a look-ahead buffer 160, 2x384 155, 2x288 147 -- 96 registers, no spills)
#ifndef LPC_TMA_NTHG_DEF
#define LPC_TMA_NTHG_DEF 288
#endif
constexpr int LPC_TMA_NG = 2, LPC_TMA_NTHG = LPC_TMA_NTHG_DEF;
using GdColsTmaKern = void (*)(GdFusedK, int, const float2*, float2*, const float2*);
struct ColsEntry { int N, npass, radix[4]; ColsKern kern[2]; GdColsKern gd; ColsTmaKern tma; GdColsTmaKern gd_tma; };
template <int... Rs> constexpr int seq_len() { return (int)sizeof...(Rs); }
#define LPC_X(N, ...)                                                                                   \
    {N, seq_len<__VA_ARGS__>(), {__VA_ARGS__}, k_rows_ifft<float, FftSeq<N, __VA_ARGS__>>,              \
     k_rows_fft<float, FftSeq<N, __VA_ARGS__>>,                                                         \
     {k_rows_spatial_fft<float, FftSeq<N, __VA_ARGS__>, 1>, nullptr, nullptr},                                       \
     {k_gd_rows<float, FftSeq<N, __VA_ARGS__>, 0>, k_gd_rows<float, FftSeq<N, __VA_ARGS__>, 1>,                   \
      k_gd_rows<float, FftSeq<N, __VA_ARGS__>, 2>, k_gd_rows<float, FftSeq<N, __VA_ARGS__>, 3>,                   \
      k_gd_rows<float, FftSeq<N, __VA_ARGS__>, 4>},                                                                \
     k_rows_ifft_tma<float, FftSeq<N, __VA_ARGS__>>, rows_ifft_tma_smem_bytes<FftSeq<N, __VA_ARGS__>>(),            \
     RowBoxes<(N / 2 + 1 + LPC_NCT - 1) / LPC_NCT>::NBOX, RowBoxes<(N / 2 + 1 + LPC_NCT - 1) / LPC_NCT>::BOXT},
const RowsEntry g_rows_menu[] = {LPC_ROW_MENU(LPC_X)};
#undef LPC_X
#define LPC_X(N, ...) \
    {N, seq_len<__VA_ARGS__>(), {__VA_ARGS__}, {k_admm_cols<float, FftSeq<N, __VA_ARGS__>, 1024, 1>, k_admm_cols<float, FftSeq<N, __VA_ARGS__>, 576, 2>}, \
     k_gd_cols<float, FftSeq<N, __VA_ARGS__>, 576, 2>,                                                              \
     k_admm_cols_tma<float, FftSeq<N, __VA_ARGS__>, LPC_TMA_NG, LPC_TMA_NTHG>,                                      \
     k_gd_cols_tma<float, FftSeq<N, __VA_ARGS__>, LPC_TMA_NG, LPC_TMA_NTHG>},
const ColsEntry g_cols_menu[] = {LPC_COL_MENU(LPC_X)};
#undef LPC_X
const RowsEntry* find_rows(int N) {
    for (const RowsEntry& e : g_rows_menu) if (e.N == N) return &e;
    return nullptr;
}
const ColsEntry* find_cols(int N) {
    for (const ColsEntry& e : g_cols_menu) if (e.N == N) return &e;
    return nullptr;
}
}  // namespace

// ------------------------------------------------------------------------------------
struct lpc_plan_s {
    int device = 0, algo = 0, dtype = 0, norm = 0, B = 1, D = 1, C = 1, h = 0, w = 0, flags = 0;
    int Hp = 0, Wp = 0, sy = 0, sx = 0, Nh = 0, Np = 0, DC = 0;
    int n_psf = 1;  // PSFs the plan holds: 1 (shared by every sample) or B (LPC_FLAG_PSF_PER_SAMPLE)
    int PQ = 0;     // transfer-function planes = n_psf * D * C; plane p uses q = p % PQ
    int NC = 0;     // columns of the per-PSF constants table = n_psf * C
    int PS = 0;     // 1: per-sample constants (lpc_const_col)
    std::string err;
    virtual ~lpc_plan_s() {}
    virtual int init() = 0;
    virtual int set_psf(const void*, cudaStream_t) = 0;
    virtual int set_admm_params(double, double, double, double, cudaStream_t) = 0;
    virtual int set_gd_params(double, double, double, cudaStream_t) = 0;
    virtual int set_admm_schedule(int, const double*, const double*, const double*, const double*, cudaStream_t) = 0;
    virtual int set_fista_schedule(int, const double*, const double*, cudaStream_t) = 0;
    virtual int set_data(const void*, int, const int64_t*, const void*, const int64_t*, void*, cudaStream_t) = 0;
    virtual int set_data_host(const void*, int, cudaStream_t) = 0;
    virtual int reset(const void*, cudaStream_t) = 0;
    virtual int run(int, cudaStream_t) = 0;
    virtual int get_image(void*, cudaStream_t) = 0;
    virtual int get_image_host(void*, cudaStream_t) = 0;
    virtual int get_state(int, void*, cudaStream_t) = 0;
    virtual int get_H(void*, cudaStream_t) = 0;
    virtual int get_alpha(void*, cudaStream_t) = 0;
    virtual int convolve(const void*, const int64_t*, void*, int, int, cudaStream_t) = 0;
    virtual int recon_error(const void*, const void*, int, int, double, void*, cudaStream_t) = 0;
    virtual int set_image_estimate(const void*, cudaStream_t) = 0;
    virtual int set_option(int, double) = 0;
    virtual int pnp_begin(void*, cudaStream_t) = 0;
    virtual int pnp_finish(const void*, cudaStream_t) = 0;
    virtual int bytes_per_iter(double*) = 0;
    virtual int launches_per_iter(int*, int*) = 0;
    virtual int profile(int, cudaStream_t, double*, int*) = 0;
    virtual const char* phase_name(int) = 0;
    virtual int engine_version() = 0;
    virtual size_t plan_bytes() = 0;
};

namespace {

template <typename T>
struct Plan : lpc_plan_s {
    using Cplx = typename Cx<T>::type;
    static constexpr int VEC = 16 / (int)sizeof(T);

    // ---- sizes
    int64_t plane = 0, splane = 0;
    int dcol0 = 0, dpitch = 0, Dd = 1;
    bool vec_ok = false;
    // ---- device memory
    std::vector<void*> allocs;
    size_t total_bytes = 0;
    T* psfpad = nullptr;   // [DC][Hp][Wp] zero-padded PSF (R2C input at set_psf)
    Cplx* Hc = nullptr;    // [DC][Hp][Nh]  H = rfft2(pad(psf)) * norm
    Cplx* Sy = nullptr;    // [Hp] ifftshift phase tables
    Cplx* Sx = nullptr;    // [Nh]
    T* consts = nullptr;   // [4C] hmax | psf min | psf max | alpha
    T* data = nullptr;     // [B*Dd*C][h][dpitch] planar measurement
    T* stage = nullptr;    // channels-last staging buffer for the *_host entry points
    size_t stage_elems = 0;
    // ADMM
    T* VH = nullptr;       // [2][Np][plane]: V, HV   (C2R output)
    T* AX = nullptr;       // [2][Np][plane]: a, xi'  (R2C input)
    T* rho = nullptr;      // [Np][plane]    rho'
    T* eta[2] = {nullptr, nullptr};  // ping-pong [2][Np][plane] eta'
    T* Rn = nullptr;       // [DC][splane]  R_divmat / (Hp*Wp), real
    Cplx* spec = nullptr;  // [nb][splane]
    int eta_cur = 0;
    // GD family / CONV
    T* xpad = nullptr;     // [Np][plane] image estimate inside a zero border (R2C input)
    T* dpad = nullptr;     // [Np][plane] residual inside a zero border (R2C input)
    T* rpad = nullptr;     // [Np][plane] C2R output
    T* mom = nullptr;      // [Np][h][w]  Nesterov p / FISTA xk_prev
    bool xpad_border_dirty = false;
    // engine v2 (lpc_fused.cuh): own FFTs fused with the spatial / spectral updates
    bool v2 = false;
    Cplx* Tin = nullptr;    // T-layout row spectra of (V, HV)/Wp      [Np][NT][Hp][2][NCT]
    Cplx* Tout = nullptr;   // T-layout row spectra of 2(a, xi')
    Cplx* Ht = nullptr;     // H in tile layout [DC][NT][Hp][NCT]
    T* Rt = nullptr;        // Rn/2 in tile layout
    Cplx* tw_row = nullptr; // per-pass twiddle tables
    Cplx* tw_col = nullptr;
    Cplx* tw_row_rev = nullptr;  // tables of the reversed radix sequences (the transform on the other side
    Cplx* tw_col_rev = nullptr;  // of an in-register fused pass runs the reversed order)
    Cplx* spec1 = nullptr;  // [Np][splane], warm start only (lazily allocated)
    // persistent TMA-pipelined column kernel (lpc_fused_tma.cuh)
    int cols_tma = 0;        // 1: persistent TMA-pipelined column kernel, 0: k_admm_cols
    Cplx* tabs_ip = nullptr; // twiddles of the in-place passes | Wj | Sy0 (copied to shared memory by the kernel)
    Cplx* Hm = nullptr;      // H in middle-stage order [DC][NT][RL][Hp/RL][NCT]
    T* Gx = nullptr;         // 2 - 2cos(2 pi kx/Wp), [NT*NCT]
    int* kymap = nullptr;    // position -> frequency after the forward passes
    ColsTmaK ctk{};          // cos/sin/phase of the innermost digit (mu's and scale are filled per launch)
    size_t smem_cols_tma = 0;
    int n_sms = 148;
    FftPlan plan_row{}, plan_col{};
    FusedK fk{};
    size_t smem_row = 0, smem_col = 0;
    bool vh_valid = false;  // VH holds the V / HV of the current iterate
    int dir_toggle = 0;     // direction of the next fused ADMM launch (FusedK::rev); alternates launch by launch
    int dir_enable = 1;
    int b2_prefetch = 256;  // B2: L2-prefetch distance in rows (0 = off).  Measured B2 per launch: off 224 us, 74 rows 223,
                            // 148: 217, 222: 213, 296: 214, 400: 218, 592: 240, 1184: 255 (LPC_B2_PF)
    // engine v2 for the GD family (lpc_fused_gd.cuh)
    bool gd2 = false;
    Cplx* Tg = nullptr;     // G-layout spectral buffer [Np][NTg][h][4], updated in place
    Cplx* Fg = nullptr;     // [2][DC][NTg][Hp][4]: S*H/(2N) and S*conj(H)/(2N)
    T* xu = nullptr;        // [Np][h][w] image estimate, unpadded planar
    GdFusedK gk{};
    int gd_tma = 0;          // GD columns on the persistent TMA pipeline (k_gd_cols_tma)
    Cplx* Fgm = nullptr;     // filters in middle-stage order [2][DC][NTg][RL][Hp/RL][4]
    size_t smem_gd_tma = 0;
    // independent planes run as concurrent chains on side streams (forked / joined inside lpc_run)
    static constexpr int MAX_CHAINS = 4;
    int n_chains = 1;
    cudaStream_t side[MAX_CHAINS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHAINS] = {nullptr, nullptr, nullptr, nullptr};
    // ---- cuFFT
    cufftHandle fwd_psf = 0, fwd = 0, inv = 0, fwd1 = 0, inv1 = 0;
    bool have_fwd_psf = false, have_fwd = false, have_inv = false, have_1 = false;
    void* work = nullptr;
    size_t work_bytes = 0;
    // ---- parameters / flags
    double mu1 = 1e-6, mu2 = 1e-5, mu3 = 4e-5, tau = 1e-4;
    // unrolled per-iteration parameters (unrolled_admm.py:181-234, unrolled_fista.py:102-106); empty = constant
    std::vector<double> s_mu1, s_mu2, s_mu3, s_tau;   // ADMM: [n]
    std::vector<double> s_tk;                          // FISTA: [n+1]
    T* alpha_sched = nullptr;                          // FISTA: device [n][4C], alpha_c of iteration i at [i*4C + 3C + c]
    size_t alpha_cap = 0;
    int n_sched = 0;
    double pm1 = 1e-6, pm2 = 1e-5, pm3 = 4e-5;         // mu of the last enqueued ADMM iteration: completes the half duals
    int gd_it = 0;                                     // FISTA schedule position (iterations since reset)
    double lip = 1.8, nmu = 0.9, tk0 = 1.0, tk = 1.0;
    bool psf_set = false, data_set = false, state_ready = false, fista_first = true;
    // plug-and-play (lpc_pnp.cuh): plain duals xi (pnp_xi), rho (rho), eta (eta[0], first Np planes), and U, X, W
    bool pnp = false, pnp_use_dual = false, pnp_open = false;
    T* pnp_xi = nullptr; T* pnp_X = nullptr; T* pnp_W = nullptr; T* pnp_U = nullptr;
    bool keep_x0 = false;  // LPC_OPT_KEEP_X0: lpc_set_psf leaves the initial-value constants (psf min / max) alone
    // ---- graph cache: a few captured iteration loops, least recently used evicted (an apply() that alternates
    //      iteration counts -- disp_iter chunks, a warm-up run next to the real one -- replays instead of re-capturing)
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int niter = -1, eta = -1, first = -1, vh = -1, gdit = -1;
        double pm[3] = {0, 0, 0};
        double tk = 0;
        cudaStream_t stream = nullptr;
        uint64_t stamp = 0;
    };
    static constexpr int MAX_GRAPHS = 4;
    std::vector<GraphEntry> graphs;
    uint64_t graph_clock = 0;

    ~Plan() override {
        cudaSetDevice(device);
        drop_graph();
        for (int i = 0; i < MAX_CHAINS; ++i) { if (side[i]) cudaStreamDestroy(side[i]); if (ev_join[i]) cudaEventDestroy(ev_join[i]); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (have_fwd_psf) cufftDestroy(fwd_psf);
        if (have_fwd) cufftDestroy(fwd);
        if (have_inv) cufftDestroy(inv);
        if (have_1) { cufftDestroy(fwd1); cufftDestroy(inv1); }
        if (work) cudaFree(work);
        for (void* p : allocs) cudaFree(p);
    }

    int fail(int code, const char* what, const char* detail) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s)", lpc_error_name(code), what, detail ? detail : "");
        err = buf;
        return code;
    }
#define LPC_CK(call)                                                                   \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) return fail(LPC_ECUDA, #call, cudaGetErrorString(e_));  \
    } while (0)
#define LPC_CF(call)                                                                   \
    do {                                                                               \
        cufftResult r_ = (call);                                                       \
        if (r_ != CUFFT_SUCCESS) {                                                     \
            char b_[32];                                                               \
            snprintf(b_, sizeof b_, "cufftResult %d", (int)r_);                        \
            return fail(LPC_ECUFFT, #call, b_);                                        \
        }                                                                              \
    } while (0)
#define LPC_KCHECK() LPC_CK(cudaGetLastError())

    template <typename U> int alloc(U** out, size_t n_elems) {
        void* p = nullptr;
        size_t bytes = n_elems * sizeof(U);
        if (bytes == 0) bytes = sizeof(U);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(LPC_ENOMEM, "cudaMalloc", cudaGetErrorString(e)); }
        e = cudaMemset(p, 0, bytes);
        if (e != cudaSuccess) return fail(LPC_ECUDA, "cudaMemset", cudaGetErrorString(e));
        allocs.push_back(p);
        total_bytes += bytes;
        *out = (U*)p;
        return LPC_OK;
    }

    void drop_graph() {
        for (GraphEntry& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }

    int make_fft(cufftHandle* hnd, bool* have, cufftType type, int batch) {
        LPC_CF(cufftCreate(hnd));
        *have = true;
        LPC_CF(cufftSetAutoAllocation(*hnd, 0));
        int n[2] = {Hp, Wp};
        size_t ws = 0;
        LPC_CF(cufftMakePlanMany(*hnd, 2, n, nullptr, 1, 0, nullptr, 1, 0, type, batch, &ws));
        if (ws > work_bytes) work_bytes = ws;
        return LPC_OK;
    }
    int bind_work() {
        drop_graph();  // captured cuFFT nodes carry the old work-area pointer
        if (work) { cudaFree(work); work = nullptr; }
        if (work_bytes) {
            cudaError_t e = cudaMalloc(&work, work_bytes);
            if (e != cudaSuccess) { cudaGetLastError(); return fail(LPC_ENOMEM, "cudaMalloc(cufft work)", cudaGetErrorString(e)); }
        }
        if (have_fwd_psf) LPC_CF(cufftSetWorkArea(fwd_psf, work));
        if (have_fwd) LPC_CF(cufftSetWorkArea(fwd, work));
        if (have_inv) LPC_CF(cufftSetWorkArea(inv, work));
        if (have_1) { LPC_CF(cufftSetWorkArea(fwd1, work)); LPC_CF(cufftSetWorkArea(inv1, work)); }
        return LPC_OK;
    }

    int init() override {
        LPC_CK(cudaSetDevice(device));
        Hp = next_fast_len_impl(2 * h - 1);
        Wp = next_fast_len_impl(2 * w - 1);
        sy = (Hp - h) / 2;
        sx = (Wp - w) / 2;
        Nh = Wp / 2 + 1;
        Np = B * D * C;
        DC = D * C;
        PS = (flags & LPC_FLAG_PSF_PER_SAMPLE) ? 1 : 0;
        n_psf = PS ? B : 1;
        PQ = n_psf * DC;
        NC = n_psf * C;
        plane = (int64_t)Hp * Wp;
        splane = (int64_t)Hp * Nh;
        vec_ok = (Wp % VEC) == 0;
        if (algo == LPC_ALGO_ADMM) {
            dcol0 = sx - (sx % VEC);
            dpitch = ((sx - dcol0) + w + VEC - 1) / VEC * VEC;
        } else {
            dcol0 = sx;
            dpitch = w;
        }
        int rc;
        if ((rc = alloc(&psfpad, (size_t)PQ * plane))) return rc;
        if ((rc = alloc(&Hc, (size_t)PQ * splane))) return rc;
        if ((rc = alloc(&Sy, (size_t)Hp))) return rc;
        if ((rc = alloc(&Sx, (size_t)Nh))) return rc;
        if ((rc = alloc(&consts, (size_t)4 * NC))) return rc;
        if ((rc = alloc(&data, (size_t)Np * h * dpitch))) return rc;  // sized for Dd == D
        int nb = Np;
        pnp = (flags & LPC_FLAG_PNP) != 0;
        if (pnp) flags |= LPC_FLAG_ENGINE_V1 | LPC_FLAG_NO_GRAPH;  // the split iteration runs on the cuFFT engine, host-driven
        if (algo == LPC_ALGO_ADMM && !(flags & LPC_FLAG_ENGINE_V1) && getenv("LPC_ENGINE_V1") == nullptr) {
            if ((rc = init_v2())) return rc;
        }
        if (algo == LPC_ALGO_ADMM) {
            nb = v2 ? 0 : 2 * Np;
            if ((rc = alloc(&VH, (size_t)2 * Np * plane))) return rc;
            if ((rc = alloc(&AX, (size_t)2 * Np * plane))) return rc;
            if ((rc = alloc(&rho, (size_t)Np * plane))) return rc;
            if ((rc = alloc(&eta[0], (size_t)2 * Np * plane))) return rc;
            if ((rc = alloc(&eta[1], (size_t)2 * Np * plane))) return rc;
            if ((rc = alloc(&Rn, (size_t)PQ * splane))) return rc;
            if (pnp) {
                if ((rc = alloc(&pnp_xi, (size_t)Np * plane))) return rc;
                if ((rc = alloc(&pnp_X, (size_t)Np * plane))) return rc;
                if ((rc = alloc(&pnp_W, (size_t)Np * plane))) return rc;
                if ((rc = alloc(&pnp_U, (size_t)Np * plane))) return rc;
            }
        } else {
            if (algo >= LPC_ALGO_GD && !(flags & LPC_FLAG_ENGINE_V1) && getenv("LPC_ENGINE_V1") == nullptr) {
                if ((rc = init_gd2())) return rc;
            }
            if (gd2) {
                nb = 0;
                if ((rc = alloc(&xu, (size_t)Np * h * w))) return rc;
                if ((rc = alloc(&mom, (size_t)Np * h * w))) return rc;
            } else {
                if ((rc = alloc(&xpad, (size_t)Np * plane))) return rc;
                if ((rc = alloc(&rpad, (size_t)Np * plane))) return rc;
                if (algo != LPC_ALGO_CONV) {
                    if ((rc = alloc(&dpad, (size_t)Np * plane))) return rc;
                    if (algo != LPC_ALGO_GD)
                        if ((rc = alloc(&mom, (size_t)Np * h * w))) return rc;
                }
            }
        }
        if (nb > 0 && (rc = alloc(&spec, (size_t)nb * splane))) return rc;
        if ((rc = make_fft(&fwd_psf, &have_fwd_psf, FftTraits<T>::R2C, PQ))) return rc;
        if (nb > 0) {
            if ((rc = make_fft(&fwd, &have_fwd, FftTraits<T>::R2C, nb))) return rc;
            if ((rc = make_fft(&inv, &have_inv, FftTraits<T>::C2R, nb))) return rc;
        }
        if ((rc = bind_work())) return rc;
        if (v2 || gd2) {
            // the persistent column kernel owns whole SMs (3 tile buffers), so it runs all planes in one chain
            int want = ((v2 && cols_tma) || (gd2 && gd_tma)) ? 1 : 3;
            if (const char* e = getenv("LPC_CHAINS")) want = atoi(e);
            if (want > MAX_CHAINS) want = MAX_CHAINS;
            if (want > Np) want = Np;
            if (want < 1) want = 1;
            n_chains = want;
            if (const char* e = getenv("LPC_PLANE_MAJOR")) plane_major = atoi(e) ? 1 : 0;
            if (const char* e = getenv("LPC_ALTERNATE")) dir_enable = atoi(e) ? 1 : 0;
            if (const char* e = getenv("LPC_B2_PF")) b2_prefetch = atoi(e);
            if (const char* e = getenv("LPC_VH_SHARED")) vh_shared = atoi(e) ? 1 : 0;
            LPC_CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            for (int i = 1; i < n_chains; ++i) {
                LPC_CK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
                LPC_CK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
            }
        }
        // shift tables do not depend on the PSF
        k_shift_table<T><<<(Hp + 127) / 128, 128>>>(Sy, Hp, Hp);
        k_shift_table<T><<<(Nh + 127) / 128, 128>>>(Sx, Nh, Wp);
        LPC_KCHECK();
        LPC_CK(cudaDeviceSynchronize());
        return LPC_OK;
    }

    // twiddle tables of both menu sequences and of their reversed orders; shared-memory sizes cover both paddings
    int upload_twiddles() {
        int rc;
        for (int which = 0; which < 4; ++which) {
            const bool row = which < 2, rev = which & 1;
            const int* rad = row ? rows_entry->radix : cols_entry->radix;
            const int np = row ? rows_entry->npass : cols_entry->npass;
            int seq[4];
            for (int i = 0; i < np; ++i) seq[i] = rev ? rad[np - 1 - i] : rad[i];
            FftPlan pl;
            if (!make_fft_plan_from(row ? Wp : Hp, seq, np, &pl)) return fail(LPC_EINVAL, "make_fft_plan_from", "menu sequence");
            std::vector<Cplx> tw;
            fill_fft_twiddles(pl, &tw);
            Cplx** dst = row ? (rev ? &tw_row_rev : &tw_row) : (rev ? &tw_col_rev : &tw_col);
            if ((rc = alloc(dst, tw.size()))) return rc;
            LPC_CK(cudaMemcpy(*dst, tw.data(), tw.size() * sizeof(Cplx), cudaMemcpyHostToDevice));
        }
        return LPC_OK;
    }
    static size_t padded_bytes(int n_complex) {  // worst case of the two padding schemes (padsh 4 / none)
        return (size_t)((fft_padded_len(n_complex, 4) + 1) & ~1) * sizeof(Cplx);
    }

    // ---------------------------------------------------------------- engine v2 set-up
    const RowsEntry* rows_entry = nullptr;
    const ColsEntry* cols_entry = nullptr;
    int ry_log2 = 0;       // B2 processes 1 << ry_log2 rows per CTA (measured: 1 row 315 us, 2 rows 400, 4 rows 465)
    int cols_variant = 1;  // A: 0 = 1024 threads x 1 CTA/SM, 1 = 576 threads x 2 CTAs/SM

    // power-of-two balancing of the two-for-one row packings (FusedK::pack_s / pack_t), as log2 exponents
    int pack_s_fixed = INT_MIN, pack_t_fixed = INT_MIN;  // LPC_PACK_S / LPC_PACK_T: overrides of the automatic choice
    float* pack_dev = nullptr;  // {pack_t, 1/pack_t}, written per PSF by k_pack_scale
    void set_pack_scales(int es) {
        fk.pack_s = std::ldexp(1.0f, es); fk.inv_pack_s = std::ldexp(1.0f, -es);
    }
    int next_dir() { const int d = dir_enable ? dir_toggle : 0; dir_toggle ^= 1; return d; }
    // B1 through the TMA unit (lpc_fused_rows_tma.cuh): a tensor map over Tin, persistent CTAs
    int b1_tma = 0;
    CUtensorMap tin_map{};
    int init_b1_tma(int max_optin) {
        if constexpr (!std::is_same<T, float>::value) {
            return LPC_OK;
        } else {
            int want = 1;
            if (const char* e = getenv("LPC_B1_TMA")) want = atoi(e) ? 1 : 0;
            if (!want || rows_entry->smem_ifft_tma > (size_t)max_optin) return LPC_OK;
            using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
                qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return LPC_OK; }
            const cuuint64_t gdim[3] = {8, (cuuint64_t)Hp, (cuuint64_t)Np * fk.NT};
            const cuuint64_t gstr[2] = {32, (cuuint64_t)Hp * 32};
            const cuuint32_t box[3] = {8, 1, (cuuint32_t)rows_entry->boxt};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = ((EncodeFn)fn)(&tin_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)Tin, gdim, gstr, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return LPC_OK;  // the LDG gather serves
            LPC_CK(cudaFuncSetAttribute((const void*)rows_entry->ifft_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)rows_entry->smem_ifft_tma));
            LPC_CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
            b1_tma = 1;
            return LPC_OK;
        }
    }
    int launch_rows_ifft(cudaStream_t st, int p0 = 0, int np = -1, T* vh = nullptr) {
        if constexpr (std::is_same<T, float>::value) {
            if (np < 0) np = Np;
            if (b1_tma) {
                FusedK k = fk; k.plane0 = p0; k.rev = next_dir();
                const int n_items = Hp * np;
                int per_sm = 3;
                if (const char* e = getenv("LPC_B1_CTAS")) per_sm = atoi(e);
                const int nblk = n_items < n_sms * per_sm ? n_items : n_sms * per_sm;
                rows_entry->ifft_tma<<<nblk, LPC_ROWS_THREADS, rows_entry->smem_ifft_tma, st>>>(k, tin_map, n_items, tw_row, vh ? vh : VH);
                LPC_KCHECK();
                return LPC_OK;
            }
            dim3 grid(Hp, np);
            FusedK k = fk; k.plane0 = p0; k.rev = next_dir();
            rows_entry->ifft<<<grid, LPC_ROWS_THREADS, smem_row, st>>>(k, tw_row, Tin, vh ? vh : VH);
            LPC_KCHECK();
        }
        return LPC_OK;
    }
    int launch_rows_fft(cudaStream_t st, T sep_scale) {
        if constexpr (std::is_same<T, float>::value) {
            dim3 grid(Hp, Np);
            rows_entry->fft<<<grid, LPC_ROWS_THREADS, smem_row, st>>>(fk, tw_row, VH, Tin, sep_scale);
            LPC_KCHECK();
        }
        return LPC_OK;
    }
    int launch_rows_spatial(cudaStream_t st, const T* eta_i, T* eta_o, int p0 = 0, int np = -1, const T* vh = nullptr, int it = 0) {
        if constexpr (std::is_same<T, float>::value) {
            if (np < 0) np = Np;
            const int ry = 1 << ry_log2;
            dim3 grid((Hp + ry - 1) / ry, np);
            FusedK k = fk; k.plane0 = p0; k.rev = next_dir(); k.pf = b2_prefetch;
            rows_entry->spatial[ry_log2]<<<grid, LPC_ROWS_THREADS, smem_row, st>>>(k, admm_k(it, true), tw_row, vh ? vh : VH, Tout, data,
                                                                          AX + (int64_t)Np * plane, rho, eta_i, eta_o);
            LPC_KCHECK();
        }
        return LPC_OK;
    }
    int launch_cols(cudaStream_t st, int p0 = 0, int np = -1, int it = 0) {
        if constexpr (std::is_same<T, float>::value) {
            if (np < 0) np = Np;
            dim3 grid(fk.NT, np);
            FusedK k = fk; k.plane0 = p0; k.rev = cols_tma ? next_dir() : 0;
            if (cols_tma) {
                const int n_items = fk.NT * np;
                const int nthr = LPC_TMA_NG * LPC_TMA_NTHG + 32;
                const int nblk = n_items < n_sms ? n_items : n_sms;
                if constexpr (std::is_same<T, float>::value) {
                    ColsTmaK ck = ctk;
                    double m[4];
                    admm_params(it, m);
                    ck.mu1 = (float)m[0]; ck.mu2 = (float)m[1]; ck.mu3 = (float)m[2];
                    ck.scale = (float)(0.5 / ((double)Hp * (double)Wp));
                    cols_entry->tma<<<nblk, nthr, smem_cols_tma, st>>>(k, ck, n_items, tabs_ip, Tout, Tin, Hm, Gx, Sx);
                }
                LPC_KCHECK();
                return LPC_OK;
            }
            if (n_sched > 0) { int rc = compute_R(st, it); if (rc) return rc; }  // the R table of THIS iteration's mu
            cols_entry->kern[cols_variant]<<<grid, cols_variant ? 576 : 1024, smem_col, st>>>(k, tw_col, tw_col_rev, Tout, Tin, Ht, Rt, Sy, Sx);
            LPC_KCHECK();
        }
        return LPC_OK;
    }

    // decides whether the fused engine applies (float, both padded sizes in the compiled menu, shared
    // memory fits) and allocates its buffers; otherwise the plan stays on engine v1
    int init_v2() {
        if constexpr (!std::is_same<T, float>::value) {
            return LPC_OK;
        } else {
            rows_entry = find_rows(Wp);
            cols_entry = find_cols(Hp);
            if (!rows_entry || !cols_entry) return LPC_OK;
            if (!make_fft_plan_from(Wp, rows_entry->radix, rows_entry->npass, &plan_row)) return LPC_OK;
            if (!make_fft_plan_from(Hp, cols_entry->radix, cols_entry->npass, &plan_col)) return LPC_OK;
            // exact size when the sequence needs no padding (odd first radix): 4 x (30 KB + 1 KB) fit the 132 KB
            // carve-out step and leave 96 KB of L1 to the twiddle tables and the +-1 rows of V, instead of 64 KB
            smem_row = (plan_row.padsh >= 30) ? (size_t)Wp * sizeof(Cplx) : padded_bytes(Wp);
            smem_col = padded_bytes(2 * LPC_NCT * Hp);
            int max_optin = 0;
            LPC_CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
            if (smem_row > (size_t)max_optin || smem_col > (size_t)max_optin) return LPC_OK;
            LPC_CK(cudaFuncSetAttribute((const void*)rows_entry->ifft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
            LPC_CK(cudaFuncSetAttribute((const void*)rows_entry->fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
            LPC_CK(cudaFuncSetAttribute((const void*)rows_entry->spatial[0], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
            for (int v = 0; v < 2; ++v)
                LPC_CK(cudaFuncSetAttribute((const void*)cols_entry->kern[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
            // non-TMA column kernel: 576 threads x 2 CTAs/SM when two tiles fit an SM, else 1024 threads x 1 (large Hp)
            cols_variant = (2 * (smem_col + 1024) <= (size_t)228 * 1024) ? 1 : 0;
            if (const char* e = getenv("LPC_COLS_VARIANT")) cols_variant = atoi(e) ? 1 : 0;

            fk.Hp = Hp; fk.Wp = Wp; fk.Nh = Nh; fk.NT = (Nh + LPC_NCT - 1) / LPC_NCT;
            if (const char* e = getenv("LPC_PACK_S")) pack_s_fixed = atoi(e);
            if (const char* e = getenv("LPC_PACK_T")) pack_t_fixed = atoi(e);
            set_pack_scales(pack_s_fixed == INT_MIN ? 0 : pack_s_fixed);
            { int rc0; if ((rc0 = alloc(&pack_dev, 2))) return rc0; fk.pack = pack_dev; }
            { const float one[2] = {1.f, 1.f}; LPC_CK(cudaMemcpy(pack_dev, one, sizeof one, cudaMemcpyHostToDevice)); }
            fk.n_planes = Np; fk.DC = PQ;
            int rc;
            const size_t tsz = (size_t)Np * fk.NT * Hp * 2 * LPC_NCT;
            if ((rc = alloc(&Tin, tsz))) return rc;
            if ((rc = alloc(&Tout, tsz))) return rc;
            if ((rc = alloc(&Ht, (size_t)PQ * fk.NT * Hp * LPC_NCT))) return rc;
            if ((rc = alloc(&Rt, (size_t)PQ * fk.NT * Hp * LPC_NCT))) return rc;
            if ((rc = upload_twiddles())) return rc;
            if ((rc = init_cols_tma(max_optin))) return rc;
            if ((rc = init_b1_tma(max_optin))) return rc;
            v2 = true;
            return LPC_OK;
        }
    }
    // persistent TMA-pipelined column kernel: tables in the order its in-register middle stage wants
    int init_cols_tma(int max_optin) {
        if constexpr (!std::is_same<T, float>::value) {
            return LPC_OK;
        } else {
            int want = 1;
            if (const char* e = getenv("LPC_COLS_TMA")) want = atoi(e) ? 1 : 0;
            const int np = cols_entry->npass, RL = cols_entry->radix[np - 1], NBL = Hp / RL;
            std::vector<Cplx> tabs;
            if (!fill_fft_ip_twiddles(Hp, cols_entry->radix, np, &tabs)) return fail(LPC_EINVAL, "fill_fft_ip_twiddles", "menu sequence");
            int tw_len = 0;
            for (int i = 0, M = Hp; i < np; ++i) { tw_len += (cols_entry->radix[i] - 1) * (M / cols_entry->radix[i]); M /= cols_entry->radix[i]; }
            tabs.resize((size_t)tw_len);  // (fill pads an empty table with one entry)
            smem_cols_tma = cols_tma_smem_bytes(Hp, RL, tw_len);
            if (want == 0 || RL > LPC_TMA_MAXR || smem_cols_tma > (size_t)max_optin) return LPC_OK;
            LPC_CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
            LPC_CK(cudaFuncSetAttribute((const void*)cols_entry->tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols_tma));
            // position -> frequency after the forward passes; ky(j, m) = ky0(j) + m*NBL
            std::vector<int> km(Hp);
            for (int pz = 0; pz < Hp; ++pz) km[pz] = fft_ip_freq_of_pos(Hp, cols_entry->radix, np, pz);
            const int h2 = Hp / 2;
            auto phase = [&](long long num, int den) {  // exp(2 pi i num/den), exact integer phase reduction
                const double a = 2.0 * M_PI * (double)(num % den) / (double)den;
                Cplx c; c.x = (T)std::cos(a); c.y = (T)std::sin(a); return c;
            };
            for (int j = 0; j < NBL; ++j) tabs.push_back(phase(km[j * RL], Hp));                    // Wj
            for (int j = 0; j < NBL; ++j) tabs.push_back(phase((long long)km[j * RL] * h2, Hp));    // Sy0
            for (int m = 0; m < RL; ++m) {
                if (km[m] != m * NBL) return fail(LPC_EINVAL, "init_cols_tma", "innermost digit is not the top digit");
                const Cplx w = phase(m, RL);
                ctk.cm[m] = w.x; ctk.sm[m] = w.y;
                const Cplx sy = phase((long long)m * NBL * h2, Hp);
                ctk.sym[m].x = sy.x; ctk.sym[m].y = sy.y;
            }
            ctk.s_folded = (Hp % 2 == 0 && Wp % 2 == 0) ? 1 : 0;
            if (const char* e = getenv("LPC_S_FOLDED")) ctk.s_folded = (atoi(e) && ctk.s_folded) ? 1 : 0;
            int rc;
            if ((rc = alloc(&tabs_ip, tabs.size()))) return rc;
            LPC_CK(cudaMemcpy(tabs_ip, tabs.data(), tabs.size() * sizeof(Cplx), cudaMemcpyHostToDevice));
            if ((rc = alloc(&kymap, (size_t)Hp))) return rc;
            LPC_CK(cudaMemcpy(kymap, km.data(), (size_t)Hp * sizeof(int), cudaMemcpyHostToDevice));
            const int NT = (Nh + LPC_NCT - 1) / LPC_NCT;
            std::vector<T> gx((size_t)NT * LPC_NCT);
            for (int kx = 0; kx < NT * LPC_NCT; ++kx) gx[kx] = (T)(2.0 - 2.0 * std::cos(2.0 * M_PI * (double)(kx % Wp) / (double)Wp));
            if ((rc = alloc(&Gx, gx.size()))) return rc;
            LPC_CK(cudaMemcpy(Gx, gx.data(), gx.size() * sizeof(T), cudaMemcpyHostToDevice));
            if ((rc = alloc(&Hm, (size_t)PQ * NT * Hp * LPC_NCT))) return rc;
            cols_tma = want;
            return LPC_OK;
        }
    }

    int init_gd_tma(int max_optin) {
        if constexpr (!std::is_same<T, float>::value) {
            return LPC_OK;
        } else {
            int want = 1;
            if (const char* e = getenv("LPC_GD_TMA")) want = atoi(e) ? 1 : 0;
            const int np = cols_entry->npass, RL = cols_entry->radix[np - 1];
            std::vector<Cplx> tabs;
            if (!fill_fft_ip_twiddles(Hp, cols_entry->radix, np, &tabs)) return fail(LPC_EINVAL, "fill_fft_ip_twiddles", "menu sequence");
            smem_gd_tma = (size_t)LPC_TMA_NBUF * Hp * LPC_GNC * sizeof(Cplx) + 64 + tabs.size() * sizeof(Cplx);
            // bulk copies need 16-byte aligned addresses: the window starts sy rows (32 bytes each) into the buffer
            if (!want || RL > LPC_TMA_MAXR || smem_gd_tma > (size_t)max_optin) return LPC_OK;
            LPC_CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
            LPC_CK(cudaFuncSetAttribute((const void*)cols_entry->gd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gd_tma));
            std::vector<int> km(Hp);
            for (int pz = 0; pz < Hp; ++pz) km[pz] = fft_ip_freq_of_pos(Hp, cols_entry->radix, np, pz);
            int rc;
            if ((rc = alloc(&tabs_ip, tabs.size()))) return rc;
            LPC_CK(cudaMemcpy(tabs_ip, tabs.data(), tabs.size() * sizeof(Cplx), cudaMemcpyHostToDevice));
            if ((rc = alloc(&kymap, (size_t)Hp))) return rc;
            LPC_CK(cudaMemcpy(kymap, km.data(), (size_t)Hp * sizeof(int), cudaMemcpyHostToDevice));
            if ((rc = alloc(&Fgm, (size_t)2 * PQ * gk.NTg * Hp * LPC_GNC))) return rc;
            gd_tma = 1;
            return LPC_OK;
        }
    }

    // ---------------------------------------------------------------- GD family on the fused engine
    int init_gd2() {
        if constexpr (!std::is_same<T, float>::value) {
            return LPC_OK;
        } else {
            rows_entry = find_rows(Wp);
            cols_entry = find_cols(Hp);
            if (!rows_entry || !cols_entry) return LPC_OK;
            if (!make_fft_plan_from(Wp, rows_entry->radix, rows_entry->npass, &plan_row)) return LPC_OK;
            if (!make_fft_plan_from(Hp, cols_entry->radix, cols_entry->npass, &plan_col)) return LPC_OK;
            smem_row = padded_bytes(Wp);
            smem_col = padded_bytes(LPC_GNC * Hp);
            int max_optin = 0;
            LPC_CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
            if (smem_row > (size_t)max_optin || smem_col > (size_t)max_optin) return LPC_OK;
            for (int m = 0; m < 5; ++m)
                LPC_CK(cudaFuncSetAttribute((const void*)rows_entry->gd[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
            LPC_CK(cudaFuncSetAttribute((const void*)cols_entry->gd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
            gk.Hp = Hp; gk.Wp = Wp; gk.Nh = Nh; gk.NTg = (Nh + LPC_GNC - 1) / LPC_GNC;
            gk.h = h; gk.w = w; gk.sy = sy; gk.sx = sx;
            gk.n_planes = Np; gk.DC = PQ; gk.C = C; gk.D = D; gk.Dd = Dd; gk.NC = NC; gk.PS = PS;
            int rc;
            if ((rc = alloc(&Tg, (size_t)Np * gk.NTg * h * LPC_GNC))) return rc;
            if ((rc = alloc(&Fg, (size_t)2 * PQ * gk.NTg * Hp * LPC_GNC))) return rc;
            if ((rc = upload_twiddles())) return rc;
            if ((rc = init_gd_tma(max_optin))) return rc;
            gd2 = true;
            return LPC_OK;
        }
    }
    int gd2_rows(int mode, T scalar, int first, cudaStream_t st, int p0 = 0, int np = -1, const T* cst = nullptr) {
        if constexpr (std::is_same<T, float>::value) {
            gk.Dd = Dd;
            if (np < 0) np = Np;
            dim3 grid((h + 1) / 2, np);
            GdFusedK k = gk; k.plane0 = p0;
            rows_entry->gd[mode]<<<grid, LPC_ROWS_THREADS, smem_row, st>>>(k, tw_row, tw_row_rev, Tg, data, xu, mom, cst ? cst : consts, scalar, first);
            LPC_KCHECK();
        }
        return LPC_OK;
    }
    int gd2_cols(int adjoint, cudaStream_t st, int p0 = 0, int np = -1) {
        if constexpr (std::is_same<T, float>::value) {
            if (np < 0) np = Np;
            dim3 grid(gk.NTg, np);
            GdFusedK k = gk; k.plane0 = p0;
            if (gd_tma) {
                const Cplx* Fm_ = Fgm + (adjoint ? (int64_t)PQ * gk.NTg * Hp * LPC_GNC : 0);
                const int n_items = gk.NTg * np;
                cols_entry->gd_tma<<<n_items < n_sms ? n_items : n_sms, LPC_TMA_NG * LPC_TMA_NTHG + 32, smem_gd_tma, st>>>(k, n_items, tabs_ip, Tg, Fm_);
                LPC_KCHECK();
                return LPC_OK;
            }
            const Cplx* F = Fg + (adjoint ? (int64_t)PQ * gk.NTg * Hp * LPC_GNC : 0);
            cols_entry->gd<<<grid, 576, smem_col, st>>>(k, tw_col, tw_col_rev, Tg, F);
            LPC_KCHECK();
        }
        return LPC_OK;
    }
    // scalars of the next FISTA update: the t_k recursion of gd.py:238-239, or the next entry of an unrolled
    // schedule (unrolled_fista.py:102-106: alpha[i][c], (t_i - 1)/t_{i+1}, no first-iteration aliasing)
    struct FistaStep { T beta; int first; const T* consts; };
    FistaStep fista_step() {
        FistaStep f;
        if (n_sched > 0) {
            const int i = gd_it < n_sched ? gd_it : n_sched - 1;
            f.beta = (T)((s_tk[i] - 1.0) / s_tk[i + 1]);
            f.first = 0;
            f.consts = alpha_sched + (size_t)i * 4 * NC;
            ++gd_it;
        } else {
            const double tn = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;  // gd.py:238
            f.beta = (T)((tk - 1.0) / tn);
            f.first = fista_first ? 1 : 0;
            f.consts = consts;
            tk = tn;
            fista_first = false;
        }
        return f;
    }
    int gd2_iteration(cudaStream_t st, cudaEvent_t* ev) {
#define LPC_EV(i) do { if (ev) LPC_CK(cudaEventRecord(ev[i], st)); } while (0)
        int rc;
        LPC_EV(0);
        if ((rc = gd2_cols(0, st))) return rc;                       // convolve: columns
        LPC_EV(1);
        if ((rc = gd2_rows(LPC_GD_RESID, T(0), 0, st))) return rc;   // rows back, - y, rows forward
        LPC_EV(2);
        if ((rc = gd2_cols(1, st))) return rc;                       // deconvolve: columns
        LPC_EV(3);
        if (algo == LPC_ALGO_GD) {
            rc = gd2_rows(LPC_GD_UPD_GD, T(0), 0, st);
        } else if (algo == LPC_ALGO_NESTEROV) {
            rc = gd2_rows(LPC_GD_UPD_NESTEROV, (T)nmu, 0, st);
        } else {
            const FistaStep f = fista_step();
            rc = gd2_rows(LPC_GD_UPD_FISTA, f.beta, f.first, st, 0, -1, f.consts);
        }
        if (rc) return rc;
        LPC_EV(4);
#undef LPC_EV
        return LPC_OK;
    }

    int tile_HR(cudaStream_t st) {
        k_tile_HR<T><<<1184, 256, 0, st>>>(Hc, Rn, Ht, Rt, PQ, Hp, Nh, fk.NT);
        LPC_KCHECK();
        if constexpr (std::is_same<T, float>::value) {
            if (cols_tma) {
                const bool fold = ctk.s_folded != 0;
                k_tile_H_mid<T><<<1184, 256, 0, st>>>(Hc, kymap, Hm, PQ, Hp, Nh, fk.NT, cols_entry->radix[cols_entry->npass - 1],
                                                     fold ? Sy : nullptr, fold ? Sx : nullptr);
                LPC_KCHECK();
            }
        }
        return LPC_OK;
    }
    // V, HV of the current iterate -> VH (they are not stored by the fused engine)
    int materialise_vh(cudaStream_t st) {
        if (!v2 || vh_valid) return LPC_OK;
        int rc = launch_rows_ifft(st);
        if (rc) return rc;
        vh_valid = true;
        return LPC_OK;
    }

    // ---------------------------------------------------------------- layout helpers
    static Strides5 contiguous(int Dn, int hh, int ww, int Cn) {
        Strides5 s;
        s.c = 1; s.x = Cn; s.y = (int64_t)ww * Cn; s.d = (int64_t)hh * ww * Cn; s.b = (int64_t)Dn * hh * ww * Cn;
        return s;
    }
    static Strides5 from_array(const int64_t* a) { Strides5 s; s.b = a[0]; s.d = a[1]; s.y = a[2]; s.x = a[3]; s.c = a[4]; return s; }

    int pack(const T* src, Strides5 ss, const T* bg, Strides5 bs, T* dst, int nB, int Dn, int hh, int ww,
             int64_t plane_stride, int pitch, int oy, int ox, cudaStream_t st, T* echo = nullptr) {
        dim3 grid((ww * C + 255) / 256, hh, nB * Dn);
        k_pack<T><<<grid, 256, 0, st>>>(src, ss, bg, bs, dst, echo, Dn, C, hh, ww, plane_stride, pitch, oy, ox);
        LPC_KCHECK();
        return LPC_OK;
    }
    int unpack(const T* src, int64_t plane_stride, int pitch, int oy, int ox, T* dst, int nBD, int hh, int ww,
               int clamp, cudaStream_t st) {
        dim3 grid((ww * C + 255) / 256, hh, nBD);
        k_unpack<T><<<grid, 256, 0, st>>>(src, plane_stride, pitch, oy, ox, dst, C, hh, ww, clamp);
        LPC_KCHECK();
        return LPC_OK;
    }

    // parameters of iteration `it` of a run (the schedule restarts at 0 with every lpc_run)
    void admm_params(int it, double* m) const {
        if (n_sched > 0 && algo == LPC_ALGO_ADMM) {
            const int i = it < n_sched ? it : n_sched - 1;
            m[0] = s_mu1[i]; m[1] = s_mu2[i]; m[2] = s_mu3[i]; m[3] = s_tau[i];
        } else {
            m[0] = mu1; m[1] = mu2; m[2] = mu3; m[3] = tau;
        }
    }
    AdmmK<T> admm_k(int it = 0, bool use_prev = false) const {
        double m[4];
        admm_params(it, m);
        const double mu1 = m[0], mu2 = m[1], mu3 = m[2], tau = m[3];
        AdmmK<T> k;
        k.mu1 = (T)mu1; k.mu2 = (T)mu2; k.mu3 = (T)mu3;
        k.mu1p = use_prev ? (T)pm1 : k.mu1; k.mu2p = use_prev ? (T)pm2 : k.mu2; k.mu3p = use_prev ? (T)pm3 : k.mu3;
        k.thr = (T)(tau / mu2);
        k.inv_mu2 = T(1) / (T)mu2;
        k.inv_mu3 = T(1) / (T)mu3;
        // X_divmat = 1/(pad(ones) + mu1) in the working precision (admm.py:193)
        k.xdiv_in = T(1) / (T(1) + (T)mu1);
        k.xdiv_out = T(1) / (T(0) + (T)mu1);
        k.Hp = Hp; k.Wp = Wp; k.h = h; k.w = w; k.sy = sy; k.sx = sx;
        k.C = C; k.D = D; k.Dd = Dd; k.dcol0 = dcol0; k.dpitch = dpitch;
        k.plane = plane; k.dplane = (int64_t)h * dpitch; k.n_planes = Np;
        return k;
    }
    GdK gd_k() const {
        GdK k;
        k.Hp = Hp; k.Wp = Wp; k.Nh = Nh; k.h = h; k.w = w; k.sy = sy; k.sx = sx;
        k.C = C; k.D = D; k.Dd = Dd; k.dpitch = dpitch; k.plane = plane; k.dplane = (int64_t)h * dpitch;
        k.n_planes = Np; k.NC = NC; k.PS = PS;
        return k;
    }

    // ---------------------------------------------------------------- PSF
    int compute_R(cudaStream_t st, int it = 0) {
        double m[4];
        admm_params(it, m);
        k_admm_R<T><<<1184, 256, 0, st>>>(Hc, Rn, PQ, Hp, Wp, Nh, m[0], m[1], m[2]);
        LPC_KCHECK();
        if (v2) return tile_HR(st);
        return LPC_OK;
    }
    int compute_alpha(cudaStream_t st) {
        k_alpha<T><<<(NC + 127) / 128, 128, 0, st>>>(consts, NC, (T)lip);
        LPC_KCHECK();
        return LPC_OK;
    }

    int set_psf(const void* dev_psf, cudaStream_t st) override {
        LPC_CK(cudaSetDevice(device));
        if (!dev_psf) return fail(LPC_EINVAL, "set_psf", "null PSF");
        // (the captured iteration graph stays valid: a new PSF rewrites the H / filter tables in place, no kernel
        //  argument depends on it -- per-sample PSFs of a multi-mask evaluation loop, eval/benchmark.py:244-245)
        const T* psf = (const T*)dev_psf;
        int rc;
        // pad (rfft_convolve.py:84-100) -> rfft2 (rfft_convolve.py:121-123)
        if ((rc = pack(psf, contiguous(D, h, w, C), nullptr, Strides5{}, psfpad, n_psf, D, h, w, plane, Wp, sy, sx, st))) return rc;
        LPC_CF(cufftSetStream(fwd_psf, st));
        LPC_CF(FftTraits<T>::fwd(fwd_psf, psfpad, Hc));
        if (norm != LPC_NORM_BACKWARD) {
            const double n = (double)Hp * (double)Wp;
            const T scale = (T)(norm == LPC_NORM_ORTHO ? 1.0 / std::sqrt(n) : 1.0 / n);
            k_scale_complex<T><<<1184, 256, 0, st>>>(Hc, (int64_t)PQ * splane, scale);
            LPC_KCHECK();
        }
        // constants derived from H / PSF
        k_init_consts<T><<<(NC + 127) / 128, 128, 0, st>>>(consts, psf, C, NC, (int64_t)D * h * w * C, keep_x0 ? 1 : 0);
        LPC_KCHECK();
        {
            dim3 g(64, PQ);
            k_hh_max<T><<<g, 256, 0, st>>>(Hc, consts, C, DC, PS, splane);
            if (!keep_x0) {
                dim3 g2(64, C, n_psf);
                k_psf_minmax<T><<<g2, 256, 0, st>>>(psf, (int64_t)D * h * w, C, consts + NC, consts + 2 * NC);
            }
            LPC_KCHECK();
        }
        if ((rc = compute_alpha(st))) return rc;
        if (v2) {
            k_pack_scale<T><<<1, 32, 0, st>>>(consts, NC, pack_dev, pack_t_fixed == INT_MIN ? -100000 : pack_t_fixed);
            LPC_KCHECK();
        }
        if (algo == LPC_ALGO_ADMM && (rc = compute_R(st))) return rc;
        if (gd2) {
            k_gd_tile_filter<T><<<1184, 256, 0, st>>>(Hc, Sy, Sx, Fg, PQ, Hp, Wp, Nh, gk.NTg);
            LPC_KCHECK();
            if constexpr (std::is_same<T, float>::value) {
                if (gd_tma) {
                    k_gd_tile_filter_mid<T><<<1184, 256, 0, st>>>(Hc, Sy, Sx, kymap, Fgm, PQ, Hp, Wp, Nh, gk.NTg,
                                                                 cols_entry->radix[cols_entry->npass - 1]);
                    LPC_KCHECK();
                }
            }
        }
        psf_set = true;
        state_ready = false;
        return LPC_OK;
    }

    int set_admm_params(double a, double b, double c, double d, cudaStream_t st) override {
        if (algo != LPC_ALGO_ADMM) return fail(LPC_ESTATE, "set_admm_params", "not an ADMM plan");
        if (!(a > 0) || !(b > 0) || !(c > 0)) return fail(LPC_EINVAL, "set_admm_params", "mu must be > 0");
        LPC_CK(cudaSetDevice(device));
        mu1 = a; mu2 = b; mu3 = c; tau = d;
        drop_graph();
        if (psf_set) return compute_R(st);
        return LPC_OK;
    }
    int set_gd_params(double lip_fact, double nesterov_mu, double fista_tk, cudaStream_t st) override {
        if (algo < LPC_ALGO_GD) return fail(LPC_ESTATE, "set_gd_params", "not a gradient-descent plan");
        LPC_CK(cudaSetDevice(device));
        lip = lip_fact; nmu = nesterov_mu; tk0 = fista_tk;
        drop_graph();
        if (psf_set) return compute_alpha(st);
        return LPC_OK;
    }

    // unrolled per-iteration parameters; n == 0 clears the schedule
    int set_admm_schedule(int n, const double* a, const double* b, const double* c, const double* d, cudaStream_t st) override {
        if (algo != LPC_ALGO_ADMM) return fail(LPC_ESTATE, "set_admm_schedule", "not an ADMM plan");
        if (n < 0 || (n > 0 && (!a || !b || !c || !d))) return fail(LPC_EINVAL, "set_admm_schedule", "bad arguments");
        for (int i = 0; i < n; ++i)
            if (!(a[i] > 0) || !(b[i] > 0) || !(c[i] > 0)) return fail(LPC_EINVAL, "set_admm_schedule", "mu must be > 0");
        LPC_CK(cudaSetDevice(device));
        drop_graph();
        s_mu1.assign(a, a + n); s_mu2.assign(b, b + n); s_mu3.assign(c, c + n); s_tau.assign(d, d + n);
        n_sched = n;
        { double m[4]; admm_params(0, m); pm1 = m[0]; pm2 = m[1]; pm3 = m[2]; }
        if (n == 0 && psf_set) return compute_R(st);  // back to the constant-parameter table
        return LPC_OK;
    }
    int set_fista_schedule(int n, const double* alpha, const double* tks, cudaStream_t st) override {
        if (algo != LPC_ALGO_FISTA) return fail(LPC_ESTATE, "set_fista_schedule", "not a FISTA plan");
        if (n < 0 || (n > 0 && (!alpha || !tks))) return fail(LPC_EINVAL, "set_fista_schedule", "bad arguments");
        LPC_CK(cudaSetDevice(device));
        drop_graph();
        n_sched = 0;
        state_ready = false;  // the previous-iterate convention differs: reset before the next run
        if (n == 0) return LPC_OK;
        std::vector<T> tab((size_t)n * 4 * NC, T(0));  // one constants table per iteration; the steps are shared by all PSFs
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < NC; ++j) tab[(size_t)i * 4 * NC + 3 * NC + j] = (T)alpha[(size_t)i * C + j % C];
        if (tab.size() > alpha_cap) {  // (an outgrown table is released with the plan)
            int rc = alloc(&alpha_sched, tab.size());
            if (rc) return rc;
            alpha_cap = tab.size();
        }
        LPC_CK(cudaMemcpyAsync(alpha_sched, tab.data(), tab.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        LPC_CK(cudaStreamSynchronize(st));  // `tab` is a pageable host temporary
        s_tk.assign(tks, tks + n + 1);
        n_sched = n;
        return LPC_OK;
    }

    // ---------------------------------------------------------------- data
    int set_data(const void* dev_data, int Dd_in, const int64_t* strides, const void* dev_bg,
                 const int64_t* bg_strides, void* dev_echo, cudaStream_t st) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "set_data", "convolver plans take their input in lpc_convolve");
        if (!dev_data) return fail(LPC_EINVAL, "set_data", "null data");
        if (Dd_in != 1 && Dd_in != D) return fail(LPC_EINVAL, "set_data", "data depth must be 1 or the PSF depth");
        LPC_CK(cudaSetDevice(device));
        if (Dd_in != Dd) drop_graph();
        Dd = Dd_in;
        Strides5 ss = strides ? from_array(strides) : contiguous(Dd, h, w, C);
        Strides5 bs = bg_strides ? from_array(bg_strides) : contiguous(Dd, h, w, C);
        // the data planes were zero-filled at allocation and only the true window is ever
        // written, so the alignment slack of the ADMM layout stays zero
        int rc = pack((const T*)dev_data, ss, (const T*)dev_bg, bs, data, B, Dd, h, w, (int64_t)h * dpitch, dpitch, 0,
                      sx - dcol0, st, (T*)dev_echo);
        if (rc) return rc;
        data_set = true;
        return LPC_OK;
    }
    int ensure_stage(size_t elems) {
        if (elems <= stage_elems) return LPC_OK;
        int rc = alloc(&stage, elems);  // the old one (if any) is released with the plan
        if (rc) return rc;
        stage_elems = elems;
        return LPC_OK;
    }
    int set_data_host(const void* host, int Dd_in, cudaStream_t st) override {
        if (!host) return fail(LPC_EINVAL, "set_data_host", "null data");
        if (Dd_in != 1 && Dd_in != D) return fail(LPC_EINVAL, "set_data_host", "data depth must be 1 or the PSF depth");
        LPC_CK(cudaSetDevice(device));
        size_t n = (size_t)B * D * h * w * C;
        int rc = ensure_stage(n);
        if (rc) return rc;
        LPC_CK(cudaMemcpyAsync(stage, host, (size_t)B * Dd_in * h * w * C * sizeof(T), cudaMemcpyHostToDevice, st));
        return set_data(stage, Dd_in, nullptr, nullptr, nullptr, nullptr, st);
    }

    // ---------------------------------------------------------------- reset
    int ensure_single_plans() {
        if (have_1) return LPC_OK;
        int rc;
        bool h1 = false, h2 = false;
        if ((rc = make_fft(&fwd1, &h1, FftTraits<T>::R2C, Np))) return rc;
        if ((rc = make_fft(&inv1, &h2, FftTraits<T>::C2R, Np))) return rc;
        have_1 = true;
        LPC_CK(cudaDeviceSynchronize());  // the shared work area may be re-allocated
        return bind_work();
    }

    int spec_mul(Cplx* s, int nplanes, int adjoint, cudaStream_t st) {
        dim3 grid((unsigned)((splane + 255) / 256), nplanes);
        const T scale = (T)(1.0 / ((double)Hp * (double)Wp));
        k_spec_mul<T><<<grid, 256, 0, st>>>(s, Hc, Sy, Sx, PQ, Hp, Nh, adjoint, scale);
        LPC_KCHECK();
        return LPC_OK;
    }

    int reset(const void* dev_init, cudaStream_t st) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "reset", "convolver plans have no solver state");
        if (!psf_set) return fail(LPC_ESTATE, "reset", "PSF not set");
        LPC_CK(cudaSetDevice(device));
        int rc;
        if (algo == LPC_ALGO_ADMM) {
            // admm.py:150-183: every state array zero, or the warm start
            LPC_CK(cudaMemsetAsync(VH, 0, (size_t)2 * Np * plane * sizeof(T), st));
            LPC_CK(cudaMemsetAsync(AX, 0, (size_t)2 * Np * plane * sizeof(T), st));
            LPC_CK(cudaMemsetAsync(rho, 0, (size_t)Np * plane * sizeof(T), st));
            LPC_CK(cudaMemsetAsync(eta[0], 0, (size_t)2 * Np * plane * sizeof(T), st));
            eta_cur = 0;
            if (pnp) {  // plain duals and U, X, W all start at zero (admm.py:161-183)
                for (T* a : {pnp_xi, pnp_X, pnp_W, pnp_U}) LPC_CK(cudaMemsetAsync(a, 0, (size_t)Np * plane * sizeof(T), st));
                pnp_open = false;
            }
            { double m[4]; admm_params(0, m); pm1 = m[0]; pm2 = m[1]; pm3 = m[2]; }
            if (v2) {
                LPC_CK(cudaMemsetAsync(Tin, 0, (size_t)Np * fk.NT * Hp * 2 * LPC_NCT * sizeof(Cplx), st));
                vh_valid = true;  // V = HV = 0 sits in VH
            }
            Cplx* sp = spec;
            if (dev_init) {
                if ((rc = ensure_single_plans())) return rc;
                if (v2) {
                    if (!spec1 && (rc = alloc(&spec1, (size_t)Np * splane))) return rc;
                    sp = spec1;
                }
                // V = initial_est (padded shape, recon.py:296-299)
                if ((rc = pack((const T*)dev_init, contiguous(D, Hp, Wp, C), nullptr, Strides5{}, VH, B, D, Hp, Wp, plane,
                               Wp, 0, 0, st))) return rc;
                // HV = convolve(V) (admm.py:175)
                LPC_CF(cufftSetStream(fwd1, st));
                LPC_CF(cufftSetStream(inv1, st));
                LPC_CF(FftTraits<T>::fwd(fwd1, VH, sp));
                if ((rc = spec_mul(sp, Np, 0, st))) return rc;
                LPC_CF(FftTraits<T>::inv(inv1, sp, VH + (int64_t)Np * plane));
                if (!pnp) {
                    dim3 grid((Wp + 127) / 128, Hp, Np);
                    k_admm_warm_start<T><<<grid, 128, 0, st>>>(admm_k(), VH, VH + (int64_t)Np * plane, AX + (int64_t)Np * plane,
                                                              rho, eta[0]);
                    LPC_KCHECK();
                }
                if (v2) {  // row spectra of (V, HV)/Wp into Tin
                    if ((rc = launch_rows_fft(st, (T)(0.5 / (double)Wp)))) return rc;
                    vh_valid = true;
                }
            }
        } else {
            // gd.py:94-126
            T* xdst = gd2 ? xu : xpad;
            const int64_t xps = gd2 ? (int64_t)h * w : plane;
            const int xpitch = gd2 ? w : Wp, xoy = gd2 ? 0 : sy, xox = gd2 ? 0 : sx;
            if (dev_init) {
                if ((rc = pack((const T*)dev_init, contiguous(D, h, w, C), nullptr, Strides5{}, xdst, B, D, h, w, xps, xpitch,
                               xoy, xox, st))) return rc;
            } else {
                dim3 grid((w + 127) / 128, h, Np);
                k_fill_x0<T><<<grid, 128, 0, st>>>(xdst, consts, C, DC, NC, PS, h, w, xps, xpitch, xoy, xox);
                LPC_KCHECK();
            }
            if (gd2 && (rc = gd2_rows(LPC_GD_INIT, T(0), 0, st))) return rc;  // row spectra of the initial estimate
            if (mom) LPC_CK(cudaMemsetAsync(mom, 0, (size_t)Np * h * w * sizeof(T), st));
            if (n_sched > 0 && algo == LPC_ALGO_FISTA) {
                // unrolled_fista.py:95-100: the previous iterate starts as the initial estimate itself
                if (gd2) LPC_CK(cudaMemcpyAsync(mom, xu, (size_t)Np * h * w * sizeof(T), cudaMemcpyDeviceToDevice, st));
                else if ((rc = copy_window(xpad, mom, st))) return rc;
            }
            gd_it = 0;
            tk = tk0;            // gd.py:227-233
            fista_first = true;  // xk aliases x right after reset (gd.py:233)
            pnp_open = false;
        }
        state_ready = true;
        return LPC_OK;
    }

    // ---------------------------------------------------------------- iterations
    // ev (optional): n_phases+1 events recorded at the phase boundaries (lpc_profile)
    // bookkeeping after iteration `it` has been enqueued: its mu completes the next iteration's half duals
    void admm_advance(int it) {
        double m[4];
        admm_params(it, m);
        pm1 = m[0]; pm2 = m[1]; pm3 = m[2];
    }
    int admm_iteration(cudaStream_t st, cudaEvent_t* ev = nullptr, int it = 0) {
#define LPC_EV(i) do { if (ev) LPC_CK(cudaEventRecord(ev[i], st)); } while (0)
        LPC_EV(0);
        if (v2) {
            int rc = materialise_vh(st);  // B1: row IFFT -> V, HV (skipped when VH already holds them, e.g. after reset)
            if (rc) return rc;
            LPC_EV(1);
            if ((rc = launch_rows_spatial(st, eta[eta_cur], eta[eta_cur ^ 1], 0, -1, nullptr, it))) return rc;  // B2
            eta_cur ^= 1;
            vh_valid = false;
            LPC_EV(2);
            if ((rc = launch_cols(st, 0, -1, it))) return rc;  // A
            LPC_EV(3);
            admm_advance(it);
            return LPC_OK;
        }
        const AdmmK<T> k = admm_k(it, true);
        if (n_sched > 0) { int rc = compute_R(st, it); if (rc) return rc; }
        T* a = AX;
        T* xi = AX + (int64_t)Np * plane;
        const T* V = VH;
        const T* HV = VH + (int64_t)Np * plane;
        if (vec_ok) {
            constexpr int RY = 8;
            const int nvec = Wp / VEC;
            const int64_t nwarps = (int64_t)Np * ((Hp + RY - 1) / RY) * ((nvec + 31) / 32);
            const unsigned blocks = (unsigned)((nwarps + 7) / 8);
            k_admm_spatial_vec<T, RY><<<blocks, 256, 0, st>>>(k, V, HV, data, xi, rho, eta[eta_cur], eta[eta_cur ^ 1], a);
        } else {
            dim3 grid((Wp + 127) / 128, Hp, Np);
            k_admm_spatial_generic<T><<<grid, 128, 0, st>>>(k, V, HV, data, xi, rho, eta[eta_cur], eta[eta_cur ^ 1], a);
        }
        LPC_KCHECK();
        eta_cur ^= 1;
        LPC_EV(1);
        LPC_CF(FftTraits<T>::fwd(fwd, AX, spec));
        LPC_EV(2);
        dim3 grid((unsigned)((splane + 255) / 256), Np);
        k_admm_spectral<T><<<grid, 256, 0, st>>>(spec, Hc, Rn, Sy, Sx, Np, PQ, Hp, Nh);
        LPC_KCHECK();
        LPC_EV(3);
        LPC_CF(FftTraits<T>::inv(inv, spec, VH));
        LPC_EV(4);
        admm_advance(it);
        return LPC_OK;
    }

    int gd_iteration(cudaStream_t st, cudaEvent_t* ev = nullptr) {
        if (gd2) return gd2_iteration(st, ev);
        const GdK k = gd_k();
        int rc;
        LPC_EV(0);
        // convolve(x): pad is implicit (x lives in xpad's window)
        LPC_CF(FftTraits<T>::fwd(fwd, xpad, spec));
        LPC_EV(1);
        if ((rc = spec_mul(spec, Np, 0, st))) return rc;
        LPC_EV(2);
        LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
        LPC_EV(3);
        dim3 grid((w + 127) / 128, h, Np);
        k_gd_residual<T><<<grid, 128, 0, st>>>(k, rpad, data, dpad);
        LPC_KCHECK();
        LPC_EV(4);
        // deconvolve(residual)
        LPC_CF(FftTraits<T>::fwd(fwd, dpad, spec));
        LPC_EV(5);
        if ((rc = spec_mul(spec, Np, 1, st))) return rc;
        LPC_EV(6);
        LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
        LPC_EV(7);
        if (algo == LPC_ALGO_GD) {
            k_gd_update<T><<<grid, 128, 0, st>>>(k, rpad, consts, xpad);
        } else if (algo == LPC_ALGO_NESTEROV) {
            k_nesterov_update<T><<<grid, 128, 0, st>>>(k, rpad, consts, xpad, mom, (T)nmu);
        } else {
            // gd.py:238-240: t_{k+1} = (1 + sqrt(1 + 4 t_k^2))/2 ; beta = (t_k - 1)/t_{k+1}
            const FistaStep f = fista_step();
            k_fista_update<T><<<grid, 128, 0, st>>>(k, rpad, f.consts, xpad, mom, f.beta, f.first);
        }
        LPC_KCHECK();
        LPC_EV(8);
        return LPC_OK;
    }
#undef LPC_EV

    // per-phase device times of the un-graphed iteration, CUDA events on the launching stream
    int profile(int n_iter, cudaStream_t st, double* ms, int* n_phases) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "profile", "convolver plans do not iterate");
        if (!psf_set || !data_set || !state_ready) return fail(LPC_ESTATE, "profile", "PSF, data and reset first");
        LPC_CK(cudaSetDevice(device));
        if (have_fwd) LPC_CF(cufftSetStream(fwd, st));
        if (have_inv) LPC_CF(cufftSetStream(inv, st));
        const int np = (algo == LPC_ALGO_ADMM) ? (v2 ? 3 : 4) : (gd2 ? 4 : 8);
        cudaEvent_t ev[9];
        for (int i = 0; i <= np; ++i) LPC_CK(cudaEventCreate(&ev[i]));
        for (int i = 0; i < np; ++i) ms[i] = 0.0;
        int rc = LPC_OK;
        for (int it = 0; it < n_iter && rc == LPC_OK; ++it) {
            rc = (algo == LPC_ALGO_ADMM) ? admm_iteration(st, ev) : gd_iteration(st, ev);
            if (rc != LPC_OK) break;
            LPC_CK(cudaEventSynchronize(ev[np]));
            for (int i = 0; i < np; ++i) {
                float t = 0.f;
                LPC_CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
                ms[i] += (double)t;
            }
        }
        for (int i = 0; i <= np; ++i) cudaEventDestroy(ev[i]);
        if (rc == LPC_OK && n_iter > 0) for (int i = 0; i < np; ++i) ms[i] /= n_iter;
        *n_phases = np;
        return rc;
    }

    // plane range of chain c
    void chain_range(int c, int* p0, int* np) const {
        const int base = Np / n_chains, extra = Np % n_chains;
        *p0 = c * base + (c < extra ? c : extra);
        *np = base + (c < extra ? 1 : 0);
    }
    // The planes of a plan are independent problems, so the fused engine runs them as
    // n_chains concurrent kernel chains: the DRAM-bound spatial kernel of one plane overlaps with the
    // latency-bound column kernel of another.  Forked from / joined to `st` with events (also under
    // stream capture, where this becomes a forked CUDA graph).
    int iterate_chains(int n_iter, cudaStream_t st) {
        int rc = LPC_OK;
        LPC_CK(cudaEventRecord(ev_fork, st));
        for (int c = 1; c < n_chains; ++c) LPC_CK(cudaStreamWaitEvent(side[c], ev_fork, 0));
        for (int it = 0; it < n_iter; ++it) {
            T scalar = T(0);
            const T* cst = nullptr;
            int first = 0, upd_mode = LPC_GD_UPD_GD;
            if (gd2) {
                if (algo == LPC_ALGO_NESTEROV) { scalar = (T)nmu; upd_mode = LPC_GD_UPD_NESTEROV; }
                if (algo == LPC_ALGO_FISTA) {
                    const FistaStep f = fista_step();
                    scalar = f.beta; first = f.first; cst = f.consts;
                    upd_mode = LPC_GD_UPD_FISTA;
                }
            }
            for (int c = 0; c < n_chains; ++c) {
                cudaStream_t cs = c == 0 ? st : side[c];
                int p0, np;
                chain_range(c, &p0, &np);
                if (v2) {
                    if (!vh_valid && (rc = launch_rows_ifft(cs, p0, np))) return rc;
                    if ((rc = launch_rows_spatial(cs, eta[eta_cur], eta[eta_cur ^ 1], p0, np, nullptr, it))) return rc;
                    if ((rc = launch_cols(cs, p0, np, it))) return rc;
                } else {
                    if ((rc = gd2_cols(0, cs, p0, np))) return rc;
                    if ((rc = gd2_rows(LPC_GD_RESID, T(0), 0, cs, p0, np))) return rc;
                    if ((rc = gd2_cols(1, cs, p0, np))) return rc;
                    if ((rc = gd2_rows(upd_mode, scalar, first, cs, p0, np, cst))) return rc;
                }
            }
            if (v2) { eta_cur ^= 1; vh_valid = false; admm_advance(it); }
        }
        for (int c = 1; c < n_chains; ++c) {
            LPC_CK(cudaEventRecord(ev_join[c], side[c]));
            LPC_CK(cudaStreamWaitEvent(st, ev_join[c], 0));
        }
        return LPC_OK;
    }

    // Plane-major order (fused ADMM engine, EXPERIMENT, off by default): the planes of a plan are independent
    // problems, so plane p could run ALL its iterations before plane p+1 starts; every producer -> consumer hand-off
    // (A -> B1 through Tin, B1 -> B2 through V/HV, B2 -> A through Tout) is then 66 MB written and immediately read
    // back, inside the 126 MB L2, and V/HV of every plane share ONE scratch plane pair.  MEASURED slower on B200
    // (0.533 vs 0.490 ms per iteration): the per-plane kernels lose ~13 % to partial waves and pipeline prologues
    // (A 3 x 58 us vs 142 us), the L2 hand-off wins back ~4 %.  LPC_PLANE_MAJOR=1 enables it.
    int plane_major = 0, vh_shared = 1;
    int iterate_plane_major(int n_iter, cudaStream_t st) {
        int rc = LPC_OK;
        const int eta0 = eta_cur;
        const double q1 = pm1, q2 = pm2, q3 = pm3;
        for (int p = 0; p < Np; ++p) {
            T* vh = vh_shared ? VH - (int64_t)p * plane : VH;  // plane p of this view is plane 0 of VH (V) / of its HV half
            pm1 = q1; pm2 = q2; pm3 = q3;
            for (int it = 0; it < n_iter; ++it) {
                const int ei = (eta0 + it) & 1;
                if ((rc = launch_rows_ifft(st, p, 1, vh))) return rc;
                if ((rc = launch_rows_spatial(st, eta[ei], eta[ei ^ 1], p, 1, vh, it))) return rc;
                if ((rc = launch_cols(st, p, 1, it))) return rc;
                admm_advance(it);
            }
        }
        eta_cur = (eta0 + n_iter) & 1;
        vh_valid = false;
        return LPC_OK;
    }

    int iterate(int n_iter, cudaStream_t st) {
        if (v2 && plane_major && n_chains == 1) return iterate_plane_major(n_iter, st);
        if ((v2 || gd2) && n_chains > 1) return iterate_chains(n_iter, st);
        int rc = LPC_OK;
        for (int i = 0; i < n_iter && rc == LPC_OK; ++i) rc = (algo == LPC_ALGO_ADMM) ? admm_iteration(st, nullptr, i) : gd_iteration(st);
        return rc;
    }

    // ---------------------------------------------------------------- plug-and-play halves (lpc_pnp.cuh)
    int pnp_begin(void* dev_out, cudaStream_t st) override {
        if (!pnp) return fail(LPC_ESTATE, "pnp_begin", "not a plug-and-play plan (LPC_FLAG_PNP)");
        if (!psf_set || !data_set || !state_ready) return fail(LPC_ESTATE, "pnp_begin", "PSF, data and reset first");
        if (pnp_open) return fail(LPC_ESTATE, "pnp_begin", "the previous iteration was not finished");
        if (!dev_out) return fail(LPC_EINVAL, "pnp_begin", "null output");
        if (algo == LPC_ALGO_ADMM) {
            dim3 grid((Wp + 127) / 128, Hp, Np);
            k_pnp_admm_pre<T><<<grid, 128, 0, st>>>(admm_k(), VH, pnp_U, eta[0], (T*)dev_out, pnp_use_dual ? 1 : 0);
            LPC_KCHECK();
        } else {
            // gradient (gd.py:128-130) on the cuFFT engine, then the step up to the projection
            const GdK k = gd_k();
            int rc;
            LPC_CF(cufftSetStream(fwd, st));
            LPC_CF(cufftSetStream(inv, st));
            LPC_CF(FftTraits<T>::fwd(fwd, xpad, spec));
            if ((rc = spec_mul(spec, Np, 0, st))) return rc;
            LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
            dim3 grid((w + 127) / 128, h, Np);
            k_gd_residual<T><<<grid, 128, 0, st>>>(k, rpad, data, dpad);
            LPC_CF(FftTraits<T>::fwd(fwd, dpad, spec));
            if ((rc = spec_mul(spec, Np, 1, st))) return rc;
            LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
            T* momp = mom ? mom : xpad;  // plain GD has no momentum buffer (never dereferenced for it)
            k_pnp_gd_step<T><<<grid, 128, 0, st>>>(k, algo, rpad, consts, xpad, momp, (T)nmu, fista_first ? 1 : 0, (T*)dev_out);
            LPC_KCHECK();
        }
        pnp_open = true;
        return LPC_OK;
    }
    int pnp_finish(const void* dev_den, cudaStream_t st) override {
        if (!pnp) return fail(LPC_ESTATE, "pnp_finish", "not a plug-and-play plan (LPC_FLAG_PNP)");
        if (!pnp_open) return fail(LPC_ESTATE, "pnp_finish", "lpc_pnp_begin first");
        if (!dev_den) return fail(LPC_EINVAL, "pnp_finish", "null input");
        pnp_open = false;
        if (algo == LPC_ALGO_ADMM) {
            const AdmmK<T> k = admm_k();
            T* V = VH; T* HV = VH + (int64_t)Np * plane;
            T* a = AX; T* nbq = AX + (int64_t)Np * plane;
            dim3 grid((Wp + 127) / 128, Hp, Np);
            k_pnp_admm_mid<T><<<grid, 128, 0, st>>>(k, (const T*)dev_den, V, HV, data, pnp_xi, rho, eta[0], pnp_U, pnp_X, pnp_W, a, nbq,
                                                     pnp_use_dual ? 1 : 0);
            LPC_KCHECK();
            LPC_CF(cufftSetStream(fwd, st));
            LPC_CF(cufftSetStream(inv, st));
            LPC_CF(FftTraits<T>::fwd(fwd, AX, spec));
            dim3 g2((unsigned)((splane + 255) / 256), Np);
            k_admm_spectral<T><<<g2, 256, 0, st>>>(spec, Hc, Rn, Sy, Sx, Np, PQ, Hp, Nh);
            LPC_KCHECK();
            LPC_CF(FftTraits<T>::inv(inv, spec, VH));
            k_pnp_admm_post<T><<<grid, 128, 0, st>>>(k, V, HV, pnp_X, pnp_W, pnp_U, pnp_xi, rho, eta[0], pnp_use_dual ? 1 : 0);
            LPC_KCHECK();
        } else {
            T beta = T(0);
            if (algo == LPC_ALGO_FISTA) {
                const double tn = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;  // gd.py:238
                beta = (T)((tk - 1.0) / tn);
                tk = tn;
                fista_first = false;
            }
            dim3 grid((w + 127) / 128, h, Np);
            T* momp = mom ? mom : xpad;
            k_pnp_gd_finish<T><<<grid, 128, 0, st>>>(gd_k(), algo, (const T*)dev_den, xpad, momp, beta);
            LPC_KCHECK();
        }
        return LPC_OK;
    }

    int run(int n_iter, cudaStream_t st) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "run", "convolver plans do not iterate");
        if (pnp) return fail(LPC_ESTATE, "run", "plug-and-play plans iterate through lpc_pnp_begin / lpc_pnp_finish");
        if (!psf_set || !data_set) return fail(LPC_ESTATE, "run", "PSF and data must be set first");
        if (!state_ready) return fail(LPC_ESTATE, "run", "call lpc_reset first");
        if (n_iter < 0) return fail(LPC_EINVAL, "run", "n_iter < 0");
        if (n_iter == 0) return LPC_OK;
        if (n_sched > 0 && algo == LPC_ALGO_ADMM && n_iter > n_sched)
            return fail(LPC_EINVAL, "run", "n_iter exceeds the unrolled schedule (the schedule restarts at 0 with every run)");
        if (n_sched > 0 && algo == LPC_ALGO_FISTA && gd_it + n_iter > n_sched)
            return fail(LPC_EINVAL, "run", "iterations since reset exceed the unrolled schedule");
        LPC_CK(cudaSetDevice(device));
        if (have_fwd) LPC_CF(cufftSetStream(fwd, st));
        if (have_inv) LPC_CF(cufftSetStream(inv, st));
        static const bool env_no_graph = (getenv("LPC_NO_GRAPH") != nullptr);
        const bool use_graph = !(flags & LPC_FLAG_NO_GRAPH) && !env_no_graph && st != nullptr;
        if (!use_graph) return iterate(n_iter, st);

        const int first = fista_first ? 1 : 0;
        const int vh0 = vh_valid ? 1 : 0;  // the first captured iteration skips B1 when VH already holds V, HV
        GraphEntry* hit = nullptr;
        for (GraphEntry& g : graphs) {
            if (g.exec && g.niter == n_iter && g.eta == eta_cur && g.stream == st && g.vh == vh0 &&
                (algo != LPC_ALGO_FISTA || (g.tk == tk && g.first == first && g.gdit == gd_it)) &&
                (algo != LPC_ALGO_ADMM || (g.pm[0] == pm1 && g.pm[1] == pm2 && g.pm[2] == pm3))) { hit = &g; break; }
        }
        if (!hit) {
            if ((int)graphs.size() >= MAX_GRAPHS) {  // evict the least recently used entry
                size_t lru = 0;
                for (size_t i = 1; i < graphs.size(); ++i) if (graphs[i].stamp < graphs[lru].stamp) lru = i;
                if (graphs[lru].exec) cudaGraphExecDestroy(graphs[lru].exec);
                graphs.erase(graphs.begin() + lru);
            }
            const int eta0 = eta_cur, gdit0 = gd_it;
            const double tk_start = tk, pm0[3] = {pm1, pm2, pm3};
            cudaGraph_t graph = nullptr;
            LPC_CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = iterate(n_iter, st);
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc != LPC_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
            if (e != cudaSuccess) return fail(LPC_ECUDA, "cudaStreamEndCapture", cudaGetErrorString(e));
            GraphEntry g;
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(LPC_ECUDA, "cudaGraphInstantiate", cudaGetErrorString(e));
            g.niter = n_iter; g.eta = eta0; g.tk = tk_start; g.first = first; g.stream = st; g.vh = vh0;
            g.gdit = gdit0; g.pm[0] = pm0[0]; g.pm[1] = pm0[1]; g.pm[2] = pm0[2];
            graphs.push_back(g);
            hit = &graphs.back();
            // iterate() advanced the host-side bookkeeping while capturing; nothing ran yet
        } else {
            // replay: advance the host-side bookkeeping exactly as iterate() would
            if (algo == LPC_ALGO_ADMM) { eta_cur ^= (n_iter & 1); admm_advance(n_iter - 1); }
            if (algo == LPC_ALGO_FISTA) {
                if (n_sched > 0) gd_it += n_iter;
                else for (int i = 0; i < n_iter; ++i) tk = (1.0 + std::sqrt(1.0 + 4.0 * tk * tk)) / 2.0;
                fista_first = false;
            }
        }
        hit->stamp = ++graph_clock;
        LPC_CK(cudaGraphLaunch(hit->exec, st));
        vh_valid = false;  // V / HV in VH are stale as soon as an iteration has run (engine v2 keeps only spectra)
        return LPC_OK;
    }

    // ---------------------------------------------------------------- outputs
    int get_image(void* dev_out, cudaStream_t st) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "get_image", "convolver plans have no image");
        if (!state_ready) return fail(LPC_ESTATE, "get_image", "no state: call lpc_reset");
        if (!dev_out) return fail(LPC_EINVAL, "get_image", "null output");
        LPC_CK(cudaSetDevice(device));
        if (algo == LPC_ALGO_ADMM) { int rc = materialise_vh(st); if (rc) return rc; }
        if (gd2) return unpack(xu, (int64_t)h * w, w, 0, 0, (T*)dev_out, B * D, h, w, 1, st);
        const T* src = (algo == LPC_ALGO_ADMM) ? VH : xpad;
        return unpack(src, plane, Wp, sy, sx, (T*)dev_out, B * D, h, w, 1, st);
    }
    int get_image_host(void* host_out, cudaStream_t st) override {
        if (!host_out) return fail(LPC_EINVAL, "get_image_host", "null output");
        size_t n = (size_t)B * D * h * w * C;
        int rc = ensure_stage(n);
        if (rc) return rc;
        if ((rc = get_image(stage, st))) return rc;
        LPC_CK(cudaMemcpyAsync(host_out, stage, n * sizeof(T), cudaMemcpyDeviceToHost, st));
        return LPC_OK;
    }
    int get_state(int which, void* dev_out, cudaStream_t st) override {
        if (!state_ready) return fail(LPC_ESTATE, "get_state", "no state: call lpc_reset");
        if (!dev_out) return fail(LPC_EINVAL, "get_state", "null output");
        LPC_CK(cudaSetDevice(device));
        if (algo == LPC_ALGO_ADMM) {
            if (which == LPC_STATE_IMAGE_EST || which == LPC_STATE_FORWARD) { int rc = materialise_vh(st); if (rc) return rc; }
            const T* src = nullptr;
            switch (which) {
                case LPC_STATE_IMAGE_EST: src = VH; break;
                case LPC_STATE_FORWARD: src = VH + (int64_t)Np * plane; break;
                case LPC_STATE_XI: src = AX + (int64_t)Np * plane; break;
                case LPC_STATE_RHO: src = rho; break;
                case LPC_STATE_ETA0: src = eta[eta_cur]; break;
                case LPC_STATE_ETA1: src = eta[eta_cur] + (int64_t)Np * plane; break;
                default: return fail(LPC_EINVAL, "get_state", "unknown selector");
            }
            return unpack(src, plane, Wp, 0, 0, (T*)dev_out, B * D, Hp, Wp, 0, st);
        }
        if (algo >= LPC_ALGO_GD && which == LPC_STATE_IMAGE_EST) {
            if (gd2) return unpack(xu, (int64_t)h * w, w, 0, 0, (T*)dev_out, B * D, h, w, 0, st);
            return unpack(xpad, plane, Wp, sy, sx, (T*)dev_out, B * D, h, w, 0, st);
        }
        return fail(LPC_EINVAL, "get_state", "selector not available for this algorithm");
    }
    int get_H(void* dev_out, cudaStream_t st) override {
        if (!psf_set) return fail(LPC_ESTATE, "get_H", "PSF not set");
        LPC_CK(cudaSetDevice(device));
        k_unpack_complex<T><<<1184, 256, 0, st>>>(Hc, (Cplx*)dev_out, n_psf * D, C, Hp, Nh);
        LPC_KCHECK();
        return LPC_OK;
    }
    int get_alpha(void* dev_out, cudaStream_t st) override {
        if (!psf_set) return fail(LPC_ESTATE, "get_alpha", "PSF not set");
        LPC_CK(cudaSetDevice(device));
        LPC_CK(cudaMemcpyAsync(dev_out, consts + 3 * NC, NC * sizeof(T), cudaMemcpyDeviceToDevice, st));
        return LPC_OK;
    }

    // ---------------------------------------------------------------- bare operator
    int convolve(const void* dev_in, const int64_t* in_strides, void* dev_out, int adjoint, int pad,
                 cudaStream_t st) override {
        if (algo != LPC_ALGO_CONV) return fail(LPC_ESTATE, "convolve", "not a convolver plan");
        if (!psf_set) return fail(LPC_ESTATE, "convolve", "PSF not set");
        if (!dev_in || !dev_out) return fail(LPC_EINVAL, "convolve", "null pointer");
        LPC_CK(cudaSetDevice(device));
        int rc;
        const int hh = pad ? h : Hp, ww = pad ? w : Wp, oy = pad ? sy : 0, ox = pad ? sx : 0;
        if (pad && xpad_border_dirty) {
            LPC_CK(cudaMemsetAsync(xpad, 0, (size_t)Np * plane * sizeof(T), st));
            xpad_border_dirty = false;
        }
        if (!pad) xpad_border_dirty = true;
        Strides5 ss = in_strides ? from_array(in_strides) : contiguous(D, hh, ww, C);
        if ((rc = pack((const T*)dev_in, ss, nullptr, Strides5{}, xpad, B, D, hh, ww, plane, Wp, oy, ox, st))) return rc;
        LPC_CF(cufftSetStream(fwd, st));
        LPC_CF(cufftSetStream(inv, st));
        LPC_CF(FftTraits<T>::fwd(fwd, xpad, spec));
        if ((rc = spec_mul(spec, Np, adjoint, st))) return rc;
        LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
        return unpack(rpad, plane, Wp, oy, ox, (T*)dev_out, B * D, hh, ww, 0, st);
    }

    // reconstruction_error (recon.py:607-653): convolve the prediction (ALWAYS through pad -> convolve -> crop: with
    // pad=False the reference pads the prediction itself and crops the result, :633-638), then min/max-normalise
    // per (batch, depth) and reduce the squared error against the measurement per batch entry.  dev_out: T[B].
    T* err_mm = nullptr;
    double* err_acc = nullptr;
    int recon_error(const void* dev_pred, const void* dev_lensless, int Dl, int normalize, double npix, void* dev_out,
                    cudaStream_t st) override {
        if (algo != LPC_ALGO_CONV) return fail(LPC_ESTATE, "reconstruction_error", "not a convolver plan");
        if (!psf_set) return fail(LPC_ESTATE, "reconstruction_error", "PSF not set");
        if (!dev_pred || !dev_lensless || !dev_out) return fail(LPC_EINVAL, "reconstruction_error", "null pointer");
        if (Dl != 1 && Dl != D) return fail(LPC_EINVAL, "reconstruction_error", "measurement depth must be 1 or the PSF depth");
        if (!(npix > 0)) return fail(LPC_EINVAL, "reconstruction_error", "npix must be > 0");
        int rc;
        if (!err_mm) {
            if ((rc = alloc(&err_mm, (size_t)2 * B * D))) return rc;
            if ((rc = alloc(&err_acc, (size_t)B))) return rc;
        }
        if (xpad_border_dirty) {
            LPC_CK(cudaMemsetAsync(xpad, 0, (size_t)Np * plane * sizeof(T), st));
            xpad_border_dirty = false;
        }
        if ((rc = pack((const T*)dev_pred, contiguous(D, h, w, C), nullptr, Strides5{}, xpad, B, D, h, w, plane, Wp, sy, sx, st))) return rc;
        LPC_CF(cufftSetStream(fwd, st));
        LPC_CF(cufftSetStream(inv, st));
        LPC_CF(FftTraits<T>::fwd(fwd, xpad, spec));
        if ((rc = spec_mul(spec, Np, 0, st))) return rc;
        LPC_CF(FftTraits<T>::inv(inv, spec, rpad));
        const int G = B * D;
        k_err_init<T><<<(G + B + 127) / 128, 128, 0, st>>>(err_mm, G, err_acc, B);
        dim3 grid(64, G);
        if (normalize) k_err_minmax<T><<<grid, 256, 0, st>>>(rpad, plane, Wp, sy, sx, C, h, w, err_mm);
        k_err_sse<T><<<grid, 256, 0, st>>>(rpad, plane, Wp, sy, sx, C, D, Dl, h, w, err_mm, normalize, (const T*)dev_lensless, err_acc);
        k_err_final<T><<<(B + 127) / 128, 128, 0, st>>>(err_acc, (T*)dev_out, B, 1.0 / npix);
        LPC_KCHECK();
        return LPC_OK;
    }

    // set_image_estimate (recon.py:415-442): overwrite the current iterate, leave everything else alone.
    // GD family: x <- est; Nesterov's p, FISTA's t_k and previous iterate stay.  Right after a reset FISTA's
    // previous iterate ALIASES the estimate (gd.py:233); assigning a new tensor breaks the alias, so the old
    // estimate is what the next update's momentum term sees.
    // ADMM: implemented as a warm start (V <- est, HV, Psi V recomputed, duals zeroed) -- DEVIATION: the reference
    // keeps its duals and, for one iteration, the stale HV / Psi V of the previous iterate.
    int set_image_estimate(const void* dev_est, cudaStream_t st) override {
        if (algo == LPC_ALGO_CONV) return fail(LPC_ESTATE, "set_image_estimate", "convolver plans have no solver state");
        if (!dev_est) return fail(LPC_EINVAL, "set_image_estimate", "null estimate");
        if (algo == LPC_ALGO_ADMM || !state_ready) return reset(dev_est, st);
        int rc;
        T* xdst = gd2 ? xu : xpad;
        if (algo == LPC_ALGO_FISTA && fista_first && n_sched == 0) {
            if (gd2) LPC_CK(cudaMemcpyAsync(mom, xu, (size_t)Np * h * w * sizeof(T), cudaMemcpyDeviceToDevice, st));
            else if ((rc = copy_window(xpad, mom, st))) return rc;
            fista_first = false;
        }
        const int64_t xps = gd2 ? (int64_t)h * w : plane;
        const int xpitch = gd2 ? w : Wp, xoy = gd2 ? 0 : sy, xox = gd2 ? 0 : sx;
        if ((rc = pack((const T*)dev_est, contiguous(D, h, w, C), nullptr, Strides5{}, xdst, B, D, h, w, xps, xpitch, xoy, xox, st))) return rc;
        if (gd2 && (rc = gd2_rows(LPC_GD_INIT, T(0), 0, st))) return rc;  // row spectra of the new estimate
        return LPC_OK;
    }
    // window of padded planar planes -> unpadded planar [Np][h][w]
    int copy_window(const T* src, T* dst, cudaStream_t st) {
        LPC_CK(cudaMemcpy2DAsync(dst, (size_t)w * sizeof(T), src + (int64_t)sy * Wp + sx, (size_t)Wp * sizeof(T), (size_t)w * sizeof(T), (size_t)h,
                                 cudaMemcpyDeviceToDevice, st));
        for (int p = 1; p < Np; ++p)
            LPC_CK(cudaMemcpy2DAsync(dst + (int64_t)p * h * w, (size_t)w * sizeof(T), src + (int64_t)p * plane + (int64_t)sy * Wp + sx,
                                     (size_t)Wp * sizeof(T), (size_t)w * sizeof(T), (size_t)h, cudaMemcpyDeviceToDevice, st));
        return LPC_OK;
    }

    int set_option(int option, double value) override {
        switch (option) {
            case LPC_OPT_KEEP_X0: keep_x0 = value != 0.0; return LPC_OK;
            case LPC_OPT_PNP_USE_DUAL: pnp_use_dual = value != 0.0; return LPC_OK;
            case LPC_OPT_B2_PREFETCH: b2_prefetch = (int)value; drop_graph(); return LPC_OK;
            default: return fail(LPC_EINVAL, "set_option", "unknown option");
        }
    }

    // ---------------------------------------------------------------- introspection
    int bytes_per_iter(double* out) override {
        const double P = (double)sizeof(T) * Np * (double)plane;
        const double Sp = 2.0 * sizeof(T) * Np * (double)splane;
        const double d = (double)sizeof(T) * Np * (double)h * w;
        if (algo == LPC_ALGO_ADMM) *out = 16.0 * P + 9.5 * Sp + d;
        else if (algo >= LPC_ALGO_GD) *out = 10.0 * d + 6.0 * Sp;
        else *out = 2.0 * d + 2.0 * Sp;
        return LPC_OK;
    }
    int launches_per_iter(int* total, int* own) override {
        int o = 0, f = 0;
        if (algo == LPC_ALGO_ADMM) { o = v2 ? 3 : 2; f = v2 ? 0 : 2; }
        else if (algo >= LPC_ALGO_GD) { o = 4; f = gd2 ? 0 : 4; }
        else { o = 3; f = 2; }
        if (own) *own = o;
        if (total) *total = o + f;  // each cuFFT exec is at least one kernel
        return LPC_OK;
    }
    size_t plan_bytes() override { return total_bytes + work_bytes; }
    int engine_version() override { return (v2 || gd2) ? 2 : 1; }
    const char* phase_name(int phase) override {
        static const char* admm1[] = {"admm_spatial(K1)", "cufft_r2c", "admm_spectral(K2)", "cufft_c2r"};
        static const char* admm2[] = {"rows_ifft(B1)", "rows_spatial_fft(B2)", "admm_cols(A)"};
        static const char* gd[] = {"cufft_r2c", "spec_mul_H", "cufft_c2r", "residual", "cufft_r2c", "spec_mul_Hadj", "cufft_c2r", "update"};
        static const char* gdf[] = {"gd_cols(H)", "gd_rows(residual)", "gd_cols(Hadj)", "gd_rows(update)"};
        if (algo >= LPC_ALGO_GD && gd2 && phase >= 0 && phase < 4) return gdf[phase];
        if (algo == LPC_ALGO_ADMM && v2 && phase >= 0 && phase < 3) return admm2[phase];
        if (algo == LPC_ALGO_ADMM && !v2 && phase >= 0 && phase < 4) return admm1[phase];
        if (algo >= LPC_ALGO_GD && phase >= 0 && phase < 8) return gd[phase];
        return "";
    }
};

}  // namespace

// ---- input side on the GPU (lpc_io.cuh; io.py:20-207, 210-388)
namespace {
// weights of one antialiased axis (ATen UpSampleKernel.cpp: _compute_indices_min_size_weights_aa with the triangle
// filter, float arithmetic as for float32 images); returns max_w
int aa_axis_weights(int n_in, int n_out, std::vector<int2>* idx, std::vector<float>* wts) {
    // expression types as in ATen (scalar_t = float for float32 images; the 0.5 / 1.0 literals are doubles there too)
    const float scale = (float)n_in / (float)n_out;  // area_pixel_compute_scale, align_corners = false, no scale factor
    const float support = (scale >= 1.0) ? (float)((2 * 0.5) * scale) : (float)(2 * 0.5);
    const int max_w = (int)std::ceil(support) * 2 + 1;
    idx->assign(n_out, int2{0, 0});
    wts->assign((size_t)n_out * max_w, 0.0f);
    const float invscale = (scale >= 1.0) ? (float)(1.0 / scale) : 1.0f;
    for (int i = 0; i < n_out; ++i) {
        const float center = (float)(scale * (i + 0.5));
        int64_t xmin = std::max((int64_t)(center - support + 0.5), (int64_t)0);
        int64_t xsize = std::min((int64_t)(center + support + 0.5), (int64_t)n_in) - xmin;
        xsize = std::min(std::max(xsize, (int64_t)0), (int64_t)max_w);
        float total = 0.0f;
        float* w = wts->data() + (size_t)i * max_w;
        for (int64_t j = 0; j < xsize; ++j) {
            float x = (float)((j + xmin - center + 0.5) * invscale);
            x = std::fabs(x);
            w[j] = x < 1.0f ? 1.0f - x : 0.0f;
            total += w[j];
        }
        if (total != 0.0f) for (int64_t j = 0; j < xsize; ++j) w[j] /= total;
        (*idx)[i] = int2{(int)xmin, (int)xsize};
    }
    return max_w;
}
struct IoScratch {  // stream-ordered scratch, freed on the same stream
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit IoScratch(cudaStream_t s) : st(s) {}
    template <typename U> U* get(size_t n) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, n * sizeof(U) + 16, st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (U*)p;
    }
    ~IoScratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
};
// [D,H,W,C] -> [D,h,w,C]: width pass then height pass (ATen separable_upsample_generic_Nd_kernel_impl)
int io_resize(IoScratch& sc, const float* in, float* out, int D, int H, int W, int C, int h, int w, cudaStream_t st) {
    const float* cur = in;
    int curW = W;
    if (w != W) {
        std::vector<int2> idx; std::vector<float> wts;
        const int mw = aa_axis_weights(W, w, &idx, &wts);
        int2* didx = sc.get<int2>(idx.size()); float* dw = sc.get<float>(wts.size());
        float* tmp = (h == H) ? out : sc.get<float>((size_t)D * H * w * C);
        if (!didx || !dw || !tmp) return LPC_ENOMEM;
        cudaMemcpyAsync(didx, idx.data(), idx.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(dw, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);  // idx / wts are pageable host temporaries
        const int64_t n = (int64_t)D * H * w * C;
        k_io_resize_axis<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 32), 256, 0, st>>>(cur, tmp, (int64_t)D * H, W, w, C, didx, dw, mw);
        cur = tmp; curW = w;
    }
    if (h != H) {
        std::vector<int2> idx; std::vector<float> wts;
        const int mw = aa_axis_weights(H, h, &idx, &wts);
        int2* didx = sc.get<int2>(idx.size()); float* dw = sc.get<float>(wts.size());
        if (!didx || !dw) return LPC_ENOMEM;
        cudaMemcpyAsync(didx, idx.data(), idx.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(dw, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        const int64_t n = (int64_t)D * h * curW * C;
        k_io_resize_axis<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 32), 256, 0, st>>>(cur, out, D, H, h, (int64_t)curW * C, didx, dw, mw);
        cur = out;
    }
    if (cur != out) cudaMemcpyAsync(out, cur, (size_t)D * h * w * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return cudaGetLastError() == cudaSuccess ? LPC_OK : LPC_ECUDA;
}
template <typename RAW>
int io_load(int is_psf, const RAW* raw, int H, int W, int C, int flags, int bg0, int bg1, const float* dev_bg_in, int out_h, int out_w,
            int normalize, float* dev_out, float* dev_bg_out, cudaStream_t st) {
    IoScratch sc(st);
    const int64_t n = (int64_t)H * W * C;
    float* f = sc.get<float>((size_t)n);
    unsigned int* rmax = sc.get<unsigned int>(4);
    float* maxval = sc.get<float>(4);
    double* sumsq = sc.get<double>(2);
    if (!f || !rmax || !maxval || !sumsq) return LPC_ENOMEM;
    cudaMemsetAsync(rmax, 0, 16, st);
    cudaMemsetAsync(sumsq, 0, 16, st);
    const unsigned nb = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 32);
    k_io_raw_to_float<RAW><<<nb, 256, 0, st>>>(raw, f, H, W, C, (flags & LPC_IO_BGR) ? 1 : 0, (flags & LPC_IO_FLIP_UD) ? 1 : 0,
                                              (flags & LPC_IO_FLIP_LR) ? 1 : 0, rmax);
    k_io_max_val<<<1, 32, 0, st>>>(rmax, maxval);
    if (is_psf) {
        if (bg1 > bg0) {  // io.py:343-357
            if (bg1 > H || bg1 > W || bg0 < 0 || !dev_bg_out) return LPC_EINVAL;
            k_io_region_mean<<<C, 256, 0, st>>>(f, 1, H, W, C, bg0, bg1, dev_bg_out);
            k_io_sub_clip<<<nb, 256, 0, st>>>(f, n, C, dev_bg_out, nullptr);
        }
    } else if (dev_bg_in) {  // io.py:164-171: bg is the PSF's normalised background level, scaled back to counts
        k_io_sub_clip<<<nb, 256, 0, st>>>(f, n, C, dev_bg_in, maxval);
    }
    unsigned int* vmax = rmax + 1;
    const bool norm_first = !is_psf && normalize && (flags & LPC_IO_NORMALIZE_FIRST);
    if (norm_first) {  // load_data normalises the full-resolution measurement BEFORE resizing it to the PSF's shape (io.py:509-513,545-547)
        k_io_reduce<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(f, n, sumsq, vmax);
        k_io_normalize<<<nb, 256, 0, st>>>(f, n, sumsq, vmax, 1);
    }
    int Cc = C;
    float* src = f;
    const int64_t n_out = (int64_t)out_h * out_w * C;
    float* resized = (is_psf && (flags & LPC_IO_SINGLE_PSF) && C > 1) ? sc.get<float>((size_t)n_out) : dev_out;
    if (!resized) return LPC_ENOMEM;
    int rc = io_resize(sc, src, resized, 1, H, W, C, out_h, out_w, st);
    if (rc) return rc;
    float* fin = resized;
    if (is_psf && (flags & LPC_IO_SINGLE_PSF) && C > 1) {  // io.py:363-368
        k_io_sum_channels<<<nb, 256, 0, st>>>(resized, dev_out, (int64_t)out_h * out_w, C);
        fin = dev_out; Cc = 1;
    }
    const int64_t nf = (int64_t)out_h * out_w * Cc;
    if (is_psf || (normalize && !norm_first)) {
        k_io_reduce<<<(unsigned)std::min<int64_t>((nf + 255) / 256, 148 * 8), 256, 0, st>>>(fin, nf, sumsq, vmax);
        k_io_normalize<<<(unsigned)std::min<int64_t>((nf + 255) / 256, 148 * 32), 256, 0, st>>>(fin, nf, sumsq, vmax, is_psf ? 0 : 1);
    }
    if (is_psf && bg1 > bg0) k_io_scale_vec<<<1, 32, 0, st>>>(dev_bg_out, C, maxval);  // io.py:376: bg /= max_val
    return cudaGetLastError() == cudaSuccess ? LPC_OK : LPC_ECUDA;
}
}  // namespace

// ====================================================================================
// C-ABI
// ====================================================================================
extern "C" {

int lpc_version(void) { return LPC_VERSION; }

const char* lpc_error_name(int code) {
    switch (code) {
        case LPC_OK: return "LPC_OK";
        case LPC_EINVAL: return "LPC_EINVAL";
        case LPC_ECUDA: return "LPC_ECUDA";
        case LPC_ECUFFT: return "LPC_ECUFFT";
        case LPC_ESTATE: return "LPC_ESTATE";
        case LPC_ENOMEM: return "LPC_ENOMEM";
        case LPC_ENOTIMPL: return "LPC_ENOTIMPL";
        default: return "LPC_E?";
    }
}

const char* lpc_last_error(lpc_plan_t plan) { return plan ? plan->err.c_str() : g_create_error.c_str(); }

int lpc_next_fast_len(int n) { return next_fast_len_impl(n); }

int lpc_geometry(int h, int w, int* Hp, int* Wp, int* start_y, int* start_x) {
    if (h <= 0 || w <= 0) return LPC_EINVAL;
    const int hp = next_fast_len_impl(2 * h - 1), wp = next_fast_len_impl(2 * w - 1);
    if (Hp) *Hp = hp;
    if (Wp) *Wp = wp;
    if (start_y) *start_y = (hp - h) / 2;
    if (start_x) *start_x = (wp - w) / 2;
    return LPC_OK;
}

int lpc_plan_create(lpc_plan_t* out, int device, int algo, int dtype, int norm, int B, int D, int C, int h, int w,
                    int flags) {
    if (!out) return LPC_EINVAL;
    *out = nullptr;
    auto bad = [&](int code, const char* msg) { g_create_error = std::string(lpc_error_name(code)) + ": " + msg; return code; };
    if (algo < LPC_ALGO_CONV || algo > LPC_ALGO_FISTA) return bad(LPC_EINVAL, "unknown algorithm");
    if (dtype != LPC_F32 && dtype != LPC_F64) return bad(LPC_EINVAL, "dtype must be LPC_F32 or LPC_F64 (recon.py:264-290)");
    if (norm < LPC_NORM_BACKWARD || norm > LPC_NORM_FORWARD) return bad(LPC_EINVAL, "unknown norm");
    if (C != 1 && C != 3) return bad(LPC_EINVAL, "PSF must either be rgb (3) or grayscale (1) (recon.py:256)");
    if (B < 1 || D < 1 || h < 1 || w < 1) return bad(LPC_EINVAL, "B, D, h, w must be >= 1");
    if ((int64_t)B * D * C > 65535) return bad(LPC_EINVAL, "B*D*C must be <= 65535 planes per plan");
    lpc_plan_s* p = nullptr;
    if (dtype == LPC_F32) p = new (std::nothrow) Plan<float>();
    else p = new (std::nothrow) Plan<double>();
    if (!p) return bad(LPC_ENOMEM, "host allocation failed");
    p->device = device; p->algo = algo; p->dtype = dtype; p->norm = norm;
    p->B = B; p->D = D; p->C = C; p->h = h; p->w = w; p->flags = flags;
    DeviceGuard guard(device);
    int rc = p->init();
    if (rc != LPC_OK) {
        g_create_error = p->err;
        delete p;
        return rc;
    }
    *out = p;
    return LPC_OK;
}

int lpc_menu_query(int Wp, int Hp, int* has_rows, int* has_cols) {
    if (has_rows) *has_rows = find_rows(Wp) ? 1 : 0;
    if (has_cols) *has_cols = find_cols(Hp) ? 1 : 0;
    return LPC_OK;
}

int lpc_plan_destroy(lpc_plan_t plan) {
    if (!plan) return LPC_OK;
    DeviceGuard guard(plan->device);
    delete plan;
    return LPC_OK;
}

#define LPC_NEED(plan)                \
    if (!(plan)) return LPC_EINVAL;   \
    DeviceGuard lpc_guard_((plan)->device)

int lpc_set_psf(lpc_plan_t plan, const void* dev_psf, void* stream) {
    LPC_NEED(plan);
    return plan->set_psf(dev_psf, (cudaStream_t)stream);
}
int lpc_set_admm_params(lpc_plan_t plan, double mu1, double mu2, double mu3, double tau, void* stream) {
    LPC_NEED(plan);
    return plan->set_admm_params(mu1, mu2, mu3, tau, (cudaStream_t)stream);
}
int lpc_set_gd_params(lpc_plan_t plan, double lip_fact, double nesterov_mu, double fista_tk, void* stream) {
    LPC_NEED(plan);
    return plan->set_gd_params(lip_fact, nesterov_mu, fista_tk, (cudaStream_t)stream);
}
int lpc_set_admm_schedule(lpc_plan_t plan, int n, const double* mu1, const double* mu2, const double* mu3,
                          const double* tau, void* stream) {
    LPC_NEED(plan);
    return plan->set_admm_schedule(n, mu1, mu2, mu3, tau, (cudaStream_t)stream);
}
int lpc_set_fista_schedule(lpc_plan_t plan, int n, const double* alpha, const double* tk, void* stream) {
    LPC_NEED(plan);
    return plan->set_fista_schedule(n, alpha, tk, (cudaStream_t)stream);
}
int lpc_set_data(lpc_plan_t plan, const void* dev_data, int Dd, const int64_t* strides, const void* dev_background,
                 const int64_t* bg_strides, void* dev_echo, void* stream) {
    LPC_NEED(plan);
    return plan->set_data(dev_data, Dd, strides, dev_background, bg_strides, dev_echo, (cudaStream_t)stream);
}
int lpc_set_data_host(lpc_plan_t plan, const void* host_data, int Dd, void* stream) {
    LPC_NEED(plan);
    return plan->set_data_host(host_data, Dd, (cudaStream_t)stream);
}
int lpc_reset(lpc_plan_t plan, const void* dev_initial_est, void* stream) {
    LPC_NEED(plan);
    return plan->reset(dev_initial_est, (cudaStream_t)stream);
}
int lpc_run(lpc_plan_t plan, int n_iter, void* stream) {
    LPC_NEED(plan);
    return plan->run(n_iter, (cudaStream_t)stream);
}
int lpc_get_image(lpc_plan_t plan, void* dev_out, void* stream) {
    LPC_NEED(plan);
    return plan->get_image(dev_out, (cudaStream_t)stream);
}
int lpc_get_image_host(lpc_plan_t plan, void* host_out, void* stream) {
    LPC_NEED(plan);
    return plan->get_image_host(host_out, (cudaStream_t)stream);
}
int lpc_get_state(lpc_plan_t plan, int which, void* dev_out, void* stream) {
    LPC_NEED(plan);
    return plan->get_state(which, dev_out, (cudaStream_t)stream);
}
int lpc_get_H(lpc_plan_t plan, void* dev_out, void* stream) {
    LPC_NEED(plan);
    return plan->get_H(dev_out, (cudaStream_t)stream);
}
int lpc_get_alpha(lpc_plan_t plan, void* dev_out, void* stream) {
    LPC_NEED(plan);
    return plan->get_alpha(dev_out, (cudaStream_t)stream);
}
int lpc_convolve(lpc_plan_t plan, const void* dev_in, const int64_t* in_strides, void* dev_out, int adjoint, int pad,
                 void* stream) {
    LPC_NEED(plan);
    return plan->convolve(dev_in, in_strides, dev_out, adjoint, pad, (cudaStream_t)stream);
}
int lpc_reconstruction_error(lpc_plan_t plan, const void* dev_prediction, const void* dev_lensless, int Dl, int normalize,
                             double npix, void* dev_out, void* stream) {
    LPC_NEED(plan);
    return plan->recon_error(dev_prediction, dev_lensless, Dl, normalize, npix, dev_out, (cudaStream_t)stream);
}
int lpc_set_image_estimate(lpc_plan_t plan, const void* dev_est, void* stream) {
    LPC_NEED(plan);
    return plan->set_image_estimate(dev_est, (cudaStream_t)stream);
}
int lpc_pnp_begin(lpc_plan_t plan, void* dev_denoiser_in, void* stream) {
    LPC_NEED(plan);
    return plan->pnp_begin(dev_denoiser_in, (cudaStream_t)stream);
}
int lpc_pnp_finish(lpc_plan_t plan, const void* dev_denoised, void* stream) {
    LPC_NEED(plan);
    return plan->pnp_finish(dev_denoised, (cudaStream_t)stream);
}
int lpc_set_option(lpc_plan_t plan, int option, double value) {
    LPC_NEED(plan);
    return plan->set_option(option, value);
}
// ---- module-level helpers of admm.py as plan-less entry points (lpc_ops.cuh)
#define LPC_OP_PROLOGUE(devidx)                                    \
    if (dtype != LPC_F32 && dtype != LPC_F64) return LPC_EINVAL;   \
    DeviceGuard lpc_guard_(devidx);                                \
    cudaStream_t st = (cudaStream_t)stream
#define LPC_OP_EPILOGUE() return cudaGetLastError() == cudaSuccess ? LPC_OK : LPC_ECUDA
static inline unsigned op_blocks(int64_t n) { int64_t b = (n + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b)); }

int lpc_op_soft_thresh(int device, int dtype, const void* dev_x, void* dev_out, int64_t n, double thresh, void* stream) {
    if (!dev_x || !dev_out || n < 0) return LPC_EINVAL;
    LPC_OP_PROLOGUE(device);
    if (n == 0) return LPC_OK;
    if (dtype == LPC_F32) k_op_soft_thresh<float><<<op_blocks(n), 256, 0, st>>>((const float*)dev_x, (float*)dev_out, n, (float)thresh);
    else k_op_soft_thresh<double><<<op_blocks(n), 256, 0, st>>>((const double*)dev_x, (double*)dev_out, n, thresh);
    LPC_OP_EPILOGUE();
}
int lpc_op_finite_diff(int device, int dtype, const void* dev_x, void* dev_out, int64_t outer, int H, int W, int C, void* stream) {
    if (!dev_x || !dev_out || outer < 0 || H < 1 || W < 1 || C < 1) return LPC_EINVAL;
    LPC_OP_PROLOGUE(device);
    const int64_t n = outer * H * W * C;
    if (n == 0) return LPC_OK;
    if (dtype == LPC_F32) k_op_finite_diff<float><<<op_blocks(n), 256, 0, st>>>((const float*)dev_x, (float*)dev_out, outer, H, W, C);
    else k_op_finite_diff<double><<<op_blocks(n), 256, 0, st>>>((const double*)dev_x, (double*)dev_out, outer, H, W, C);
    LPC_OP_EPILOGUE();
}
int lpc_op_finite_diff_adj(int device, int dtype, const void* dev_u, void* dev_out, int64_t outer, int H, int W, int C, void* stream) {
    if (!dev_u || !dev_out || outer < 0 || H < 1 || W < 1 || C < 1) return LPC_EINVAL;
    LPC_OP_PROLOGUE(device);
    const int64_t n = outer * H * W * C;
    if (n == 0) return LPC_OK;
    if (dtype == LPC_F32) k_op_finite_diff_adj<float><<<op_blocks(n), 256, 0, st>>>((const float*)dev_u, (float*)dev_out, outer, H, W, C);
    else k_op_finite_diff_adj<double><<<op_blocks(n), 256, 0, st>>>((const double*)dev_u, (double*)dev_out, outer, H, W, C);
    LPC_OP_EPILOGUE();
}
int lpc_op_finite_diff_gram(int device, int dtype, void* dev_out, int D, int Hp, int Wp, int C, void* stream) {
    if (!dev_out || D < 1 || Hp < 1 || Wp < 1 || C < 1) return LPC_EINVAL;
    LPC_OP_PROLOGUE(device);
    const int Nh = Wp / 2 + 1;
    const int64_t n = (int64_t)D * Hp * Nh * C;
    if (dtype == LPC_F32) k_op_finite_diff_gram<float><<<op_blocks(n), 256, 0, st>>>((float2*)dev_out, D, Hp, Wp, Nh, C);
    else k_op_finite_diff_gram<double><<<op_blocks(n), 256, 0, st>>>((double2*)dev_out, D, Hp, Wp, Nh, C);
    LPC_OP_EPILOGUE();
}


int lpc_io_load_psf(int device, int raw_bits, const void* dev_raw, int H, int W, int C, int flags, int bg_pix0, int bg_pix1,
                    int out_h, int out_w, void* dev_psf_out, void* dev_bg_out, void* stream) {
    if (!dev_raw || !dev_psf_out || H < 1 || W < 1 || (C != 1 && C != 3) || out_h < 1 || out_w < 1) return LPC_EINVAL;
    if (raw_bits != 8 && raw_bits != 16) return LPC_EINVAL;
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (raw_bits == 8) return io_load<uint8_t>(1, (const uint8_t*)dev_raw, H, W, C, flags, bg_pix0, bg_pix1, nullptr, out_h, out_w, 1,
                                               (float*)dev_psf_out, (float*)dev_bg_out, st);
    return io_load<uint16_t>(1, (const uint16_t*)dev_raw, H, W, C, flags, bg_pix0, bg_pix1, nullptr, out_h, out_w, 1,
                             (float*)dev_psf_out, (float*)dev_bg_out, st);
}
int lpc_io_load_image(int device, int raw_bits, const void* dev_raw, int H, int W, int C, int flags, const void* dev_bg,
                      int out_h, int out_w, int normalize, void* dev_out, void* stream) {
    if (!dev_raw || !dev_out || H < 1 || W < 1 || (C != 1 && C != 3) || out_h < 1 || out_w < 1) return LPC_EINVAL;
    if (raw_bits != 8 && raw_bits != 16) return LPC_EINVAL;
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (raw_bits == 8) return io_load<uint8_t>(0, (const uint8_t*)dev_raw, H, W, C, flags, 0, 0, (const float*)dev_bg, out_h, out_w, normalize,
                                               (float*)dev_out, nullptr, st);
    return io_load<uint16_t>(0, (const uint16_t*)dev_raw, H, W, C, flags, 0, 0, (const float*)dev_bg, out_h, out_w, normalize,
                             (float*)dev_out, nullptr, st);
}
int lpc_io_rgb2gray(int device, const void* dev_rgb, void* dev_gray, int64_t npix, void* stream) {
    if (!dev_rgb || !dev_gray || npix < 0) return LPC_EINVAL;
    DeviceGuard guard(device);
    if (npix == 0) return LPC_OK;
    k_io_rgb2gray<<<op_blocks(npix), 256, 0, (cudaStream_t)stream>>>((const float*)dev_rgb, (float*)dev_gray, npix);
    return cudaGetLastError() == cudaSuccess ? LPC_OK : LPC_ECUDA;
}

int lpc_bytes_per_iter(lpc_plan_t plan, double* algorithmic_bytes) {
    LPC_NEED(plan);
    if (!algorithmic_bytes) return LPC_EINVAL;
    return plan->bytes_per_iter(algorithmic_bytes);
}
int lpc_launches_per_iter(lpc_plan_t plan, int* total, int* own) {
    LPC_NEED(plan);
    return plan->launches_per_iter(total, own);
}
int lpc_profile(lpc_plan_t plan, int n_iter, void* stream, double* phase_ms, int* n_phases) {
    LPC_NEED(plan);
    if (!phase_ms || !n_phases || n_iter < 1) return LPC_EINVAL;
    return plan->profile(n_iter, (cudaStream_t)stream, phase_ms, n_phases);
}
const char* lpc_phase_name(lpc_plan_t plan, int phase) { return plan ? plan->phase_name(phase) : ""; }
int lpc_engine_version(lpc_plan_t plan) { return plan ? plan->engine_version() : 0; }
int lpc_plan_bytes(lpc_plan_t plan, size_t* bytes) {
    LPC_NEED(plan);
    if (!bytes) return LPC_EINVAL;
    *bytes = plan->plan_bytes();
    return LPC_OK;
}

}  // extern "C"
