/*
 * lpc.h -- C-ABI of the B200-native iterative deconvolution engine
 *          (liblpc.so, built from lenslesspicam_b200/csrc/ for sm_100a).
 *
 * The reference (LCAV/LenslessPiCam v1.0.8) is 100 % Python and has no FFI of
 * its own; the boundary it exposes for this path is the Python class surface
 *   lensless.recon.rfft_convolve.RealFFTConvolve2D   (rfft_convolve.py:26-223)
 *   lensless.recon.recon.ReconstructionAlgorithm     (recon.py:179-653)
 *   lensless.ADMM                                    (admm.py:24-338)
 *   lensless.GradientDescent / NesterovGradientDescent / FISTA (gd.py:62-241)
 * the lenslesspicam_b200 Python package mirrors that surface and binds the entry points below
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns LPC_OK (0) or a negative LPC_E* code; no C++
 *     exception crosses the boundary; lpc_last_error() gives the message.
 *   - every pointer named dev_* is a DEVICE pointer valid on the plan's device;
 *     host_* pointers are host memory (pinned recommended).  The library never
 *     takes ownership of caller memory.
 *   - every call enqueues on the caller's stream (a cudaStream_t passed as
 *     void*; NULL = legacy default stream) and returns WITHOUT synchronising,
 *     except where stated.
 *   - caller arrays are in the reference's channels-last layout
 *     data [B, D, h, w, C], PSF [D, h, w, C] (recon.py:223-224), contiguous unless
 *     explicit element strides are given; `dtype` selects float or double.
 *   - a plan is not re-entrant: one host thread / stream at a time per plan.
 */
#ifndef LPC_H_
#define LPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPC_VERSION 100

/* error codes */
#define LPC_OK 0
#define LPC_EINVAL -1   /* bad argument / shape / enum                                   */
#define LPC_ECUDA -2    /* CUDA runtime error (message in lpc_last_error)                */
#define LPC_ECUFFT -3   /* cuFFT error                                                    */
#define LPC_ESTATE -4   /* call order: PSF or data not set, wrong algorithm for the call  */
#define LPC_ENOMEM -5   /* device allocation failed                                       */
#define LPC_ENOTIMPL -6 /* combination the reference raises NotImplementedError for       */

/* algorithms (reference class each one replaces) */
#define LPC_ALGO_CONV 0     /* RealFFTConvolve2D only         rfft_convolve.py:26  */
#define LPC_ALGO_ADMM 1     /* lensless.ADMM                  admm.py:24           */
#define LPC_ALGO_GD 2       /* lensless.GradientDescent       gd.py:62             */
#define LPC_ALGO_NESTEROV 3 /* NesterovGradientDescent        gd.py:143            */
#define LPC_ALGO_FISTA 4    /* lensless.FISTA                 gd.py:191            */

/* dtype (recon.py:264-290: only float32 / float64 exist in the reference) */
#define LPC_F32 0
#define LPC_F64 1

/* FFT normalisation of H (rfft_convolve.py:27,121-123; only H carries `norm`) */
#define LPC_NORM_BACKWARD 0
#define LPC_NORM_ORTHO 1
#define LPC_NORM_FORWARD 2

/* lpc_get_state selectors (planar -> channels-last, padded shape [B,D,Hp,Wp,C] unless noted) */
#define LPC_STATE_IMAGE_EST 0 /* ADMM: V (padded). GD family: x (unpadded [B,D,h,w,C]) */
#define LPC_STATE_FORWARD 1   /* ADMM: H V (padded)                                     */
#define LPC_STATE_XI 2        /* ADMM: xi' = xi - mu1 X  (half-updated dual)  */
#define LPC_STATE_RHO 3       /* ADMM: rho' = rho - mu3 W                                */
#define LPC_STATE_ETA0 4      /* ADMM: eta'[...,0]                                       */
#define LPC_STATE_ETA1 5      /* ADMM: eta'[...,1]                                       */

/* lpc_plan_create flags */
#define LPC_FLAG_NONE 0
#define LPC_FLAG_NO_GRAPH 1   /* never capture the iteration loop in a CUDA graph        */
#define LPC_FLAG_ENGINE_V1 4  /* K1/K2/K3 + cuFFT engine instead of the fused own-FFT engine (A/B, bisecting) */
#define LPC_FLAG_PSF_PER_SAMPLE 8 /* the plan holds ONE PSF PER BATCH ENTRY: lpc_set_psf takes [B,D,h,w,C]; H, R_divmat,
                                     alpha and the initial value are per sample (multi-mask batches: the reference's
                                     forward(batch, psfs) builds RealFFTConvolve2D(psfs[B,...]), trainable_recon.py:347-350;
                                     eval/benchmark.py:186-190,255-259) */

#define LPC_FLAG_PNP 16 /* plug-and-play plan: a caller-side denoiser replaces the prox / projection, the iteration is driven
                           half by half through lpc_pnp_begin / lpc_pnp_finish (cuFFT engine, no CUDA graph) */

/* lpc_set_option selectors */
#define LPC_OPT_KEEP_X0 1      /* != 0: lpc_set_psf keeps the GD-family initial value (psf min/max, gd.py:99-105) of the
                                  PSF set BEFORE: UnrolledFISTA.forward(psfs) swaps the convolver only, its _image_init
                                  stays the constructor PSF's (trainable_recon.py:347-350, unrolled_fista.py:55-59) */
#define LPC_OPT_PNP_USE_DUAL 3 /* PnP-ADMM: denoiser["use_dual"] (admm.py:127, 236-243, 266-275) */
#define LPC_OPT_B2_PREFETCH 2  /* tuning: L2-prefetch distance (rows) of the fused spatial kernel, 0 = off */

typedef struct lpc_plan_s* lpc_plan_t;

int lpc_version(void);
const char* lpc_error_name(int code);
/* message of the last failure on `plan` (or, with plan == NULL, of the last
 * failed lpc_plan_create on this thread). Valid until the next call. */
const char* lpc_last_error(lpc_plan_t plan);

/* ---- geometry: integer work, bit-exact with rfft_convolve.py:111-117 -------- */
/* smallest 5-smooth integer >= n  (scipy.fftpack.next_fast_len, rfft_convolve.py:16,112) */
int lpc_next_fast_len(int n);
/* padded = next_fast_len(2*dim-1); start = (padded-dim)//2 */
int lpc_geometry(int h, int w, int* Hp, int* Wp, int* start_y, int* start_x);

/* ---- plan life cycle ------------------------------------------------------- */
/* Allocates planar solver state, cuFFT plans and constant tables on `device`
 * for B images of PSF-shape [D,h,w,C].  Replaces the constructor chain
 * ADMM.__init__ (admm.py:35-135) / GradientDescent.__init__ (gd.py:67-92) ->
 * ReconstructionAlgorithm.__init__ (recon.py:203-329) -> RealFFTConvolve2D.__init__
 * (rfft_convolve.py:27-77).  `norm` is the convolver's norm (ADMM: backward,
 * GD family: ortho).  `pad` mirrors RealFFTConvolve2D.pad and is only
 * meaningful for LPC_ALGO_CONV (ADMM always optimises the padded image, the GD
 * family the unpadded one). */
int lpc_plan_create(lpc_plan_t* out, int device, int algo, int dtype, int norm, int B, int D,
                    int C, int h, int w, int flags);
int lpc_plan_destroy(lpc_plan_t plan);
/* Does this library's compiled menu of the fused engine hold the padded row length Wp / column length Hp?  (A size
 * outside it runs the cuFFT engine; the Python mirror compiles a size-specific library on demand instead.) */
int lpc_menu_query(int Wp, int Hp, int* has_rows, int* has_cols);

/* RealFFTConvolve2D.set_psf (rfft_convolve.py:102-131): pad PSF, H = rfft2(.)*norm,
 * plus the algorithm constants derived from H: ADMM R_divmat (admm.py:186-190),
 * GD step alpha and initial estimate (gd.py:96-112).  dev_psf is [D,h,w,C]. */
int lpc_set_psf(lpc_plan_t plan, const void* dev_psf, void* stream);  /* [B,D,h,w,C] with LPC_FLAG_PSF_PER_SAMPLE */

/* ADMM hyper-parameters (admm.py:39-42); recomputes R_divmat / X_divmat. */
int lpc_set_admm_params(lpc_plan_t plan, double mu1, double mu2, double mu3, double tau,
                        void* stream);
/* GD family: lip_fact (gd.py:67), Nesterov mu (gd.py:153, always 0.9 after reset --
 * SURVEY 8b quirk v), FISTA initial tk (gd.py:200). */
int lpc_set_gd_params(lpc_plan_t plan, double lip_fact, double nesterov_mu, double fista_tk,
                      void* stream);

/* Unrolled solvers: per-iteration hyper-parameters (inference only, no autograd).
 * ADMM: mu1[i], mu2[i], mu3[i], tau[i], i < n -- the |.| of UnrolledADMM's parameter vectors
 * (unrolled_admm.py:80-99,141-144); iteration i of every lpc_run uses entry i (R_divmat and X_divmat of
 * that iteration, unrolled_admm.py:166-179,181-234); n_iter <= n.  n == 0 restores the constant
 * parameters of lpc_set_admm_params.
 * FISTA: alpha[i*C + c] and tk[0..n] (unrolled_fista.py:60-83,102-106): x <- x - alpha_i grad,
 * xk = proj(x), x = xk + (tk_i - 1)/tk_{i+1} (xk - xk_prev), with xk_prev initialised to the initial
 * estimate (unrolled_fista.py:95-100; lensless.FISTA's first-iteration aliasing does not apply).
 * The position advances across runs until lpc_reset.  Host arrays, copied. */
int lpc_set_admm_schedule(lpc_plan_t plan, int n, const double* mu1, const double* mu2, const double* mu3,
                          const double* tau, void* stream);
int lpc_set_fista_schedule(lpc_plan_t plan, int n, const double* alpha, const double* tk, void* stream);

/* ReconstructionAlgorithm.set_data (recon.py:352-381) + optional background
 * subtraction and clamp (recon.py:553-555; dev_background may be NULL).
 * dev_data is [B, Dd, h, w, C] with Dd in {1, D}; strides are ELEMENT strides
 * for (B, D, h, w, C); pass strides == NULL for contiguous.  dev_echo (may be
 * NULL) receives the background-subtracted data, contiguous [B,Dd,h,w,C]: the
 * reference overwrites self._data with it (recon.py:554). */
int lpc_set_data(lpc_plan_t plan, const void* dev_data, int Dd, const int64_t* strides,
                 const void* dev_background, const int64_t* bg_strides, void* dev_echo,
                 void* stream);
/* same from HOST memory: stages through an internal device buffer (cudaMemcpyAsync) */
int lpc_set_data_host(lpc_plan_t plan, const void* host_data, int Dd, void* stream);

/* reset() (admm.py:150-230, gd.py:94-126, 174-181, 227-233).  dev_initial_est is
 * NULL or the warm start: ADMM [B,D,Hp,Wp,C] (padded! recon.py:296-299),
 * GD family [B,D,h,w,C]. */
int lpc_reset(lpc_plan_t plan, const void* dev_initial_est, void* stream);

/* n_iter calls of _update (recon.py:575-576; admm.py:313-329; gd.py:132-134,
 * 183-188, 235-241). Captured once per (plan, n_iter) into a CUDA graph unless
 * LPC_FLAG_NO_GRAPH. */
int lpc_run(lpc_plan_t plan, int n_iter, void* stream);

/* _form_image (admm.py:331-338: crop + clamp; gd.py:136-140: non_neg) written
 * channels-last to dev_out [B,D,h,w,C] (contiguous). */
int lpc_get_image(lpc_plan_t plan, void* dev_out, void* stream);
/* same into HOST memory; does NOT synchronise (caller syncs the stream). */
int lpc_get_image_host(lpc_plan_t plan, void* host_out, void* stream);

/* debug / parity: copy one state array out, channels-last (see LPC_STATE_*) */
int lpc_get_state(lpc_plan_t plan, int which, void* dev_out, void* stream);
/* H as the reference stores it: complex [D,Hp,Wp/2+1,C] interleaved (re,im)  (rfft_convolve.py:121-124) */
int lpc_get_H(lpc_plan_t plan, void* dev_out, void* stream);
/* GD step sizes alpha[C] (gd.py:108-112) */
int lpc_get_alpha(lpc_plan_t plan, void* dev_out, void* stream);

/* ---- RealFFTConvolve2D.convolve / deconvolve (rfft_convolve.py:133-223) ------ */
/* LPC_ALGO_CONV plans only.  pad != 0: dev_in/out are [B,Dd,h,w,C] / [B,D,h,w,C];
 * pad == 0: both are padded [B,*,Hp,Wp,C].  in_strides as in lpc_set_data
 * (stride 0 broadcasts a size-1 depth or channel axis, test_convolver.py:32-59). */
int lpc_convolve(lpc_plan_t plan, const void* dev_in, const int64_t* in_strides, void* dev_out,
                 int adjoint, int pad, void* stream);

/* ReconstructionAlgorithm.reconstruction_error (recon.py:607-653) on an LPC_ALGO_CONV plan built with the solver's
 * convolver parameters: Hx = crop(convolve(pad(prediction))); normalize != 0: Hx = (Hx - min)/max(Hx - min) per
 * (batch, depth) over (h,w,C) (recon.py:640-644); dev_out[b] = sum_{D,h,w,C} (Hx - lensless)^2 / npix, T[B].
 * dev_prediction [B,D,h,w,C] and dev_lensless [B,Dl,h,w,C] (Dl in {1,D}) contiguous.  npix = prod(psf.shape)
 * (recon.py:251: self._npix). */
int lpc_reconstruction_error(lpc_plan_t plan, const void* dev_prediction, const void* dev_lensless, int Dl,
                             int normalize, double npix, void* dev_out, void* stream);

/* ReconstructionAlgorithm.set_image_estimate (recon.py:415-442): overwrite the current iterate only.  GD family:
 * momentum / t_k / previous iterate are kept.  ADMM: a warm start (duals zeroed) -- documented deviation.
 * dev_est as in lpc_reset. */
int lpc_set_image_estimate(lpc_plan_t plan, const void* dev_est, void* stream);

/* ---- module-level helpers of lensless/recon/admm.py (imported by callers such as unrolled_admm.py:10), plan-less.
 * Arrays are the caller's contiguous channels-last [..., H, W, C] (outer = product of the leading axes).
 *   soft_thresh      admm.py:341-346     out[i] = sign(x) max(|x| - thresh, 0), n elements
 *   finite_diff      admm.py:349-359     [outer,H,W,C] -> [outer,H,W,C,2]   (periodic backward differences)
 *   finite_diff_adj  admm.py:362-370     [outer,H,W,C,2] -> [outer,H,W,C]   (periodic forward differences, summed)
 *   finite_diff_gram admm.py:373-397     -> complex [D,Hp,Wp/2+1,C] = rfft2 of the stencil, closed form */
int lpc_op_soft_thresh(int device, int dtype, const void* dev_x, void* dev_out, int64_t n, double thresh, void* stream);
int lpc_op_finite_diff(int device, int dtype, const void* dev_x, void* dev_out, int64_t outer, int H, int W, int C, void* stream);
int lpc_op_finite_diff_adj(int device, int dtype, const void* dev_u, void* dev_out, int64_t outer, int H, int W, int C, void* stream);
int lpc_op_finite_diff_gram(int device, int dtype, void* dev_out, int D, int Hp, int Wp, int C, void* stream);

/* ---- plug-and-play iterations (LPC_FLAG_PNP plans).  The reference lets a torch denoiser replace the U prox of ADMM
 * (admm.py:126-133, 235-243, 266-275) and the projection of the gradient-descent family (gd.py:89-92, 136-140); the engine
 * splits an iteration at that call:
 *   lpc_pnp_begin   writes the denoiser's input, channels-last: ADMM [B,D,Hp,Wp,C] = U + eta/mu2 (use_dual) or V;
 *                   GD family [B,D,h,w,C] = the stepped estimate x - alpha*grad (Nesterov: with its momentum)
 *   lpc_pnp_finish  takes the denoised array (same shape) and completes the iteration (ADMM: X, W, image update with the
 *                   reference's rk expression, convolve, dual updates; FISTA: momentum).
 * One begin/finish pair = one _update (admm.py:313-329, gd.py:132-134, 183-188, 235-241). */
int lpc_pnp_begin(lpc_plan_t plan, void* dev_denoiser_in, void* stream);
int lpc_pnp_finish(lpc_plan_t plan, const void* dev_denoised, void* stream);

/* ---- the input side on the GPU: lensless/utils/io.py load_psf (:210-388) and load_image (:20-207) on a
 * decoded sensor frame resident in device memory (raw uint8 / uint16 [H,W,C], C in {1,3}; PNG decode stays on the
 * host).  float32 channels-last outputs [1, out_h, out_w, C] that feed lpc_set_psf / lpc_set_data directly.
 *   lpc_io_load_psf   BGR->RGB + flips, background = per-channel mean of the [bg_pix0, bg_pix1)^2 corner (skipped when
 *                     bg_pix1 <= bg_pix0), subtract + clip, antialiased resize (torchvision Resize(antialias=True),
 *                     utils/image.py:28-67), optional channel sum (single_psf), L2 normalisation; dev_bg_out[C] receives
 *                     the background level / max_val (io.py:376) for the measurement's subtraction
 *   lpc_io_load_image BGR->RGB + flips, subtract dev_bg[C] * max_val (io.py:164-171; NULL: none) + clip, resize,
 *                     normalize != 0: divide by the maximum (io.py:196-197)
 *   lpc_io_rgb2gray   utils/image.py:158-206 (numpy branch): [npix,3] -> [npix] with weights 0.299/0.587/0.114 */
#define LPC_IO_BGR 1        /* raw channel order is BGR (cv2.imread), io.py:151-152 */
#define LPC_IO_FLIP_UD 2
#define LPC_IO_FLIP_LR 4
#define LPC_IO_SINGLE_PSF 8 /* io.py:363-368 */
#define LPC_IO_NORMALIZE_FIRST 16 /* lpc_io_load_image: divide by the maximum BEFORE the resize -- load_data normalises the
                                     full-resolution measurement and only then resizes it to the PSF's shape (io.py:509-513,545-547) */
int lpc_io_load_psf(int device, int raw_bits, const void* dev_raw, int H, int W, int C, int flags, int bg_pix0, int bg_pix1,
                    int out_h, int out_w, void* dev_psf_out, void* dev_bg_out, void* stream);
int lpc_io_load_image(int device, int raw_bits, const void* dev_raw, int H, int W, int C, int flags, const void* dev_bg,
                      int out_h, int out_w, int normalize, void* dev_out, void* stream);
int lpc_io_rgb2gray(int device, const void* dev_rgb, void* dev_gray, int64_t npix, void* stream);

/* plan-level switches, see LPC_OPT_* */
int lpc_set_option(lpc_plan_t plan, int option, double value);

/* ---- introspection for the roofline ------------------ */
/* algorithmic bytes per iteration for the whole plan (all B images):
 * ADMM 16P + 9.5Sp + d, GD family 10d + 6Sp. */
int lpc_bytes_per_iter(lpc_plan_t plan, double* algorithmic_bytes);
/* number of kernels (ours + cuFFT's) one iteration launches, ours only in *own */
int lpc_launches_per_iter(lpc_plan_t plan, int* total, int* own);
/* Runs n_iter iterations WITHOUT the CUDA graph, bracketing every phase of the
 * iteration with CUDA events on `stream`; returns the mean device time (ms) of each
 * phase in phase_ms[0..*n_phases) (at most 8).  Synchronises.  Names: lpc_phase_name. */
int lpc_profile(lpc_plan_t plan, int n_iter, void* stream, double* phase_ms, int* n_phases);
const char* lpc_phase_name(lpc_plan_t plan, int phase);
/* 2 = fused own-FFT engine (lpc_fused.cuh), 1 = K1/K2 + cuFFT */
int lpc_engine_version(lpc_plan_t plan);
/* device bytes held by the plan */
int lpc_plan_bytes(lpc_plan_t plan, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* LPC_H_ */
