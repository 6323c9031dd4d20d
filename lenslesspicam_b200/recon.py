"""
GPU mirror of ``lensless.recon.recon.ReconstructionAlgorithm``
(reference: lensless/recon/recon.py:179-653).

Same constructor / ``set_data`` / ``apply`` / ``reset`` / ``get_image_estimate``
surface, so the callers in scripts/recon/admm.py, scripts/recon/gradient_descent.py,
profile/admm.py, profile/gradient_descent.py and lensless/eval/benchmark.py work
unchanged with a CUDA PSF.  All arithmetic of the iteration loop runs in
liblpc.so (hand-written sm_100a kernels + cuFFT) on a private planar state; this
class only holds torch tensors and passes raw device pointers.

Deliberate deviations from the reference, each also in DESIGN.md:

* CUDA tensors only; numpy / CPU tensors raise ``TypeError`` (no CPU fallback).
* ``apply`` returns a FRESH tensor, not a view into solver state, and does not
  clamp the ADMM iterate in place (reference quirks i/ii).  A user-supplied
  ``initial_est`` is never mutated (quirk iii).
* batch extension: ``set_data`` accepts ``[B, D, h, w, C]`` with B > 1 (the
  reference asserts B == 1 in ``apply``, recon.py:549-551); ``apply`` then returns
  ``[B, D, h, w, C]``.  With B == 1 the return is ``[D, h, w, C]`` as in the reference.
* ``denoiser`` (plug-and-play) and custom ``proj`` / ``psi`` callables are not
  fusable and raise ``NotImplementedError``.
"""
import abc
import ctypes
import pathlib as plib
import warnings
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._engine import EngineStream, torch_dtype_code
from .rfft_convolve import RealFFTConvolve2D


def get_dtype(dtype=None, is_torch=True):
    """utils/io.py:645-674: only None / "float32" / "float64" strings are accepted."""
    if dtype is None:
        dtype = "float32"
    assert dtype == "float32" or dtype == "float64"
    return torch.float32 if dtype == "float32" else torch.float64


class ReconstructionAlgorithm(abc.ABC):
    _ALGO = None  # LPC_ALGO_* of the subclass

    def __init__(
        self,
        psf,
        dtype=None,
        pad=True,
        n_iter=100,
        initial_est=None,
        reset=True,
        denoiser=None,
        **kwargs,
    ):
        super().__init__()
        if not isinstance(psf, torch.Tensor) or not psf.is_cuda:
            raise TypeError(
                "lenslesspicam_b200 needs a CUDA torch.Tensor PSF: the engine has no CPU path "
                "(for numpy / CPU tensors use the reference's own classes)"
            )
        self.is_torch = True
        # recon.py:255-256
        assert len(psf.shape) == 4, "PSF must be 4D: (depth, height, width, channels)."
        assert psf.shape[3] == 3 or psf.shape[3] == 1, "PSF must either be rgb (3) or grayscale (1)"
        # recon.py:308-322: plug-and-play.  The reference only knows the name "DruNet" (weights fetched from the hub);
        # here ``denoiser["network"]`` may also be any callable ``f(image, noise_level) -> image`` on torch CUDA tensors.
        # The iteration is then split around the call (lpc_pnp_begin / lpc_pnp_finish, csrc/lpc_pnp.cuh).
        self._denoiser = None
        self._denoiser_noise_level = None
        self._denoiser_use_dual = False
        if denoiser is not None:
            assert "network" in denoiser.keys()
            assert "noise_level" in denoiser.keys()
            net = denoiser["network"]
            if callable(net):
                self._denoiser = net
            elif net == "DruNet":
                self._denoiser = _load_drunet_function(psf.device)
            else:
                raise NotImplementedError(f"Unsupported denoiser: {net}")
            self._denoiser_noise_level = denoiser["noise_level"]
            self._denoiser_use_dual = bool(denoiser.get("use_dual", False))
        self._n_iter = n_iter
        self._npix = np.prod(psf.shape)
        self._psf_shape = np.array(psf.shape)

        # recon.py:263-277 (dtype is a *string* or None in the reference)
        tdtype = get_dtype(dtype, True)
        self._psf = psf.type(tdtype).contiguous()
        self._psf_ctor = self._psf
        self._dtype = tdtype
        self._complex_dtype = torch.complex64 if tdtype == torch.float32 else torch.complex128
        self._device = self._psf.device
        self._dev_index = self._device.index if self._device.index is not None else torch.cuda.current_device()

        # recon.py:292-299 (unknown kwargs ride along and are ignored, like rfft_convolve.py:27)
        self._convolver_param = {"dtype": dtype, "pad": pad, **kwargs}
        self._pad = pad
        self._norm = kwargs.get("norm", "ortho")
        if self._norm not in _lib.NORMS:
            raise ValueError(f"Invalid norm: {self._norm}")
        self._convolver_obj = None
        self._err_convolver = None  # convolver for reconstruction_error(psfs=...) with PSFs other than the solver's
        lib = _lib.load()
        Hp, Wp, sy, sx = (ctypes.c_int() for _ in range(4))
        D, h, w, C = (int(v) for v in self._psf.shape)
        _lib.check(lib.lpc_geometry(h, w, Hp, Wp, sy, sx))
        self._padded_shape = [D, Hp.value, Wp.value, C]
        self._start_idx = np.array([sy.value, sx.value])
        self._end_idx = self._start_idx + self._psf_shape[-3:-1]
        if pad:
            self._image_est_shape = self._psf_shape
        else:
            self._image_est_shape = self._padded_shape

        self._stream = EngineStream(self._device)
        # (batch size, one-PSF-per-sample) -> lpc_plan_t, least recently used first; at most _MAX_PLANS are kept (a
        # full-size ADMM plan is ~1.5 GB per image) and none exists until the first call that needs one
        self._plans = OrderedDict()
        self._plan = None     # plan of the current batch
        self._B = 1
        self._per_sample = False
        self._psfs = None     # [B, D, h, w, C] when the current plan holds one PSF per sample
        self._flags = _lib.FLAG_NO_GRAPH if kwargs.get("no_graph", False) else _lib.FLAG_NONE
        if kwargs.get("engine", None) == "v1":  # K1/K2 + cuFFT instead of the fused own-FFT engine (A/B, bisecting)
            self._flags |= _lib.FLAG_ENGINE_V1
        if self._denoiser is not None:
            self._flags |= _lib.FLAG_PNP

        if initial_est is not None:
            self._set_initial_estimate(initial_est)
        else:
            self._initial_est = None
        self._data = None
        self._data_sent = False
        self._pending_psfs = None  # per-sample PSFs of the batch being sent (forward(batch, psfs))

        # used inside trainable recon in the reference; kept for attribute compatibility
        self.compensation_branch = None
        self.compensation_branch_inputs = None

        if reset:
            self.reset()

    # ------------------------------------------------------------------ plans
    _MAX_PLANS = 2
    _warned_v1 = False

    def _select_plan(self, B, psfs=None):
        """Plan for a batch of ``B`` measurements; ``psfs`` ([B, D, h, w, C]) selects a plan that holds one PSF per
        sample.  New plans get the hyper-parameters, the PSF(s) and a reset."""
        B = int(B)
        per_sample = psfs is not None
        key = (B, per_sample)
        lib = _lib.load()
        plan = self._plans.get(key)
        created = plan is None
        if created:
            while len(self._plans) >= self._MAX_PLANS:
                _, old = self._plans.popitem(last=False)
                if self._plan is not None and old.value == self._plan.value:
                    self._plan = None
                lib.lpc_plan_destroy(old)
            plan = ctypes.c_void_p()
            D, h, w, C = (int(v) for v in self._psf.shape)
            flags = self._flags | (_lib.FLAG_PSF_PER_SAMPLE if per_sample else 0)
            creator = lib
            if self._dtype == torch.float32 and not (flags & (_lib.FLAG_ENGINE_V1 | _lib.FLAG_PNP)):
                from . import _jit  # a padded size outside the compiled menu gets its own specialised library

                creator = _jit.library_for(self._padded_shape[2], self._padded_shape[1])
            _lib.check(
                creator.lpc_plan_create(
                    ctypes.byref(plan), self._dev_index, self._ALGO, torch_dtype_code(self._dtype),
                    _lib.NORMS[self._norm], B, D, C, h, w, flags,
                )
            )
            self._plans[key] = plan
            if self._denoiser is not None:
                _lib.check(lib.lpc_set_option(plan, _lib.OPT_PNP_USE_DUAL, 1.0 if self._denoiser_use_dual else 0.0), plan)
            if (self._dtype == torch.float32 and lib.lpc_engine_version(plan) != 2 and not (self._flags & (_lib.FLAG_ENGINE_V1 | _lib.FLAG_PNP))
                    and not ReconstructionAlgorithm._warned_v1):
                ReconstructionAlgorithm._warned_v1 = True
                _, Hp, Wp, _ = self._padded_shape
                warnings.warn(
                    f"lenslesspicam_b200: padded size {Hp}x{Wp} is not in the fused engine's compiled menu "
                    "(csrc/lpc_fused.cuh: LPC_ROW_MENU / LPC_COL_MENU) and could not be compiled on demand (LPC_JIT=0, no nvcc, or "
                    "more than four radix passes); this solver runs the cuFFT-based engine v1, "
                    "about 2x slower.  `engine_version` on the solver reports which engine serves it.",
                    RuntimeWarning, stacklevel=3,
                )
        self._plans.move_to_end(key)
        self._plan, self._B, self._per_sample = plan, B, per_sample
        if created or per_sample:
            with self._stream as st:
                if created:
                    self._push_params(plan, st)
                    if self._keep_x0:
                        # the initial value stays the CONSTRUCTOR PSF's whatever PSF arrives later (UnrolledFISTA:
                        # trainable_recon.py:347-350 swaps the convolver only, unrolled_fista.py:55-59 keeps _image_init)
                        base = self._psf_ctor if not per_sample else self._psf_ctor[None].expand((B,) + tuple(self._psf_ctor.shape)).contiguous()
                        _lib.check(lib.lpc_set_psf(plan, base.data_ptr(), st), plan)
                        _lib.check(lib.lpc_set_option(plan, _lib.OPT_KEEP_X0, 1.0), plan)
                if per_sample:
                    self._psfs = psfs.to(device=self._device, dtype=self._dtype).contiguous()
                    _lib.check(lib.lpc_set_psf(plan, self._psfs.data_ptr(), st), plan)
                else:
                    _lib.check(lib.lpc_set_psf(plan, self._psf.data_ptr(), st), plan)
                if created:
                    self._after_new_plan(plan, st)
            self._reset_plan(self._initial_est)
        return plan

    _keep_x0 = False

    def _after_new_plan(self, plan, stream):
        """hook: parameters that need the PSF on the plan first (unrolled schedules)"""

    def _ensure_plan(self):
        if self._plan is None:
            self._select_plan(self._B)
        return self._plan

    @property
    def engine_version(self):
        """2 = fused own-FFT engine (fp32, padded sizes of the compiled menu); 1 = K1/K2/K3 + cuFFT."""
        return int(_lib.load().lpc_engine_version(self._ensure_plan()))

    @abc.abstractmethod
    def _push_params(self, plan, stream):
        """send the algorithm's hyper-parameters to a plan"""

    def __del__(self):
        try:
            lib = _lib.load()
            for plan in self._plans.values():
                lib.lpc_plan_destroy(plan)
            self._plans = OrderedDict()
        except Exception:
            pass

    @property
    def _convolver(self):
        """RealFFTConvolve2D mirror (built on first use; the solver itself does not need it)."""
        if self._convolver_obj is None:
            kw = {k: v for k, v in self._convolver_param.items() if k not in ("dtype", "pad", "no_graph", "engine")}
            self._convolver_obj = RealFFTConvolve2D(self._psf, dtype=self._dtype, pad=self._pad, **kw)
        return self._convolver_obj

    # ------------------------------------------------------------------ data
    def set_data(self, data):
        """recon.py:352-381"""
        assert isinstance(data, torch.Tensor)
        assert len(data.shape) >= 3, "Data must be at least 3D: [..., width, height, channel]."
        assert np.all(self._psf_shape[-3:-1] == np.array(data.shape)[-3:-1]), "PSF and data shape mismatch"
        if len(data.shape) == 3:
            self._data = data[None, None, ...]
        elif len(data.shape) == 4:
            self._data = data[None, ...]
        else:
            self._data = data
        self._data_sent = False

    def _send_data(self, background=None):
        data = self._data
        if not data.is_cuda:
            raise TypeError("data must be a CUDA tensor on the PSF's device")
        B, Dd = int(data.shape[0]), int(data.shape[1])
        D, h, w, C = (int(v) for v in self._psf.shape)
        if Dd not in (1, D) or int(data.shape[4]) != C:
            raise ValueError("data depth must be 1 or the PSF depth, and channels must match the PSF")
        data = data.to(self._dtype)
        plan = self._select_plan(B, self._pending_psfs)
        lib = _lib.load()
        echo = None
        bg_ptr, bg_strides = None, None
        if background is not None:
            bg = background.to(device=self._device, dtype=self._dtype)
            while bg.dim() < 5:
                bg = bg[None]
            bg = bg.expand(B, Dd, h, w, C)
            bg_ptr, bg_strides = bg.data_ptr(), _lib.strides_arg(bg.stride())
            echo = torch.empty((B, Dd, h, w, C), dtype=self._dtype, device=self._device)
        with self._stream as st:
            _lib.check(
                lib.lpc_set_data(
                    plan, data.data_ptr(), Dd, _lib.strides_arg(data.stride()), bg_ptr, bg_strides,
                    echo.data_ptr() if echo is not None else None, st,
                ),
                plan,
            )
        if echo is not None:
            self._data = echo  # recon.py:554-555 overwrites self._data
        self._data_sent = True

    # ------------------------------------------------------------------ estimates
    def _check_est(self, image_est):
        assert isinstance(image_est, torch.Tensor)
        assert len(image_est.shape) >= 4, "Image estimate must be at least 4D: [..., depth, width, height, channel]."
        assert np.all(
            np.array(self._image_est_shape[-3:-1]) == np.array(image_est.shape)[-3:-1]
        ), f"Image estimate must be of shape (..., width, height, channel): {self._image_est_shape[-3:-1]}"
        return image_est[None, ...] if len(image_est.shape) == 4 else image_est

    def _set_initial_estimate(self, image_est):
        """recon.py:383-413 (takes effect at the next reset())"""
        self._initial_est = self._check_est(image_est)

    def set_image_estimate(self, image_est):
        """recon.py:415-442: overwrite the current estimate, ``_initial_est`` untouched.  Gradient-descent family: exactly
        that (momentum, t_k and the previous iterate are kept).  ADMM: DEVIATION -- a warm start from ``image_est``
        (duals re-zeroed); the reference keeps its duals and for one iteration the stale H V / Psi V."""
        est = self._check_est(image_est)
        plan = self._ensure_plan()
        est = est.to(device=self._device, dtype=self._dtype)
        est = est.expand((self._B,) + tuple(est.shape[1:])).contiguous()
        lib = _lib.load()
        with self._stream as st:
            _lib.check(lib.lpc_set_image_estimate(plan, est.data_ptr(), st), plan)

    def get_image_estimate(self):
        """recon.py:444-446 -> [B, D, h, w, C]"""
        return self._form_image()

    @property
    def _image_est(self):
        """Current iterate in the reference's layout (ADMM: padded [B,D,Hp,Wp,C]; GD: [B,D,h,w,C])."""
        D, Hp, Wp, C = self._padded_shape
        shape = (self._B, D, Hp, Wp, C) if not self._pad else (self._B,) + tuple(int(v) for v in self._psf.shape)
        out = torch.empty(shape, dtype=self._dtype, device=self._device)
        lib = _lib.load()
        plan = self._ensure_plan()
        with self._stream as st:
            _lib.check(lib.lpc_get_state(plan, _lib.STATE_IMAGE_EST, out.data_ptr(), st), plan)
        return out

    def _set_psf(self, psf):
        """recon.py:448-470"""
        assert psf.shape[-1] == 3 or psf.shape[-1] == 1, "PSF must either be rgb (3) or grayscale (1)"
        assert self._psf.shape == psf.shape, "new PSF must have same shape as old PSF"
        assert isinstance(psf, type(self._psf)), "new PSF must have same type as old PSF"
        self._psf = psf.to(device=self._device, dtype=self._dtype).contiguous()
        if self._convolver_obj is not None:
            self._convolver_obj.set_psf(self._psf)  # keep its plans: a multi-mask evaluation loop changes the PSF per sample
        lib = _lib.load()
        with self._stream as st:
            for (_, per_sample), plan in self._plans.items():
                if not per_sample:
                    _lib.check(lib.lpc_set_psf(plan, self._psf.data_ptr(), st), plan)
        if self._per_sample:  # back to the shared-PSF plan
            self._plan, self._per_sample, self._psfs = None, False, None
            self._data_sent = False
        self.reset()

    # ------------------------------------------------------------------ iteration
    def _reset_plan(self, initial_est=None):
        if self._plan is None:
            return  # no plan yet: a new plan is reset when it is created
        lib = _lib.load()
        ptr = None
        if initial_est is not None:
            est = initial_est.to(device=self._device, dtype=self._dtype)
            est = est.expand((self._B,) + tuple(est.shape[1:])).contiguous()
            ptr = est.data_ptr()
        with self._stream as st:
            _lib.check(lib.lpc_reset(self._plan, ptr, st), self._plan)

    def reset(self):
        """Reset state variables (admm.py:150-230 / gd.py:94-126)."""
        self._reset_plan(self._initial_est)

    def _update(self, iter):
        """One iteration (recon.py:338-343); ``apply`` runs all of them in one call instead."""
        self._ensure_plan()
        self._run(1)

    def _run(self, n_iter):
        """``n_iter`` iterations on the current plan: one captured CUDA graph, or -- plug-and-play -- a host loop of
        begin / denoiser / finish (admm.py:235-243,266-275,313-329; gd.py:132-140)."""
        lib = _lib.load()
        plan = self._plan
        if self._denoiser is None:
            with self._stream as st:
                _lib.check(lib.lpc_run(plan, int(n_iter), st), plan)
            return
        D, Hp, Wp, C = self._padded_shape
        shape = (self._B, D, Hp, Wp, C) if not self._pad else (self._B,) + tuple(int(v) for v in self._psf.shape)
        for _ in range(int(n_iter)):
            buf = torch.empty(shape, dtype=self._dtype, device=self._device)
            with self._stream as st:
                _lib.check(lib.lpc_pnp_begin(plan, buf.data_ptr(), st), plan)
            den = self._denoiser(buf, self._denoiser_noise_level)
            den = den.to(device=self._device, dtype=self._dtype).contiguous()
            assert tuple(den.shape) == tuple(shape), "the denoiser must return an array of its input's shape"
            with self._stream as st:
                _lib.check(lib.lpc_pnp_finish(plan, den.data_ptr(), st), plan)

    def _form_image(self):
        """admm.py:331-338 / gd.py:136-140 -> fresh [B, D, h, w, C] tensor."""
        D, h, w, C = (int(v) for v in self._psf.shape)
        if self._denoiser is not None and self._pad:
            # gd.py:136-140: with a denoiser the projection IS the denoiser, applied to the current iterate
            return self._denoiser(self._image_est, self._denoiser_noise_level)
        out = torch.empty((self._B, D, h, w, C), dtype=self._dtype, device=self._device)
        lib = _lib.load()
        plan = self._ensure_plan()
        with self._stream as st:
            _lib.check(lib.lpc_get_image(plan, out.data_ptr(), st), plan)
        return out

    def _progress(self):
        return

    def _get_numpy_data(self, data):
        return data.detach().cpu().numpy()

    def apply(
        self,
        n_iter=None,
        disp_iter=-1,
        plot_pause=0.2,
        plot=False,
        save=False,
        gamma=None,
        ax=None,
        reset=True,
        background=None,
        **kwargs,
    ):
        """recon.py:498-605.  Extra keyword arguments are accepted and ignored
        (eval/benchmark.py:247-253 passes ``output_intermediate``)."""
        assert self._data is not None, "Must set data with `set_data()`"
        if background is not None or not self._data_sent or int(self._data.shape[0]) != self._B or self._plan is None:
            self._send_data(background)
        if reset:
            self.reset()
        if n_iter is None:
            n_iter = self._n_iter

        lib = _lib.load()
        if (plot or save) and disp_iter is not None:
            plot_image, plt = _import_plotting()
            if ax is None:
                ax = plot_image(self._get_numpy_data(self._form_image()[0]), gamma=gamma)
            if disp_iter is None or disp_iter <= 0:
                disp_iter = n_iter + 1
            done = 0
            while done < n_iter:
                step = min(disp_iter - (done % disp_iter), n_iter - done)
                self._run(step)
                done += step
                if done % disp_iter == 0:
                    self._progress()
                    ax = plot_image(self._get_numpy_data(self._form_image()[0]), ax=ax, gamma=gamma)
                    title = "Reconstruction after iteration {}".format(done)
                    (ax[0, 0] if hasattr(ax, "__len__") else ax).set_title(title)
                    if save:
                        plt.savefig(plib.Path(save) / f"{done}.png")
                    if plot:
                        plt.draw()
                        plt.pause(plot_pause)
        else:
            ax = None
            self._run(n_iter)

        img = self._form_image()
        final_im = img[0] if self._B == 1 else img
        if plot:
            plot_image, plt = _import_plotting()
            ax = plot_image(self._get_numpy_data(final_im), ax=ax, gamma=gamma)
            title = "Final reconstruction after {} iterations".format(n_iter)
            (ax[0, 0] if hasattr(ax, "__len__") else ax).set_title(title)
            if save:
                plt.savefig(plib.Path(save) / f"{n_iter}.png")
            return final_im, ax
        return final_im

    def forward(self, batch, psfs=None, background=None):
        """Batched inference in the shape of ``TrainableReconstructionAlgorithm.forward``
        (trainable_recon.py:297-405, no learnable parameters): ``[B, D|1, h, w, C] -> [B, D, h, w, C]``."""
        return self._forward_batch(batch, psfs, background)

    def _forward_batch(self, batch, psfs, background, **apply_kw):
        """``psfs``: None, one PSF ``[D, h, w, C]`` / ``[1, D, h, w, C]``, or one per sample ``[B, D, h, w, C]``
        (multi-mask datasets, eval/benchmark.py:186-190,255-259; the reference builds ``RealFFTConvolve2D(psfs)`` with
        the batch axis kept, trainable_recon.py:347-350).  Per-sample PSFs run as ONE batched launch chain on a plan that
        holds one transfer function -- and for ADMM one ``R_divmat``, for the gradient-descent family one step size
        and initial value -- per (sample, depth, channel)."""
        assert isinstance(batch, torch.Tensor) and len(batch.shape) == 5, "batch must be of shape (N, D, H, W, C)"
        apply_kw.setdefault("disp_iter", None)
        apply_kw.setdefault("plot", False)
        B = int(batch.shape[0])
        if psfs is not None and psfs.dim() == 5 and int(psfs.shape[0]) > 1:
            assert int(psfs.shape[0]) == B, "one PSF per sample"
            assert tuple(psfs.shape[1:]) == tuple(self._psf.shape), "new PSF must have same shape as old PSF"
            self._pending_psfs = psfs
            try:
                self.set_data(batch)
                out = self.apply(background=background, **apply_kw)
            finally:
                self._pending_psfs = None
            return out[None] if self._B == 1 else out
        if psfs is not None:
            self._set_psf(psfs[0] if psfs.dim() == 5 else psfs)
        elif self._per_sample:
            self._plan, self._per_sample, self._psfs = None, False, None
        self.set_data(batch)
        out = self.apply(background=background, **apply_kw)
        return out[None] if self._B == 1 else out

    # ------------------------------------------------------------------ metrics
    def reconstruction_error(self, prediction=None, lensless=None, psfs=None, normalize=True):
        """recon.py:607-653, on the device in one engine call: convolve (pad -> rfft2 -> H -> irfft2 -> crop), then the
        min/max normalisation per (batch, depth) and the squared-error sum per batch entry as two fused reduction
        kernels (csrc/lpc_metrics.cuh).  ``psfs`` may be one PSF or one per sample ``[B, D, h, w, C]``."""
        if prediction is None:
            prediction = self.get_image_estimate()
        if lensless is None:
            lensless = self._data
        if psfs is None:
            psfs = self._psfs if (self._per_sample and self._psfs is not None) else self._psf
        if psfs is self._psf:
            convolver = self._convolver  # cached plan of the solver's own PSF (a benchmark loop calls this per sample)
        else:
            if self._err_convolver is None:
                kw = {k: v for k, v in self._convolver_param.items() if k not in ("no_graph", "engine")}
                kw["pad"] = True
                self._err_convolver = RealFFTConvolve2D(psfs.to(prediction.device), **kw)
            else:
                self._err_convolver.set_psf(psfs.to(prediction.device))
            convolver = self._err_convolver
        return convolver.reconstruction_error(prediction, lensless, self._npix, normalize=normalize)

    # ------------------------------------------------------------------ host-buffer path (bench e2e)
    def apply_host(self, host_data, host_out, n_iter=None):
        """End-to-end call with HOST buffers: H2D copy, reset, n_iter iterations, crop/clamp,
        D2H copy -- all enqueued on one stream; the caller synchronises.  ``host_data`` is
        ``[B, Dd, h, w, C]`` (pinned recommended), ``host_out`` ``[B, D, h, w, C]``."""
        assert host_data.dim() == 5 and not host_data.is_cuda and not host_out.is_cuda
        assert host_data.dtype == self._dtype and host_out.dtype == self._dtype
        assert host_data.is_contiguous() and host_out.is_contiguous()
        B, Dd = int(host_data.shape[0]), int(host_data.shape[1])
        plan = self._select_plan(B)
        lib = _lib.load()
        if n_iter is None:
            n_iter = self._n_iter
        with self._stream as st:
            _lib.check(lib.lpc_set_data_host(plan, host_data.data_ptr(), Dd, st), plan)
            _lib.check(lib.lpc_reset(plan, None, st), plan)
            _lib.check(lib.lpc_run(plan, int(n_iter), st), plan)
            _lib.check(lib.lpc_get_image_host(plan, host_out.data_ptr(), st), plan)
        self._data_sent = False
        return host_out

    def apply_host_stream(self, host_inputs, host_outputs, n_iter=None):
        """Reconstruct a STREAM of host-resident measurements: ``host_inputs[k]`` (pinned ``[B, Dd, h, w, C]``) ->
        ``host_outputs[k]`` (pinned ``[B, D, h, w, C]``).  Every measurement still pays its own H2D copy, reset,
        ``n_iter`` iterations, crop/clamp and D2H copy, but the copies run on two side streams against double-buffered
        device staging, so the copy of measurement k+1 and the read-back of k-1 overlap the iterations of k.
        Returns after the last image has landed in host memory."""
        assert len(host_inputs) == len(host_outputs) and len(host_inputs) > 0
        if n_iter is None:
            n_iter = self._n_iter
        first = host_inputs[0]
        assert first.dim() == 5 and not first.is_cuda and first.dtype == self._dtype
        B, Dd = int(first.shape[0]), int(first.shape[1])
        D, h, w, C = (int(v) for v in self._psf.shape)
        plan = self._select_plan(B)
        lib = _lib.load()
        dev = self._device
        if getattr(self, "_hs", None) is None or self._hs["key"] != (B, Dd):
            self._hs = {
                "key": (B, Dd),
                "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev),
                "din": [torch.empty((B, Dd, h, w, C), dtype=self._dtype, device=dev) for _ in range(2)],
                "dout": [torch.empty((B, D, h, w, C), dtype=self._dtype, device=dev) for _ in range(2)],
            }
        hs = self._hs
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]
        out_ready = [torch.cuda.Event() for _ in range(2)]
        out_free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        for ev in in_free + out_free:
            ev.record(cur)
        hs["h2d"].wait_stream(cur)
        hs["d2h"].wait_stream(cur)
        for k, (hin, hout) in enumerate(zip(host_inputs, host_outputs)):
            s = k & 1
            with torch.cuda.stream(hs["h2d"]):
                hs["h2d"].wait_event(in_free[s])
                hs["din"][s].copy_(hin, non_blocking=True)
                in_ready[s].record(hs["h2d"])
            cur.wait_event(in_ready[s])
            with self._stream as st:
                _lib.check(lib.lpc_set_data(plan, hs["din"][s].data_ptr(), Dd, None, None, None, None, st), plan)
            in_free[s].record(cur)
            cur.wait_event(out_free[s])
            with self._stream as st:
                _lib.check(lib.lpc_reset(plan, None, st), plan)
                _lib.check(lib.lpc_run(plan, int(n_iter), st), plan)
                _lib.check(lib.lpc_get_image(plan, hs["dout"][s].data_ptr(), st), plan)
            out_ready[s].record(cur)
            with torch.cuda.stream(hs["d2h"]):
                hs["d2h"].wait_event(out_ready[s])
                hout.copy_(hs["dout"][s], non_blocking=True)
                out_free[s].record(hs["d2h"])
        hs["d2h"].synchronize()
        self._data_sent = False
        return host_outputs

    def bytes_per_iter(self):
        """Algorithmic HBM bytes one iteration moves for the current batch."""
        lib = _lib.load()
        v = ctypes.c_double()
        _lib.check(lib.lpc_bytes_per_iter(self._ensure_plan(), ctypes.byref(v)), self._plan)
        return v.value


def _load_drunet_function(device):
    """recon.py:317-319: ``denoiser["network"] == "DruNet"``.  The weights come from the reference's model hub
    (lensless.recon.utils.load_drunet); without the reference package and network access this raises."""
    try:
        from lensless.recon.utils import get_drunet_function_v2, load_drunet
    except Exception as e:  # pragma: no cover
        raise NotImplementedError(
            'denoiser["network"] == "DruNet" needs the reference package (lensless.recon.utils.load_drunet) and its '
            "downloaded weights; pass a callable f(image, noise_level) instead"
        ) from e
    model = load_drunet(requires_grad=False).to(device)
    return get_drunet_function_v2(model, mode="inference")


def _import_plotting():
    try:
        import matplotlib.pyplot as plt
        from lensless.utils.plot import plot_image
    except Exception as e:  # pragma: no cover - matplotlib is not in the image
        raise ImportError(
            "plot=True / save=... need matplotlib and the reference's lensless.utils.plot.plot_image "
            "(visualisation is out of the engine's scope)"
        ) from e
    return plot_image, plt
