"""
GPU mirror of ``lensless.recon.rfft_convolve.RealFFTConvolve2D``
(reference: lensless/recon/rfft_convolve.py:26-223).

Same constructor, attributes and methods; the arithmetic runs in liblpc.so
(zero-pad + cuFFT R2C, one spectral-multiply kernel that also applies the
ifftshift as a phase, cuFFT C2R, crop) on planar device buffers.  CUDA tensors
only -- there is no CPU path in this package.
"""
import numpy as np
import torch

from . import _lib
from ._engine import EngineStream, torch_dtype_code


class RealFFTConvolve2D:
    def __init__(self, psf, dtype=None, pad=True, norm="ortho", rgb=None, **kwargs):
        # rfft_convolve.py:49-77 (unknown kwargs are swallowed there too, :27)
        if not isinstance(psf, torch.Tensor) or not psf.is_cuda:
            raise TypeError(
                "lenslesspicam_b200.RealFFTConvolve2D needs a CUDA torch.Tensor PSF "
                "(numpy / CPU tensors are the reference's own path; no CPU fallback here)"
            )
        self.is_torch = True
        assert len(psf.shape) >= 4, "Expected 4D PSF of shape ([batch], depth, width, height, channels)"
        if len(psf.shape) > 5:
            raise ValueError("Expected a 4D or 5D PSF")
        self._use_3d = psf.shape[-4] != 1
        self._is_rgb = (psf.shape[-1] == 3) if rgb is None else rgb
        assert self._is_rgb or psf.shape[-1] == 1
        if norm not in _lib.NORMS:
            raise ValueError(f"Invalid norm: {norm}")
        self.norm = norm
        if dtype is None:
            dtype = torch.float32
        if isinstance(dtype, str):
            dtype = {"float32": torch.float32, "float64": torch.float64}[dtype]
        self.dtype = dtype
        self.pad = pad
        self._plans = {}
        self._stream = EngineStream(psf.device)
        self.set_psf(psf)

    # ------------------------------------------------------------------ geometry
    def set_psf(self, psf):
        # rfft_convolve.py:102-131
        # a 5-D PSF [B, D, h, w, C] is one PSF per batch entry (rfft_convolve.py:121: H keeps the batch axis);
        # the engine holds it in a plan created with LPC_FLAG_PSF_PER_SAMPLE
        psf = psf.type(self.dtype).contiguous()
        n_psf = int(psf.shape[0]) if psf.dim() == 5 else 0
        if n_psf == 1:
            psf, n_psf = psf[0], 0
        if n_psf != getattr(self, "_n_psf", n_psf):
            self._drop_plans()
        self._n_psf = n_psf
        self._psf = psf
        self._psf_shape = np.array(self._psf.shape)
        lib = _lib.load()
        import ctypes

        Hp, Wp, sy, sx = (ctypes.c_int() for _ in range(4))
        h, w = int(self._psf_shape[-3]), int(self._psf_shape[-2])
        _lib.check(lib.lpc_geometry(h, w, Hp, Wp, sy, sx))
        C = 3 if self._is_rgb else 1
        self._padded_shape = [int(self._psf_shape[-4]), Hp.value, Wp.value, C]
        self._start_idx = np.array([sy.value, sx.value])
        self._end_idx = self._start_idx + self._psf_shape[-3:-1]
        for plan in self._plans.values():
            self._set_plan_psf(plan)
        self._H_cache = None

    def _set_plan_psf(self, plan):
        lib = _lib.load()
        with self._stream as st:
            _lib.check(lib.lpc_set_psf(plan, self._psf.data_ptr(), st), plan)

    def _plan(self, B):
        plan = self._plans.get(B)
        if plan is None:
            import ctypes

            lib = _lib.load()
            plan = ctypes.c_void_p()
            D, h, w, C = (int(v) for v in self._psf.shape[-4:])
            if self._n_psf and B != self._n_psf:
                raise ValueError(f"{self._n_psf} PSFs for a batch of {B}: one PSF per sample is needed")
            while len(self._plans) >= 2:  # bounded: a plan holds three padded arrays per image
                lib.lpc_plan_destroy(self._plans.pop(next(iter(self._plans))))
            _lib.check(
                lib.lpc_plan_create(
                    ctypes.byref(plan), self._psf.device.index or 0, _lib.ALGO_CONV, torch_dtype_code(self.dtype),
                    _lib.NORMS[self.norm], B, D, C, h, w, _lib.FLAG_PSF_PER_SAMPLE if self._n_psf else _lib.FLAG_NONE,
                )
            )
            self._plans[B] = plan
            self._set_plan_psf(plan)
        return plan

    def _drop_plans(self):
        try:
            lib = _lib.load()
            for plan in getattr(self, "_plans", {}).values():
                lib.lpc_plan_destroy(plan)
        except Exception:
            pass
        self._plans = {}

    def __del__(self):
        self._drop_plans()

    @property
    def _H(self):
        """complex [D, Hp, Wp//2+1, C], as rfft_convolve.py:121-123 stores it."""
        if self._H_cache is None:
            lib = _lib.load()
            D, Hp, Wp, C = self._padded_shape
            cdt = torch.complex64 if self.dtype == torch.float32 else torch.complex128
            lead = (self._n_psf,) if self._n_psf else ()
            out = torch.empty(lead + (D, Hp, Wp // 2 + 1, C), dtype=cdt, device=self._psf.device)
            plan = self._plan(self._n_psf or 1)
            with self._stream as st:
                _lib.check(lib.lpc_get_H(plan, out.data_ptr(), st), plan)
            self._H_cache = out
        return self._H_cache

    @property
    def _Hadj(self):
        return torch.conj(self._H)

    # ------------------------------------------------------------------ pad / crop
    def _crop(self, x):
        # rfft_convolve.py:79-82 (a view, like the reference)
        return x[..., self._start_idx[0] : self._end_idx[0], self._start_idx[1] : self._end_idx[1], :]

    def _pad(self, v):
        # rfft_convolve.py:84-100: pure indexing (a copy into a zeroed array), no arithmetic
        if len(v.shape) == 5:
            shape = [v.shape[0]] + self._padded_shape
        elif len(v.shape) == 4:
            shape = self._padded_shape
        else:
            raise ValueError("Expected 4D or 5D tensor")
        vpad = torch.zeros(size=shape, dtype=v.dtype, device=v.device)
        vpad[..., self._start_idx[0] : self._end_idx[0], self._start_idx[1] : self._end_idx[1], :] = v
        return vpad

    # ------------------------------------------------------------------ operator
    def _apply(self, x, adjoint):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise TypeError("expected a CUDA tensor")
        squeeze = False
        if x.dim() == 4:
            x = x[None]
            squeeze = True
        if x.dim() != 5:
            raise ValueError("Expected 4D or 5D tensor")
        D, Hp, Wp, C = self._padded_shape
        hh, ww = (int(self._psf.shape[-3]), int(self._psf.shape[-2])) if self.pad else (Hp, Wp)
        assert tuple(x.shape[-3:-1]) == (hh, ww), "input spatial shape does not match the convolver"
        B = int(x.shape[0])
        if self._n_psf and B == 1:
            B = self._n_psf  # one input against a batch of PSFs broadcasts over the batch (torch semantics)
        if x.shape[1] not in (1, D) or x.shape[4] not in (1, C):
            raise ValueError("depth / channel axes must be 1 or match the PSF")
        x = x.to(self.dtype)
        xe = x.expand(B, D, hh, ww, C)  # stride-0 broadcast of size-1 depth / channel axes
        out = torch.empty((B, D, hh, ww, C), dtype=self.dtype, device=x.device)
        lib = _lib.load()
        plan = self._plan(B)
        with self._stream as st:
            _lib.check(
                lib.lpc_convolve(
                    plan, xe.data_ptr(), _lib.strides_arg(xe.stride()), out.data_ptr(),
                    1 if adjoint else 0, 1 if self.pad else 0, st,
                ),
                plan,
            )
        assert out.shape[-3:-1] == x.shape[-3:-1]
        return out[0] if squeeze else out

    def convolve(self, x, return_fft=False):
        """rfft_convolve.py:133-176"""
        if return_fft:
            raise NotImplementedError("return_fft is not exposed by the fused engine")
        return self._apply(x, adjoint=False)

    def deconvolve(self, y, return_fft=False):
        """rfft_convolve.py:178-223"""
        if return_fft:
            raise NotImplementedError("return_fft is not exposed by the fused engine")
        return self._apply(y, adjoint=True)

    # ------------------------------------------------------------------ metric
    def reconstruction_error(self, prediction, lensless, npix, normalize=True):
        """The arithmetic of ``ReconstructionAlgorithm.reconstruction_error`` (recon.py:607-653) in one call of the
        engine: pad -> convolve -> crop (with ``pad=False`` the reference pads the prediction itself and crops the
        result, recon.py:633-638: the same numbers), min/max normalisation per (batch, depth) and the squared-error
        sum against the measurement as two fused reduction kernels.  Returns ``[B]``."""
        D, Hp, Wp, C = self._padded_shape
        h, w = int(self._psf.shape[-3]), int(self._psf.shape[-2])
        prediction = prediction.to(self.dtype)
        lensless = lensless.to(device=prediction.device, dtype=self.dtype)
        while prediction.dim() < 5:
            prediction = prediction[None]
        while lensless.dim() < 5:
            lensless = lensless[None]
        B = int(prediction.shape[0])
        assert tuple(prediction.shape[1:]) == (D, h, w, C), "prediction must be [B, D, h, w, C]"
        Dl = int(lensless.shape[1])
        assert Dl in (1, D) and tuple(lensless.shape[2:]) == (h, w, C)
        lensless = lensless.expand(B, Dl, h, w, C).contiguous()
        prediction = prediction.contiguous()
        out = torch.empty((B,), dtype=self.dtype, device=prediction.device)
        lib = _lib.load()
        plan = self._plan(B)
        with self._stream as st:
            _lib.check(
                lib.lpc_reconstruction_error(
                    plan, prediction.data_ptr(), lensless.data_ptr(), Dl, 1 if normalize else 0, float(npix),
                    out.data_ptr(), st,
                ),
                plan,
            )
        return out
