"""
Inference-only GPU mirrors of the reference's unrolled solvers:

* ``UnrolledADMM``  -- lensless/recon/unrolled_admm.py:20-251: ADMM whose mu1 / mu2 / mu3 / tau are
  vectors with one entry per iteration (``_mu1_p`` ...; their absolute value is used,
  unrolled_admm.py:141-144), ``R_divmat`` / ``X_divmat`` per iteration (unrolled_admm.py:166-179).
* ``UnrolledFISTA`` -- lensless/recon/unrolled_fista.py:19-106: projected gradient descent with a step
  ``alpha[i][c]`` per iteration and channel and the momentum weights ``(tk[i] - 1) / tk[i + 1]``.

``forward(batch, psfs=None, background=None)`` has the shape contract of
``TrainableReconstructionAlgorithm.forward`` (trainable_recon.py:297-405): ``[B, D|1, h, w, C]`` in,
``[B, D, h, w, C]`` out.  The same fused kernels run as for ADMM / FISTA, with the parameters of
iteration ``i`` passed as kernel arguments (``lpc_set_admm_schedule`` / ``lpc_set_fista_schedule``).
There is no autograd: learned parameter vectors are loaded (``set_parameters``), not trained.
Pre-/post-processors, compensation branches and PSF networks are torch modules outside the hot path
and raise ``NotImplementedError`` when requested.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .admm import ADMM
from .gd import FISTA, non_neg

_UNSUPPORTED = (
    "pre_process", "post_process", "background_network", "psf_network", "compensation",
    "integrated_background_subtraction", "learned_background_subtraction", "unetres_input_background",
)


def _vec(value, n, name):
    """scalar or length-n sequence / tensor -> float64 numpy vector of length n"""
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu().numpy()
    arr = np.asarray(value, dtype=np.float64)
    if arr.ndim == 0:
        arr = np.full((n,), float(arr))
    assert arr.shape == (n,), f"{name} must be a scalar or have one entry per iteration ({n})"
    return arr


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _reject_unsupported(kwargs):
    for key in _UNSUPPORTED:
        if kwargs.pop(key, None):
            raise NotImplementedError(f"`{key}` is a torch module outside the fused hot path (SURVEY 8f)")
    for key in ("skip_unrolled", "skip_pre", "skip_post", "return_intermediate", "legacy_denoiser",
                "compensation_residual", "psf_residual"):
        kwargs.pop(key, None)
    return kwargs


class _UnrolledMixin:
    direct_background_subtraction = False

    def _prepare_batch(self, batch, background):
        """trainable_recon.py:313-321: data = clamp(data - background, 0, 1) when direct subtraction is on"""
        assert isinstance(batch, torch.Tensor) and len(batch.shape) == 5, "batch must be of shape (N, D, H, W, C)"
        if self.direct_background_subtraction:
            assert background is not None, "If direct_background_subtraction is True, background must be defined."
            batch = torch.clamp(batch - background.to(batch.device), 0, 1)
        return batch

    def forward(self, batch, psfs=None, background=None):
        """trainable_recon.py:297-405 without processors: reset, n_iter updates, _form_image."""
        batch = self._prepare_batch(batch, background)
        if psfs is not None:
            assert psfs.shape[-3:] == batch.shape[-3:], "psfs must have the same shape as batch"
        return self._forward_batch(batch, psfs, None, n_iter=self._n_iter)

    batch_call = forward  # trainable_recon.py: alias kept for callers


class UnrolledADMM(_UnrolledMixin, ADMM):
    def __init__(self, psf, dtype=None, n_iter=5, mu1=1e-6, mu2=1e-5, mu3=4e-5, tau=0.0001, psi=None,
                 psi_adj=None, psi_gram=None, pad=False, norm="backward", direct_background_subtraction=False,
                 **kwargs):
        kwargs = _reject_unsupported(dict(kwargs))
        self.direct_background_subtraction = bool(direct_background_subtraction)
        # unrolled_admm.py:80-99: one entry per iteration
        self._mu1_p = torch.as_tensor(_vec(mu1, n_iter, "mu1"))
        self._mu2_p = torch.as_tensor(_vec(mu2, n_iter, "mu2"))
        self._mu3_p = torch.as_tensor(_vec(mu3, n_iter, "mu3"))
        self._tau_p = torch.as_tensor(_vec(tau, n_iter, "tau"))
        super().__init__(psf, dtype, mu1=float(abs(self._mu1_p[0])), mu2=float(abs(self._mu2_p[0])),
                         mu3=float(abs(self._mu3_p[0])), tau=float(abs(self._tau_p[0])), psi=psi, psi_adj=psi_adj,
                         psi_gram=psi_gram, pad=pad, norm=norm, n_iter=n_iter, **kwargs)

    def set_parameters(self, mu1=None, mu2=None, mu3=None, tau=None):
        """Load (learned) per-iteration parameter vectors; takes effect at the next reset()."""
        n = self._n_iter
        for name, val in (("_mu1_p", mu1), ("_mu2_p", mu2), ("_mu3_p", mu3), ("_tau_p", tau)):
            if val is not None:
                setattr(self, name, torch.as_tensor(_vec(val, n, name)))

    def _push_params(self, plan, stream):
        # constant-parameter slots = entry 0 (self._mu1 ... are per-iteration tensors here, as in the reference)
        first = [float(abs(p[0])) for p in (self._mu1_p, self._mu2_p, self._mu3_p, self._tau_p)]
        lib = _lib.load()
        _lib.check(lib.lpc_set_admm_params(plan, *first, stream), plan)
        self._push_schedule(plan, stream)

    def _push_schedule(self, plan, stream):
        # unrolled_admm.py:141-144: the absolute values are what the iteration uses
        vecs = [np.abs(_vec(p, self._n_iter, "parameter")) for p in (self._mu1_p, self._mu2_p, self._mu3_p, self._tau_p)]
        self._mu1, self._mu2, self._mu3, self._tau = (torch.as_tensor(v) for v in vecs)
        key = b"".join(v.tobytes() for v in vecs)
        pushed = self.__dict__.setdefault("_pushed", {})
        if pushed.get(plan.value) == key:
            return
        lib = _lib.load()
        _lib.check(lib.lpc_set_admm_schedule(plan, self._n_iter, *[_dptr(v) for v in vecs], stream), plan)
        pushed[plan.value] = key

    def reset(self, batch_size=1):
        with self._stream as st:
            for plan in self._plans.values():
                self._push_schedule(plan, st)
        super().reset()


class UnrolledFISTA(_UnrolledMixin, FISTA):
    _keep_x0 = True  # forward(psfs=...) swaps the convolver only; x0 stays the constructor PSF's (unrolled_fista.py:55-59)

    def _after_new_plan(self, plan, stream):
        self._push_schedule(plan, stream)

    def __init__(self, psf, n_iter=5, dtype=None, proj=non_neg, learn_tk=True, tk=1, direct_background_subtraction=False,
                 **kwargs):
        kwargs = _reject_unsupported(dict(kwargs))
        self.direct_background_subtraction = bool(direct_background_subtraction)
        self._alpha_p = None  # filled from the engine's 1.8 / max|H*H| after the PSF is set (unrolled_fista.py:60-73)
        # unrolled_fista.py:75-81
        tks = [float(tk)]
        for i in range(n_iter):
            tks.append((1 + math.sqrt(1 + 4 * tks[i] ** 2)) / 2)
        self._tk_p = torch.as_tensor(np.asarray(tks, dtype=np.float64))
        self._schedule_ready = False
        super().__init__(psf, dtype, proj, tk=float(tk), n_iter=n_iter, **kwargs)

    def set_parameters(self, alpha=None, tk=None):
        """Load (learned) step sizes ``alpha[n_iter, C]`` and momentum schedule ``tk[n_iter + 1]``."""
        C = int(self._psf.shape[3])
        if alpha is not None:
            a = alpha.detach().cpu().numpy() if isinstance(alpha, torch.Tensor) else np.asarray(alpha)
            a = np.asarray(a, dtype=np.float64)
            assert a.shape == (self._n_iter, C), f"alpha must be [{self._n_iter}, {C}]"
            self._alpha_p = torch.as_tensor(a)
        if tk is not None:
            t = tk.detach().cpu().numpy() if isinstance(tk, torch.Tensor) else np.asarray(tk)
            t = np.asarray(t, dtype=np.float64)
            assert t.shape == (self._n_iter + 1,), f"tk must have {self._n_iter + 1} entries"
            self._tk_p = torch.as_tensor(t)

    def _default_alpha(self):
        # unrolled_fista.py:60-73: computed ONCE from the constructor PSF (a shared-PSF plan of batch size 1)
        C = int(self._psf.shape[3])
        out = torch.empty((C,), dtype=self._dtype, device=self._device)
        lib = _lib.load()
        keep = (self._plan, self._B, self._per_sample)
        plan = self._select_plan(1)
        with self._stream as st:
            _lib.check(lib.lpc_get_alpha(plan, out.data_ptr(), st), plan)
        self._plan, self._B, self._per_sample = keep
        a = out.double().cpu().numpy()  # lip_fact / max|H*H| with lip_fact = 1.8 (gd.py:108-112 == unrolled_fista.py:63-66)
        return torch.as_tensor(np.tile(a[None, :], (self._n_iter, 1)))

    def _push_schedule(self, plan, stream):
        if self._alpha_p is None:
            self._alpha_p = self._default_alpha()
        alpha = np.ascontiguousarray(np.abs(self._alpha_p.numpy()), dtype=np.float64)  # unrolled_fista.py:99-100
        tks = np.ascontiguousarray(np.abs(self._tk_p.numpy()), dtype=np.float64)
        self._alpha_sched, self._tk_sched = torch.as_tensor(alpha), torch.as_tensor(tks)
        key = alpha.tobytes() + tks.tobytes()
        pushed = self.__dict__.setdefault("_pushed", {})
        if pushed.get(plan.value) == key:
            return
        lib = _lib.load()
        _lib.check(lib.lpc_set_fista_schedule(plan, self._n_iter, _dptr(alpha), _dptr(tks), stream), plan)
        pushed[plan.value] = key

    def reset(self, batch_size=1):
        with self._stream as st:
            for plan in self._plans.values():
                self._push_schedule(plan, st)
        super().reset()
