"""
GPU mirrors of ``lensless.GradientDescent`` / ``NesterovGradientDescent`` / ``FISTA``
(reference: lensless/recon/gd.py:62-241), projection = non-negativity.

Per iteration the fused engine (fp32, menu sizes) runs four hand-written kernels that own their FFTs:
columns(H) -> rows(crop, - y, zero-pad) -> columns(conj H) -> rows(step + projection + momentum); only the
window rows of the zero-padded problem are ever transformed along rows, stored or loaded
(lenslesspicam_b200/csrc/lpc_fused_gd.cuh, lpc_fused_tma.cuh).  Other sizes / fp64 run cuFFT + four
kernels (lpc_gd.cuh).
"""
import inspect

from . import _lib
from .recon import ReconstructionAlgorithm


class GradientDescentUpdate:
    """gd.py:24-38"""

    VANILLA = "vanilla"
    NESTEROV = "nesterov"
    FISTA = "fista"

    @staticmethod
    def all_values():
        vals = []
        for i in inspect.getmembers(GradientDescentUpdate):
            if not i[0].startswith("_") and not callable(i[1]):
                vals.append(i[1])
        return vals


def non_neg(xi):
    """gd.py:41-59 -- the one projection the engine fuses (marker object for ``proj=``)."""
    import torch

    return torch.clamp(xi, min=0)


class GradientDescent(ReconstructionAlgorithm):
    _ALGO = _lib.ALGO_GD

    def __init__(self, psf, dtype=None, proj=non_neg, lip_fact=1.8, **kwargs):
        assert callable(proj)
        self._proj = proj
        if proj is not non_neg and getattr(proj, "__name__", "") != "non_neg" and kwargs.get("denoiser") is None:
            # only the non-negativity projection is fused into the update kernels; any other callable runs between the
            # two halves of the split iteration, like a plug-and-play denoiser (csrc/lpc_pnp.cuh)
            kwargs["denoiser"] = {"network": (lambda x, _noise_level, _p=proj: _p(x)), "noise_level": None}
        self._lip_fact = lip_fact
        kwargs.setdefault("pad", True)
        if not kwargs["pad"]:
            raise NotImplementedError("the gradient-descent family optimises the unpadded image (pad=True)")
        if not hasattr(self, "_nesterov_mu"):
            self._nesterov_mu = 0.9
        if not hasattr(self, "_initial_tk"):
            self._initial_tk = 1.0
        super(GradientDescent, self).__init__(psf, dtype, **kwargs)

    def _push_params(self, plan, stream):
        lib = _lib.load()
        _lib.check(lib.lpc_set_gd_params(plan, self._lip_fact, self._nesterov_mu, self._initial_tk, stream), plan)

    @property
    def _alpha(self):
        """step sizes per channel (gd.py:108-112)"""
        import torch

        out = torch.empty((int(self._psf.shape[3]),), dtype=self._dtype, device=self._device)
        lib = _lib.load()
        with self._stream as st:
            _lib.check(lib.lpc_get_alpha(self._plan, out.data_ptr(), st), self._plan)
        return out


class NesterovGradientDescent(GradientDescent):
    """gd.py:143-188.  As in the reference, ``reset()`` always restores p=0 and mu=0.9, so the
    constructor's ``p`` / ``mu`` are never honoured (SURVEY 8b quirk v) -- mirrored."""

    _ALGO = _lib.ALGO_NESTEROV

    def __init__(self, psf, dtype=None, proj=non_neg, p=0, mu=0.9, **kwargs):
        self._p = 0
        self._mu = 0.9
        self._nesterov_mu = 0.9
        super(NesterovGradientDescent, self).__init__(psf, dtype, proj, **kwargs)

    def reset(self, p=0, mu=0.9):
        self._p = p
        self._mu = mu
        if mu != self._nesterov_mu:
            self._nesterov_mu = mu
            for plan in self._plans.values():
                with self._stream as st:
                    self._push_params(plan, st)
        if p != 0:
            raise NotImplementedError("a non-zero initial momentum buffer is not supported")
        super(NesterovGradientDescent, self).reset()


class FISTA(GradientDescent):
    """gd.py:191-241"""

    _ALGO = _lib.ALGO_FISTA

    def __init__(self, psf, dtype=None, proj=non_neg, tk=1.0, **kwargs):
        self._initial_tk = tk
        super(FISTA, self).__init__(psf, dtype, proj, **kwargs)
        self._tk = tk

    def reset(self, tk=None):
        if tk and tk != self._initial_tk:
            # gd.py:227-233: an explicit tk only applies to this reset; the plan keeps one start value
            raise NotImplementedError("reset(tk=...) differing from the constructor's tk is not supported")
        super(FISTA, self).reset()
        self._tk = self._initial_tk


def apply_gradient_descent(psf_fp, data_fp, n_iter, verbose=False, proj=non_neg, **kwargs):
    """gd.py:244-263.  ``psf_fp`` / ``data_fp`` are file paths as in the reference (loaded by the GPU loader,
    ``lenslesspicam_b200.io.load_data``, with ``**kwargs``) or already-loaded CUDA tensors."""
    import time

    import torch

    if isinstance(psf_fp, str):
        from .io import load_data

        psf, data = load_data(psf_fp=psf_fp, data_fp=data_fp, plot=False, **kwargs)
        kwargs = {}
    else:
        psf, data = psf_fp, data_fp
    recon = GradientDescent(psf, n_iter=n_iter, proj=proj, **kwargs)
    recon.set_data(data)
    torch.cuda.synchronize()
    start_time = time.time()
    res = recon.apply(plot=False)
    torch.cuda.synchronize()
    proc_time = time.time() - start_time
    if verbose:
        print(f"Reconstruction time : {proc_time} s")
        print(f"Reconstruction shape: {res.shape}")
    return res
