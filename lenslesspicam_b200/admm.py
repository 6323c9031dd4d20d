"""
GPU mirror of ``lensless.ADMM`` (reference: lensless/recon/admm.py:24-338):
ADMM with an anisotropic-TV prior and a non-negativity constraint.

Per iteration the fused engine (default, fp32, padded sizes in the compiled menu) runs THREE hand-written
kernels that own their FFTs (the reference: ~360 torch op dispatches and six FFTs):

    B1  row IFFT of the (V, HV) row spectra                            -> V, HV
    B2  fused spatial pass -- dual completion, U/X/W proxes, new half-duals, FFT operand a
        (admm.py:232-263, 277-279, 298-311) -> row FFT                 -> row spectra of (a, xi')
    A   persistent TMA-pipelined column kernel: column FFT -> Vhat = R (A - S conj(H) Xi'),
        HVhat = S H Vhat (admm.py:280-294, 320; R evaluated on the fly) -> column IFFT

Other sizes / fp64 / ``engine="v1"`` run K1 (spatial) + cuFFT + K2 (spectral).
See lenslesspicam_b200/csrc/lpc_fused.cuh, lpc_fused_tma.cuh, lpc_admm.cuh and DESIGN.md.
"""
import time

import torch

from . import _lib
from .recon import ReconstructionAlgorithm


class ADMM(ReconstructionAlgorithm):
    _ALGO = _lib.ALGO_ADMM

    def __init__(
        self,
        psf,
        dtype=None,
        mu1=1e-6,
        mu2=1e-5,
        mu3=4e-5,
        tau=0.0001,
        psi=None,
        psi_adj=None,
        psi_gram=None,
        pad=False,
        norm="backward",
        # PnP
        denoiser=None,
        **kwargs,
    ):
        # admm.py:86-89
        self._mu1 = mu1
        self._mu2 = mu2
        self._mu3 = mu3
        self._tau = tau

        # admm.py:91-96 -- the reference has no 3-D ADMM.  ``allow_depth=True`` opts in to the
        # engine's extension: D independent 2-D problems against the same measurement, the
        # semantics of the reference's own 3-D gradient descent (SURVEY 8a note).
        assert len(psf.shape) == 4, "PSF must be 4D: (depth, height, width, channels)."
        if psf.shape[0] > 1 and not kwargs.pop("allow_depth", False):
            raise NotImplementedError("3D ADMM is not supported yet, use gradient descent or APGD instead.")
        if psi is not None or psi_adj is not None or psi_gram is not None:
            raise NotImplementedError("custom psi / psi_adj / psi_gram cannot be fused; only the TV prior is built")
        if pad:
            raise NotImplementedError("ADMM optimises the padded image (pad=False), as in the reference default")

        super(ADMM, self).__init__(psf, dtype, pad=pad, norm=norm, denoiser=denoiser, reset=False, **kwargs)
        self.reset()  # admm.py:135

    def _push_params(self, plan, stream):
        lib = _lib.load()
        _lib.check(lib.lpc_set_admm_params(plan, self._mu1, self._mu2, self._mu3, self._tau, stream), plan)


# ---------------------------------------------------------------------------------------------
# module-level helpers (admm.py:341-397; unrolled_admm.py:10 and user code import them).  Inside the solvers they are
# fused into the spatial kernels; these are the stand-alone forms, each one hand-written kernel (csrc/lpc_ops.cuh).
# ---------------------------------------------------------------------------------------------
def _op_args(x):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("lenslesspicam_b200 helpers need CUDA tensors (the reference's own functions serve numpy / CPU)")
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("float32 / float64 only")
    x = x.contiguous()
    from ._engine import torch_dtype_code

    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    return x, dev, torch_dtype_code(x.dtype), torch.cuda.current_stream(x.device).cuda_stream


def soft_thresh(x, thresh):
    """admm.py:341-346: ``sign(x) * max(|x| - thresh, 0)``"""
    x, dev, code, st = _op_args(x)
    out = torch.empty_like(x)
    _lib.check(_lib.load().lpc_op_soft_thresh(dev, code, x.data_ptr(), out.data_ptr(), x.numel(), float(thresh), st))
    return out


def finite_diff(x):
    """admm.py:349-359: periodic backward differences along H (-3) and W (-2), stacked on a new last axis."""
    x, dev, code, st = _op_args(x)
    assert x.dim() >= 3, "expected [..., H, W, C]"
    H, W, C = (int(v) for v in x.shape[-3:])
    outer = x.numel() // max(H * W * C, 1)
    out = torch.empty(tuple(x.shape) + (2,), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().lpc_op_finite_diff(dev, code, x.data_ptr(), out.data_ptr(), outer, H, W, C, st))
    return out


def finite_diff_adj(x):
    """admm.py:362-370: adjoint of :func:`finite_diff` -- ``[..., H, W, C, 2] -> [..., H, W, C]``"""
    x, dev, code, st = _op_args(x)
    assert x.dim() >= 4 and x.shape[-1] == 2, "expected [..., H, W, C, 2]"
    H, W, C = (int(v) for v in x.shape[-4:-1])
    outer = x.numel() // max(2 * H * W * C, 1)
    out = torch.empty(tuple(x.shape[:-1]), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().lpc_op_finite_diff_adj(dev, code, x.data_ptr(), out.data_ptr(), outer, H, W, C, st))
    return out


def finite_diff_gram(shape, dtype=None, is_torch=True, device=None):
    """admm.py:373-397: ``rfft2`` of the finite-difference Gram stencil for a padded shape ``[D, Hp, Wp, C]``, evaluated in
    closed form (``4 - 2cos(2 pi ky/Hp) - 2cos(2 pi kx/Wp)``; the reference's depth > 1 variant included).  Complex
    ``[D, Hp, Wp//2+1, C]`` on ``device`` (default: the current CUDA device; the reference returns a CPU tensor that
    ``ADMM.__init__`` then moves to the PSF's device, admm.py:122-123)."""
    if not is_torch:
        raise TypeError("no numpy path in lenslesspicam_b200")
    from .recon import get_dtype
    from ._engine import torch_dtype_code

    if dtype is None or isinstance(dtype, str):
        dtype = get_dtype(dtype, True)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    D, Hp, Wp, C = (int(v) for v in shape)
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    out = torch.empty((D, Hp, Wp // 2 + 1, C), dtype=cdt, device=device)
    dev = device.index if device.index is not None else torch.cuda.current_device()
    st = torch.cuda.current_stream(device).cuda_stream
    _lib.check(_lib.load().lpc_op_finite_diff_gram(dev, torch_dtype_code(dtype), out.data_ptr(), D, Hp, Wp, C, st))
    return out


def apply_admm(psf_fp, data_fp, n_iter, verbose=False, **kwargs):
    """admm.py:400-419.  ``psf_fp`` / ``data_fp`` are file paths as in the reference (loaded by the GPU loader,
    ``lenslesspicam_b200.io.load_data``, with ``**kwargs``) or already-loaded CUDA tensors."""
    if isinstance(psf_fp, str):
        from .io import load_data

        psf, data = load_data(psf_fp=psf_fp, data_fp=data_fp, plot=False, **kwargs)
        kwargs = {}
    else:
        psf, data = psf_fp, data_fp
    recon = ADMM(psf, n_iter=n_iter, **kwargs)
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
