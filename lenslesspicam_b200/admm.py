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


def apply_admm(psf, data, n_iter, verbose=False, **kwargs):
    """admm.py:400-419 with tensors instead of file paths (loading images is the reference's job)."""
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
