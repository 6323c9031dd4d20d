"""
lenslesspicam_b200 -- B200-native (sm_100a) iterative deconvolution engine behind
the ``lensless.ADMM`` / ``lensless.FISTA`` / ``lensless.GradientDescent`` surface of
LCAV/LenslessPiCam.  Python holds torch CUDA tensors; the arithmetic runs in
``liblpc.so`` (hand-written CUDA kernels + cuFFT) through the C-ABI of
``include/lpc.h``.

There is no CPU fallback: the first access to any solver / operator name (``from lenslesspicam_b200 import ADMM``)
loads ``liblpc.so`` and raises if the extension has not been built.  Only ``lenslesspicam_b200.dist`` (torch.distributed
plumbing: batch sharding, PSF broadcast, max-over-ranks timing) imports without the library, so the host-side
multi-GPU logic can be exercised on machines that have neither a GPU nor the CUDA libraries.
"""
import importlib

__version__ = "0.2.0"

_EXPORTS = {
    "RealFFTConvolve2D": "rfft_convolve",
    "ReconstructionAlgorithm": "recon",
    "ADMM": "admm", "apply_admm": "admm",
    "soft_thresh": "admm", "finite_diff": "admm", "finite_diff_adj": "admm", "finite_diff_gram": "admm",
    "FISTA": "gd", "GradientDescent": "gd", "GradientDescentUpdate": "gd", "NesterovGradientDescent": "gd",
    "non_neg": "gd", "apply_gradient_descent": "gd",
    "UnrolledADMM": "unrolled", "UnrolledFISTA": "unrolled",
    "load_data": "io", "load_psf": "io", "load_image": "io", "rgb2gray": "io",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    from . import _lib

    _lib.load()  # raises ImportError when liblpc.so is missing: no CPU fallback
    value = getattr(importlib.import_module(f".{mod}", __name__), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS))
