"""
lenslesspicam_b200 -- B200-native (sm_100a) iterative deconvolution engine behind
the ``lensless.ADMM`` / ``lensless.FISTA`` / ``lensless.GradientDescent`` surface of
LCAV/LenslessPiCam.  Python holds torch CUDA tensors; the arithmetic runs in
``liblpc.so`` (hand-written CUDA kernels + cuFFT) through the C-ABI of
``include/lpc.h``.  There is no CPU fallback: importing the solver classes fails
loudly if the extension has not been built.
"""
from . import _lib

_lib.load()  # raise now, not at first use, if liblpc.so is missing

from .rfft_convolve import RealFFTConvolve2D  # noqa: E402
from .recon import ReconstructionAlgorithm  # noqa: E402
from .admm import ADMM, apply_admm  # noqa: E402
from .gd import (  # noqa: E402
    FISTA,
    GradientDescent,
    GradientDescentUpdate,
    NesterovGradientDescent,
    non_neg,
)

from .unrolled import UnrolledADMM, UnrolledFISTA  # noqa: E402

__version__ = "0.1.0"
__all__ = [
    "ADMM",
    "FISTA",
    "GradientDescent",
    "GradientDescentUpdate",
    "NesterovGradientDescent",
    "RealFFTConvolve2D",
    "UnrolledADMM",
    "UnrolledFISTA",
    "ReconstructionAlgorithm",
    "apply_admm",
    "non_neg",
]
