"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/lpc.h
declares, and its integer geometry is bit-exact with the reference (no compute calls)."""
import ctypes
import os
import re

import pytest

import lensless_oracle as orc
from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "lpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpc_[a-z_A-Z0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lenslesspicam_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"liblpc.so does not export {n}"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in _lib.py"
    assert set(_lib.PROTOTYPES) == set(names)


def test_version_and_error_names():
    from lenslesspicam_b200 import _lib

    lib = _lib.load()
    assert lib.lpc_version() == 100
    assert lib.lpc_error_name(0) == b"LPC_OK"
    assert lib.lpc_error_name(-1) == b"LPC_EINVAL"


def test_next_fast_len_bit_exact():
    from lenslesspicam_b200 import _lib
    from scipy.fftpack import next_fast_len

    lib = _lib.load()
    for n in list(range(1, 6000)) + [6079, 8111, 12345, 65537]:
        assert lib.lpc_next_fast_len(n) == next_fast_len(n) == orc.next_fast_len(n)


@pytest.mark.parametrize(
    "h,w,expect",
    [
        (1080, 1920, (2160, 3840, 540, 960)),
        (380, 507, (768, 1024, 194, 258)),  # bundled data, downsample 8
        (760, 1014, (1536, 2048, 388, 517)),
        (203, 61, (405, 125, 101, 32)),  # odd padded sizes
        (47, 29, (96, 60, 24, 15)),  # test_convolver.py shapes
    ],
)
def test_geometry_bit_exact(h, w, expect):
    from lenslesspicam_b200 import _lib

    lib = _lib.load()
    v = [ctypes.c_int() for _ in range(4)]
    assert lib.lpc_geometry(h, w, *v) == 0
    got = tuple(x.value for x in v)
    assert got == expect == orc.padded_geometry(h, w)


def test_bad_arguments_are_rejected_without_a_gpu():
    from lenslesspicam_b200 import _lib

    lib = _lib.load()
    p = ctypes.c_void_p()
    # C must be 1 or 3 (recon.py:256), algorithm / dtype enums checked before any CUDA call
    assert lib.lpc_plan_create(ctypes.byref(p), 0, _lib.ALGO_ADMM, _lib.F32, 0, 1, 1, 2, 8, 8, 0) == -1
    assert b"rgb" in lib.lpc_last_error(None)
    assert lib.lpc_plan_create(ctypes.byref(p), 0, 9, _lib.F32, 0, 1, 1, 3, 8, 8, 0) == -1
    assert lib.lpc_plan_create(ctypes.byref(p), 0, _lib.ALGO_ADMM, 7, 0, 1, 1, 3, 8, 8, 0) == -1
    assert lib.lpc_geometry(0, 5, None, None, None, None) == -1
    assert lib.lpc_run(None, 1, None) == -1


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "lenslesspicam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "lensless_oracle" not in txt and "oracle/" not in txt, f


def test_cpu_tensor_is_refused_loudly():
    import torch

    import lenslesspicam_b200 as L

    with pytest.raises(TypeError):
        L.ADMM(torch.rand(1, 8, 8, 3))
    with pytest.raises(TypeError):
        L.FISTA(torch.rand(1, 8, 8, 3).numpy())
    with pytest.raises(TypeError):
        L.RealFFTConvolve2D(torch.rand(1, 8, 8, 3))
