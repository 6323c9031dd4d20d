"""Host harness for the hand-written FFT core (lenslesspicam_b200/csrc/lpc_fft.cuh).

Compiles tests/csrc/test_fft_host.cpp with nvcc as a HOST program (the butterflies, index math,
padding and twiddle tables are __host__ __device__) and runs it: in-register DFTs, Stockham
passes and the in-place DIF/DIT passes against a naive long-double DFT.  No GPU needed.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_fft_core_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "test_fft_host"
    src = os.path.join(ROOT, "tests", "csrc", "test_fft_host.cpp")
    inc = os.path.join(ROOT, "lenslesspicam_b200", "csrc")
    res = subprocess.run([nvcc, "-x", "cu", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), src], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-3000:]
    assert "ALL OK" in run.stdout
