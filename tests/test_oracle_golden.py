"""
Pin the CPU oracle (oracle/lensless_oracle.py) against outputs of the reference
itself (tests/golden/*.npz, written by tests/golden/make_golden.py).  No GPU.

Tolerances: the oracle follows the reference's numpy branch operation for
operation, so against the numpy-fp32 goldens (``n32``) it must agree to fp32
round-off (<= 2e-6 rel-to-max); against the torch-fp32 goldens (``t32``) the
only difference is the FFT library (pocketfft vs MKL/pocketfft-in-torch), for
which SURVEY 7.4 measured ~5e-6 -> bound 1e-5 (north_star's tolerance).
"""
import numpy as np
import pytest

import lensless_oracle as orc
from conftest import load_golden, rel_to_max

CASES = ["small_even_rgb", "small_odd_rgb", "small_gray", "small_mixed"]
ALGOS = ["admm", "gd", "nesterov", "fista"]


def test_next_fast_len_matches_scipy():
    from scipy.fftpack import next_fast_len

    for n in list(range(1, 5000)) + [2159, 3839, 6079, 8111]:
        assert orc.next_fast_len(n) == next_fast_len(n)


def test_geometry_headline():
    # 1080x1920 -> 2160x3840, window rows 540:1620, cols 960:2880
    assert orc.padded_geometry(1080, 1920) == (2160, 3840, 540, 960)
    assert orc.padded_geometry(380, 507) == (768, 1024, 194, 258)
    assert orc.padded_geometry(203, 61)[:2] == (405, 125)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("algo", ALGOS)
def test_oracle_matches_reference(case, algo):
    g = load_golden(case)
    for n in (1, 5, 30):
        out = orc.reconstruct(algo, g["psf"], g["data"], n)
        assert out.dtype == np.float32 and out.shape == g[f"{algo}_t32_{n}"].shape
        if f"{algo}_n32_{n}" in g:
            assert rel_to_max(out, g[f"{algo}_n32_{n}"]) <= 2e-6, (case, algo, n)
        assert rel_to_max(out, g[f"{algo}_t32_{n}"]) <= 1e-5, (case, algo, n)


def test_oracle_fp64_matches_reference_fp64():
    g = load_golden("small_odd_rgb")
    for algo in ALGOS:
        out = orc.reconstruct(algo, g["psf"].astype(np.float64), g["data"].astype(np.float64), 30, dtype="float64")
        assert out.dtype == np.float64
        assert rel_to_max(out, g[f"{algo}_t64_30"]) <= 1e-10, algo


def test_oracle_bundled_data():
    g = load_golden("bundled_ds16")
    out = orc.reconstruct("admm", g["psf"], g["data"], 5)
    assert rel_to_max(out, g["admm_n32_5"]) <= 2e-6
    assert rel_to_max(out, g["admm_t32_5"]) <= 1e-5
    out = orc.reconstruct("fista", g["psf"], g["data"], 20)
    assert rel_to_max(out, g["fista_t32_20"]) <= 1e-5


def test_oracle_variants():
    g = load_golden("variants")
    kw = dict(mu1=3e-6, mu2=2e-5, mu3=1e-4, tau=3e-4)
    out = orc.reconstruct("admm", g["psf"], g["data"], 10, **kw)
    assert rel_to_max(out, g["admm_params_t32_10"]) <= 1e-5
    out = orc.reconstruct("admm", g["psf"], g["data"], 10, initial_est=g["init"].copy())
    assert rel_to_max(out, g["admm_init_t32_10"]) <= 1e-5
    s = orc.OracleADMM(g["psf"])
    s.set_data(g["data"])
    s.apply(n_iter=5)
    assert rel_to_max(s.apply(n_iter=5, reset=False), g["admm_cont_t32_5p5"]) <= 1e-5
    s = orc.OracleADMM(g["psf"])
    s.set_data(g["data"].copy())
    assert rel_to_max(s.apply(n_iter=5, background=g["bg"]), g["admm_bg_t32_5"]) <= 1e-5
    out = orc.reconstruct("fista", g["psf"], g["data"], 10, tk=2.0, lip_fact=1.2)
    assert rel_to_max(out, g["fista_params_t32_10"]) <= 1e-5


def test_oracle_depth_gd():
    g = load_golden("depth_gd")
    for algo in ("gd", "fista"):
        out = orc.reconstruct(algo, g["psf"], g["data"], 10)
        assert out.shape == g[f"{algo}_t32_10"].shape == (4, 16, 20, 3)
        assert rel_to_max(out, g[f"{algo}_t32_10"]) <= 1e-5


def test_oracle_convolver():
    g = load_golden("convolver")
    for norm in ("ortho", "backward"):
        c = orc.OracleConvolver(g["psf"], pad=True, norm=norm)
        assert list(c.padded_shape) == list(g["padded_shape"])
        assert list(c.start) == list(g["start"])
        assert rel_to_max(c.convolve(g["data"]), g[f"conv_{norm}"]) <= 1e-5
        assert rel_to_max(c.deconvolve(g["data"]), g[f"deconv_{norm}"]) <= 1e-5
        assert rel_to_max(c.convolve(g["data_d1"]), g[f"conv_d1_{norm}"]) <= 1e-5
    c = orc.OracleConvolver(g["psf"], pad=True, norm="ortho")
    assert rel_to_max(c.H, g["H_ortho"]) <= 1e-5
    # pad/crop round trip is exact (test_convolver.py:11-28)
    padded = c.pad_to(g["data"])
    assert list(padded.shape[1:]) == list(g["padded_shape"])
    assert np.array_equal(c.crop(padded), g["data"])
    c2 = orc.OracleConvolver(g["psf"], pad=False, norm="backward")
    assert rel_to_max(c2.convolve(padded), g["conv_nopad"]) <= 1e-5
    assert rel_to_max(c2.deconvolve(padded), g["deconv_nopad"]) <= 1e-5


def test_oracle_batch_equals_unrolled():
    g = load_golden("batch")
    outs = np.stack([orc.reconstruct("admm", g["psf"], g["batch"][b], 10) for b in range(4)])
    assert rel_to_max(outs, g["admm_t32_10"]) <= 1e-5
    assert rel_to_max(outs, g["unrolled_admm_t32_10"][:, :, :, :, :].reshape(outs.shape)) <= 1e-5


def test_oracle_unrolled_matches_reference():
    """UnrolledADMM / UnrolledFISTA .forward(batch) of the reference with per-iteration parameters
    (tests/golden/unrolled.npz) vs the oracle's per-sample restatement."""
    g = load_golden("unrolled")
    n = len(g["mu1"])
    for b in range(g["batch"].shape[0]):
        out = orc.reconstruct("unrolled_admm", g["psf"], g["batch"][b], n, mu1=g["mu1"], mu2=g["mu2"], mu3=g["mu3"], tau=g["tau"])
        assert rel_to_max(out, g["unrolled_admm_t32"][b]) <= 1e-5, b
        out = orc.reconstruct("unrolled_fista", g["psf"], g["batch"][b], n, alpha=g["alpha"], tk=g["tk"])
        assert rel_to_max(out, g["unrolled_fista_t32"][b]) <= 1e-5, b
        out = orc.reconstruct("unrolled_fista", g["psf"], g["batch"][b], n)
        assert rel_to_max(out, g["unrolled_fista_default_t32"][b]) <= 1e-5, b
    out = orc.reconstruct("unrolled_admm", g["psf"].astype(np.float64), g["batch"][0].astype(np.float64), n, dtype="float64",
                          mu1=g["mu1"], mu2=g["mu2"], mu3=g["mu3"], tau=g["tau"])
    assert rel_to_max(out, g["unrolled_admm_t64"][0]) <= 1e-10
    out = orc.reconstruct("unrolled_fista", g["psf"].astype(np.float64), g["batch"][0].astype(np.float64), n, dtype="float64",
                          alpha=g["alpha"], tk=g["tk"])
    assert rel_to_max(out, g["unrolled_fista_t64"][0]) <= 1e-10


def test_torch_restatement_matches_reference():
    """oracle/lensless_oracle_torch.py (the reference's torch branch restated; timed on the GPU by bench.py as the
    'reference torch path on the same device') against the reference's own torch-fp32 outputs."""
    import torch

    import lensless_oracle_torch as ot

    for case in ("small_even_rgb", "small_odd_rgb", "small_gray"):
        g = load_golden(case)
        for algo, n in (("admm", 30), ("fista", 30)):
            s = ot.ALGOS[algo](torch.from_numpy(g["psf"]))
            s.set_data(torch.from_numpy(g["data"]))
            out = s.apply(n).numpy()
            assert rel_to_max(out, g[f"{algo}_t32_{n}"]) <= 2e-6, (case, algo)
