"""
Parity of the CUDA engine on the BASELINE.json configurations AT THEIR STATED ITERATION COUNTS
("next round" item 1).  Needs a GPU: ``pytest -m gpu``.

    cfg1  bundled tape PSF / thumbs-up, load_data(downsample=8) RGB 3x380x507 and downsample=4 gray 1x760x1014
          (profile/admm.py:12-16): ADMM-5, ADMM-100, FISTA-300 against the reference's committed torch-fp32 outputs
    cfg2  ADMM-100  @ 3x1080x1920 (padded 2160x3840)
    cfg3  FISTA-300 @ 3x1080x1920
    cfg4  ADMM-100, a batch of 4 full-size measurements (one GPU's share of the 256-image job)
    cfg5  "3-D" ADMM-50, 16-depth PSF stack x 1080x1920 (16 independent 2-D problems, SURVEY 8a note)

The truth at full size is the torch restatement of the reference (oracle/lensless_oracle_torch.py, pinned to the
reference's goldens in tests/test_oracle_golden.py) run ON THE SAME GPU in fp64; the reference trajectory is the
same restatement in fp32.  Gate (north_star 1e-5 where reachable, SURVEY 7.4 envelope otherwise):

    e(ours, fp64) <= 1e-5     or     e(ours, fp64) <= 1.5 * e(ref32, fp64)

with e = max|a-b| / max|b|.  Every measured triple (ours-vs-fp64, ref32-vs-fp64, ours-vs-ref32) is appended to
``gpurun_out/parity_r02.json`` so the numbers quoted in DESIGN.md section 5 are the test's own.
"""
import json
import os

import numpy as np
import pytest
import torch

import lensless_oracle as orc
import lensless_oracle_torch as ot
from conftest import ROOT, load_golden, rel_to_max

pytestmark = pytest.mark.gpu

TOL = 1e-5
ENVELOPE = 1.5


def L():
    import lenslesspicam_b200

    return lenslesspicam_b200


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def report(name, **vals):
    path = os.environ.get("LPC_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in vals.items()}
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except Exception:
        pass
    print(f"parity {name}: " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))


def torch_ref(algo, psf, data, n_iter, dtype):
    """the reference's torch branch restated, on the GPU, fp32 (trajectory) or fp64 (truth)"""
    s = ot.ALGOS[algo](dev(psf, dtype))
    s.set_data(dev(data, dtype))
    out = s.apply(n_iter).cpu().numpy()
    del s
    torch.cuda.empty_cache()
    return out


def gate(name, out, ref32, ref64, extra=None):
    e_ours = rel_to_max(out, ref64)
    e_ref = rel_to_max(ref32, ref64)
    e_32 = rel_to_max(out, ref32)
    report(name, ours_vs_fp64=e_ours, ref32_vs_fp64=e_ref, ours_vs_ref32=e_32, **(extra or {}))
    assert e_ours <= TOL or e_ours <= ENVELOPE * e_ref, (name, e_ours, e_ref, e_32)
    # and the distance to the fp32 reference is explained by the two fp32 round-off clouds
    assert e_32 <= TOL or e_32 <= e_ours + e_ref + 1e-7, (name, e_ours, e_ref, e_32)
    return e_ours, e_ref, e_32


def ours(algo, psf, data, n_iter, **kw):
    cls = {"admm": L().ADMM, "fista": L().FISTA}[algo]
    rec = cls(dev(psf), **kw)
    rec.set_data(dev(data))
    out = rec.apply(n_iter=n_iter, disp_iter=None, plot=False)
    from lenslesspicam_b200 import _lib

    assert _lib.load().lpc_engine_version(rec._plan) == 2, "the fused engine must serve the BASELINE shapes"
    return out.cpu().numpy(), rec


# ------------------------------------------------------------------------------ cfg 1: bundled data
@pytest.mark.parametrize("case", ["bundled_ds8", "bundled_ds4_gray"])
def test_cfg1_bundled_data_at_reference_downsamples(case):
    g = load_golden(case)
    psf, data = g["psf"], g["data"]
    for algo, n in (("admm", 5), ("admm", 100), ("fista", 300)):
        out, rec = ours(algo, psf, data, n)
        ref32 = g[f"{algo}_t32_{n}"]  # the reference itself (tests/golden/make_golden.py)
        assert out.shape == ref32.shape and out.dtype == np.float32
        ref64 = torch_ref(algo, psf, data, n, torch.float64)
        gate(f"cfg1/{case}/{algo}-{n}", out, ref32, ref64)
        if algo == "admm" and n == 5:
            # reconstruction_error of the reference on its own ADMM-5 state (recon.py:607-653)
            for norm, key in ((True, "admm5_recon_error_norm"), (False, "admm5_recon_error_raw")):
                err = rec.reconstruction_error(normalize=norm).cpu().numpy()
                assert err.shape == g[key].shape
                np.testing.assert_allclose(err, g[key], rtol=2e-4)


# ------------------------------------------------------------------------------ cfg 2 / 3: the headline shape
def test_cfg2_admm100_full_size():
    psf, data = orc.synthetic_problem(1080, 1920, 3, seed=0)
    out, _ = ours("admm", psf, data, 100)
    ref32 = torch_ref("admm", psf, data, 100, torch.float32)
    ref64 = torch_ref("admm", psf, data, 100, torch.float64)
    assert out.shape == (1, 1080, 1920, 3) and (out >= 0).all()
    gate("cfg2/admm-100@3x1080x1920", out, ref32, ref64)


def test_cfg3_fista300_full_size():
    psf, data = orc.synthetic_problem(1080, 1920, 3, seed=1)
    out, _ = ours("fista", psf, data, 300)
    ref32 = torch_ref("fista", psf, data, 300, torch.float32)
    ref64 = torch_ref("fista", psf, data, 300, torch.float64)
    gate("cfg3/fista-300@3x1080x1920", out, ref32, ref64)
    # 50 iterations: about 1e-5 from the truth for the reference's fp32 arithmetic and for ours (SURVEY 7.4 measured
    # 3.6e-6 at 380x507; the full-size transforms are 8x longer)
    out50, _ = ours("fista", psf, data, 50)
    gate("cfg3/fista-50@3x1080x1920", out50, torch_ref("fista", psf, data, 50, torch.float32), torch_ref("fista", psf, data, 50, torch.float64))


def test_gpu_fp64_truth_equals_cpu_oracle_at_full_size():
    """Independent cross-check of the truth used above: the torch restatement in fp64 on the GPU against the numpy /
    pocketfft oracle in fp64 on the host, full size, a few iterations (both deterministic fp64: agreement here
    carries to any iteration count).  LPC_SLOW=1 runs all 100 ADMM iterations on the host (minutes)."""
    psf, data = orc.synthetic_problem(1080, 1920, 3, seed=0)
    n = 100 if os.environ.get("LPC_SLOW") else 4
    cpu = orc.reconstruct("admm", psf.astype(np.float64), data.astype(np.float64), n, dtype="float64", workers=os.cpu_count())
    gpu = torch_ref("admm", psf, data, n, torch.float64)
    e = rel_to_max(gpu, cpu)
    report(f"truth/admm-{n}@3x1080x1920 torch-fp64(GPU) vs numpy-fp64(CPU)", err=e)
    assert e <= 1e-10
    if n == 100:
        out, _ = ours("admm", psf, data, n)
        e_cpu = rel_to_max(out, cpu)
        report("cfg2/admm-100 ours vs numpy-fp64(CPU)", err=e_cpu)
        assert e_cpu <= 3 * TOL
    cpu = orc.reconstruct("fista", psf.astype(np.float64), data.astype(np.float64), 4, dtype="float64", workers=os.cpu_count())
    gpu = torch_ref("fista", psf, data, 4, torch.float64)
    assert rel_to_max(gpu, cpu) <= 1e-10


# ------------------------------------------------------------------------------ cfg 4: batch at full size
def test_cfg4_batch4_admm100_full_size():
    psf, _ = orc.synthetic_problem(1080, 1920, 3, seed=0)
    datas = [orc.synthetic_problem(1080, 1920, 3, seed=40 + b)[1] for b in range(4)]
    rec = L().ADMM(dev(psf), n_iter=100)
    out = rec.forward(torch.stack([dev(d) for d in datas])).cpu().numpy()  # [4,1,h,w,3]
    assert out.shape == (4, 1, 1080, 1920, 3)
    for b in (0, 3):
        ref32 = torch_ref("admm", psf, datas[b], 100, torch.float32)
        ref64 = torch_ref("admm", psf, datas[b], 100, torch.float64)
        gate(f"cfg4/batch4[{b}] admm-100@3x1080x1920", out[b], ref32, ref64)
    # a batched run equals the single-measurement run of the same engine (planes are independent problems)
    one, _ = ours("admm", psf, datas[1], 100)
    assert rel_to_max(out[1], one) <= 1e-6


# ------------------------------------------------------------------------------ cfg 5: depth stack at full size
def test_cfg5_depth16_admm50_full_size():
    """16 independent 2-D ADMM problems against one measurement (the reference has no 3-D ADMM; its 3-D gradient
    descent has exactly these semantics -- SURVEY 8a): the oracle is 16 runs of the 2-D restatement."""
    rng = np.random.default_rng(5)
    psf = (rng.random((16, 1080, 1920, 1)) ** 8).astype(np.float32)
    psf /= np.linalg.norm(psf.ravel())
    _, data = orc.synthetic_problem(1080, 1920, 1, seed=6)
    rec = L().ADMM(dev(psf), allow_depth=True)
    rec.set_data(dev(data))
    out = rec.apply(n_iter=50, disp_iter=None, plot=False).cpu().numpy()
    assert out.shape == (16, 1080, 1920, 1)
    worst = (0.0, 0.0, 0.0)
    for d in (0, 7, 15):
        ref32 = torch_ref("admm", psf[d : d + 1], data, 50, torch.float32)
        ref64 = torch_ref("admm", psf[d : d + 1], data, 50, torch.float64)
        worst = max(worst, gate(f"cfg5/depth16[{d}] admm-50@1080x1920", out[d : d + 1], ref32, ref64))
