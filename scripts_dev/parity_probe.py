"""Where does the full-size ADMM-100 error come from?  (GPU; run through gpurun.)
Compares, against the fp64 torch restatement of the reference on the same GPU: the fp32 restatement, the fp32 restatement
with the measurement perturbed by <= 1 ulp (the problem's own fp32 conditioning), our fused engine (v2), our cuFFT engine
(v1: same fused algebra, library FFTs), at several iteration counts."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lensless_oracle as orc  # noqa: E402
import lensless_oracle_torch as ot  # noqa: E402

import lenslesspicam_b200 as L  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def ref(algo, psf, data, n, dt):
    s = ot.ALGOS[algo](psf.to(dt))
    s.set_data(data.to(dt))
    out = s.apply(n).clone()
    del s
    torch.cuda.empty_cache()
    return out


def main():
    h, w = (int(v) for v in os.environ.get("PROBE_SHAPE", "1080x1920").split("x"))
    iters = [int(v) for v in os.environ.get("PROBE_ITERS", "10,30,100").split(",")]
    for seed in (0, 1):
        psf_np, data_np = orc.synthetic_problem(h, w, 3, seed=seed)
        psf, data = torch.from_numpy(psf_np).cuda(), torch.from_numpy(data_np).cuda()
        gen = torch.Generator(device="cuda").manual_seed(5)
        # <= 1 ulp relative perturbation of the measurement
        pert = data * (1 + (torch.rand(data.shape, generator=gen, device="cuda") - 0.5) * 2 * 6e-8)
        for algo in ("admm", "fista"):
            for n in iters if algo == "admm" else [3 * i for i in iters]:
                truth = ref(algo, psf, data, n, torch.float64)
                r32 = ref(algo, psf, data, n, torch.float32)
                r32p = ref(algo, psf, pert, n, torch.float32)
                row = {"ref32": rel(r32, truth), "ref32_1ulp_input": rel(r32p, truth), "ref32p_vs_ref32": rel(r32p, r32)}
                for eng in (None, "v1"):
                    cls = L.ADMM if algo == "admm" else L.FISTA
                    rec = cls(psf, engine=eng) if eng else cls(psf)
                    rec.set_data(data)
                    o = rec.apply(n_iter=n, disp_iter=None, plot=False)
                    row["ours_" + (eng or "v2")] = rel(o, truth)
                    if eng is None:
                        row["ours_v2_vs_ref32"] = rel(o, r32)
                        rec.set_data(pert)
                        row["ours_v2_1ulp_input"] = rel(rec.apply(n_iter=n, disp_iter=None, plot=False), truth)
                    del rec
                    torch.cuda.empty_cache()
                print(f"seed {seed} {algo}-{n}: " + "  ".join(f"{k}={v:.2e}" for k, v in row.items()), flush=True)


if __name__ == "__main__":
    main()
