"""Small runs through every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck python scripts_dev/sanitize_small.py`
(fused ADMM incl. the TMA row gather and the persistent TMA column kernels, fused GD family, cuFFT engine, per-sample PSFs,
reconstruction_error, plug-and-play, module helpers, the loader)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lenslesspicam_b200 as L  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def problem(h, w, C, B=1, seed=0):
    gen = torch.Generator().manual_seed(seed)
    psf = torch.rand((1, h, w, C), generator=gen) ** 8
    psf = (psf / psf.norm()).cuda()
    data = torch.rand((B, 1, h, w, C), generator=gen).cuda()
    return psf, data


def blur(x, nl):
    return torch.clamp((1 - nl) * x + nl * (x + torch.roll(x, 1, -3) + torch.roll(x, -1, -3)) / 3, min=0)


def main():
    for (h, w, C) in ((24, 32, 3), (23, 38, 3), (135, 240, 1), (64, 100, 1), (100, 120, 1)):  # 100 -> padded 200 = 8.5.5: warp-local column passes
        psf, data = problem(h, w, C)
        for cls in (L.ADMM, L.FISTA, L.GradientDescent, L.NesterovGradientDescent):
            rec = cls(psf)
            rec.set_data(data)
            rec.apply(n_iter=3, disp_iter=None, plot=False)
            rec.apply(n_iter=2, disp_iter=None, plot=False, reset=False)
            print(cls.__name__, (h, w, C), "engine", rec.engine_version, float(rec.reconstruction_error()[0]))
    psf, data = problem(24, 32, 3, B=3)
    psfs = torch.stack([psf * (1 + 0.1 * k) for k in range(3)])
    for cls in (L.ADMM, L.FISTA):
        rec = cls(psf, n_iter=3)
        out = rec.forward(data, psfs=psfs)
        print("per-sample", cls.__name__, float(out.sum()), rec.reconstruction_error(prediction=out, lensless=data, psfs=psfs).tolist())
    psf, data = problem(24, 32, 3)
    for use_dual in (True, False):
        rec = L.ADMM(psf, denoiser={"network": blur, "noise_level": 0.3, "use_dual": use_dual})
        rec.set_data(data)
        print("pnp admm", use_dual, float(rec.apply(n_iter=3, disp_iter=None, plot=False).sum()))
    rec = L.FISTA(psf, denoiser={"network": blur, "noise_level": 0.3})
    rec.set_data(data)
    print("pnp fista", float(rec.apply(n_iter=3, disp_iter=None, plot=False).sum()))
    x = torch.randn(2, 1, 9, 7, 3, device="cuda")
    print("helpers", float(L.soft_thresh(x, 0.2).sum()), float(L.finite_diff_adj(L.finite_diff(x)).sum()), float(L.finite_diff_gram([1, 18, 14, 3]).real.sum()))
    p, d = L.load_data(os.path.join(G, "raw_psf.png"), os.path.join(G, "raw_data.png"), downsample=4, normalize=True)
    p2, d2 = L.load_data(os.path.join(G, "raw8_psf.png"), os.path.join(G, "raw8_data.png"), downsample=1.6, gray=True, normalize=True)
    print("loader", tuple(p.shape), float(d.max()), tuple(p2.shape))
    rec = L.UnrolledADMM(psf, n_iter=3, mu1=[1e-6, 2e-6, 1e-6])
    print("unrolled", float(rec.forward(data).sum()))
    torch.cuda.synchronize()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
