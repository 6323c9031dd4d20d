import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import lensless_oracle as orc
import lenslesspicam_b200 as L
def rel(a, b): return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())
for shape in [(135, 240, 3), (203, 61, 3), (100, 136, 1), (380, 507, 3)]:
    h, w, C = shape
    psf, data = orc.synthetic_problem(h, w, C, seed=11)
    for n in (1, 2, 20):
        ref = orc.reconstruct("admm", psf, data, n)
        ref64 = orc.reconstruct("admm", psf, data, n, dtype="float64") if "dtype" in orc.reconstruct.__code__.co_varnames else None
        outs = {}
        for eng in (None, "v1"):
            r = L.ADMM(torch.from_numpy(psf).cuda(), engine=eng) if eng else L.ADMM(torch.from_numpy(psf).cuda())
            r.set_data(torch.from_numpy(data).cuda())
            outs[eng] = r.apply(n_iter=n, disp_iter=None, plot=False).cpu().numpy()
        print(shape, n, "v2 vs oracle32 %.2e" % rel(outs[None], ref), "v1 vs oracle32 %.2e" % rel(outs["v1"], ref),
              "v2 vs v1 %.2e" % rel(outs[None], outs["v1"]), ("v2 vs oracle64 %.2e  oracle32 vs oracle64 %.2e" % (rel(outs[None], ref64), rel(ref, ref64))) if ref64 is not None else "")
