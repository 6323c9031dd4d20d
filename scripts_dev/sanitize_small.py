"""Small ADMM / FISTA / unrolled runs for compute-sanitizer (memcheck, racecheck, synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import lensless_oracle as orc
import lenslesspicam_b200 as L
for shape in [(24, 32, 3), (23, 38, 3), (135, 240, 1)]:
    h, w, C = shape
    psf, data = orc.synthetic_problem(h, w, C, seed=3)
    for name, cls, n in (("admm", L.ADMM, 4), ("fista", L.FISTA, 4)):
        rec = cls(torch.from_numpy(psf).cuda(), no_graph=True)
        rec.set_data(torch.from_numpy(data).cuda())
        out = rec.apply(n_iter=n, disp_iter=None, plot=False).cpu().numpy()
        ref = orc.reconstruct(name, psf, data, n)
        print(shape, name, "rel err %.2e" % (np.abs(out - ref).max() / np.abs(ref).max()))
torch.cuda.synchronize()
print("done")
