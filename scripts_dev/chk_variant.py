"""A/B correctness of an experimental build (LPC_LIB_PATH): ADMM-10 / FISTA-10 at the headline size, saved by the
first run to /tmp/chk_ref.pt and compared (max|a-b|/max|b|) by every later run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lenslesspicam_b200 as L
h, w, C = 1080, 1920, 3
g = torch.Generator().manual_seed(0)
psf = (torch.rand((1, h, w, C), generator=g) ** 8).cuda(); psf /= psf.norm()
data = torch.rand((1, 1, h, w, C), generator=g).cuda()
out = {}
for name, cls in (("admm", L.ADMM), ("fista", L.FISTA)):
    rec = cls(psf, n_iter=10)
    rec.set_data(data)
    out[name] = rec.apply(disp_iter=None, plot=False).clone().cpu()
ref = "/tmp/chk_ref.pt"
if not os.path.exists(ref):
    torch.save(out, ref); print("saved reference")
else:
    r = torch.load(ref)
    print({k: float((out[k] - r[k]).abs().max() / r[k].abs().max()) for k in out})
