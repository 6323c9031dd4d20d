"""Host+device cost of the per-sample calls of an evaluation loop (eval/benchmark.py:240-260):
_set_psf, set_data + apply, reconstruction_error -- at the headline size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lenslesspicam_b200 as L
h, w, C = 1080, 1920, 3
g = torch.Generator().manual_seed(0)
psfs = [(torch.rand((1, h, w, C), generator=g) ** 8).cuda() for _ in range(3)]
psfs = [p / p.norm() for p in psfs]
data = torch.rand((1, 1, h, w, C), generator=g).cuda()
rec = L.ADMM(psfs[0], n_iter=100)
rec.set_data(data); rec.apply(disp_iter=None, plot=False); rec.reconstruction_error(); torch.cuda.synchronize()
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("_set_psf               %.2f ms" % t(lambda i: rec._set_psf(psfs[i % 3])))
print("set_data + apply(100)  %.2f ms" % t(lambda i: (rec.set_data(data), rec.apply(disp_iter=None, plot=False))))
print("reconstruction_error   %.2f ms" % t(lambda i: rec.reconstruction_error()))
print("full per-sample loop   %.2f ms" % t(lambda i: (rec._set_psf(psfs[i % 3]), rec.set_data(data), rec.apply(disp_iter=None, plot=False), rec.reconstruction_error())))
