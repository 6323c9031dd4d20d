"""Device time of the pieces of one reconstruction step (set_data, reset, run(100), get_image) at the headline size."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lenslesspicam_b200 as L
from lenslesspicam_b200 import _lib
h, w, C = 1080, 1920, 3
g = torch.Generator().manual_seed(0)
psf = (torch.rand((1, h, w, C), generator=g) ** 8).cuda(); psf /= psf.norm()
data = torch.rand((1, 1, h, w, C), generator=g).cuda()
stream = torch.cuda.Stream()
lib = _lib.load()
with torch.cuda.stream(stream):
    rec = L.ADMM(psf, n_iter=100)
    rec.set_data(data)
    for _ in range(3): rec.apply(disp_iter=None, plot=False)
    st = stream.cuda_stream
    out = torch.empty((1, 1, h, w, C), device="cuda")
    def ev(): return torch.cuda.Event(enable_timing=True)
    names = ["set_data", "reset", "run(100)", "get_image"]
    acc = [0.0] * 4
    N = 10
    for _ in range(N):
        e = [ev() for _ in range(5)]
        e[0].record(stream)
        _lib.check(lib.lpc_set_data(rec._plan, data.data_ptr(), 1, None, None, None, None, st), rec._plan); e[1].record(stream)
        _lib.check(lib.lpc_reset(rec._plan, None, st), rec._plan); e[2].record(stream)
        _lib.check(lib.lpc_run(rec._plan, 100, st), rec._plan); e[3].record(stream)
        _lib.check(lib.lpc_get_image(rec._plan, out.data_ptr(), st), rec._plan); e[4].record(stream)
        stream.synchronize()
        for i in range(4): acc[i] += e[i].elapsed_time(e[i + 1])
    for n, a in zip(names, acc): print(f"{n:10s} {a / N:8.3f} ms")
    print("total      %8.3f ms" % (sum(acc) / N))
