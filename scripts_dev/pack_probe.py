"""ADMM-100 / ADMM-30 full-size error vs fp64 for several power-of-two packing scales (LPC_PACK_S / LPC_PACK_T are read
at plan creation)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lensless_oracle as orc  # noqa: E402
import lensless_oracle_torch as ot  # noqa: E402

import lenslesspicam_b200 as L  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    psf_np, data_np = orc.synthetic_problem(1080, 1920, 3, seed=0)
    psf, data = torch.from_numpy(psf_np).cuda(), torch.from_numpy(data_np).cuda()
    truths = {}
    for n in (30, 100):
        s = ot.TorchADMM(psf.double())
        s.set_data(data.double())
        truths[n] = s.apply(n).clone()
        del s
    s = ot.TorchADMM(psf.double()); s.set_data(data.double()); s.reset()
    for _ in range(30):
        s.update()
    print("rms V %.3e  rms HV %.3e  mean V %.3e mean HV %.3e  max V %.3e" % (s.V.pow(2).mean().sqrt(), s.HV.pow(2).mean().sqrt(), s.V.mean(), s.HV.mean(), s.V.abs().max()))
    rk = (s.mu3 * s.W - s.rho) + ot.finite_diff_adj(s.mu2 * s.U - s.eta)
    bq = s.mu1 * s.X - s.xi
    print("rms a %.3e  rms bq %.3e  mean a %.3e  mean bq %.3e" % (rk.pow(2).mean().sqrt(), bq.pow(2).mean().sqrt(), rk.mean(), bq.mean()))
    del s
    torch.cuda.empty_cache()
    combos = [(0, 0), (5, 0), (0, 4), (0, 8), (5, 8), (3, 6), (8, 8), (5, 10), (0, -4)]
    for es, et in combos:
        os.environ["LPC_PACK_S"], os.environ["LPC_PACK_T"] = str(es), str(et)
        row = []
        for n in (30, 100):
            rec = L.ADMM(psf)
            rec.set_data(data)
            o = rec.apply(n_iter=n, disp_iter=None, plot=False)
            row.append(rel(o, truths[n]))
            del rec
        print(f"pack_s=2^{es} pack_t=2^{et}: admm-30 {row[0]:.2e}  admm-100 {row[1]:.2e}", flush=True)


if __name__ == "__main__":
    main()
