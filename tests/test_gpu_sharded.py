"""
BASELINE configs[3] on hardware: ``lenslesspicam_b200.dist.reconstruct_sharded`` over 2 NCCL ranks against the single-GPU
result of the same batch, bit for bit (every plane is an independent problem and the kernels are deterministic, so the
sharded job must reproduce the single-GPU images exactly).  Needs >= 2 GPUs: ``gpurun --gpus 2 -- python -m pytest -m gpu
tests/test_gpu_sharded.py``; with one GPU the test is skipped.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(n=5, h=135, w=240, C=3):
    gen = torch.Generator().manual_seed(4)
    psf = torch.rand((1, h, w, C), generator=gen) ** 8
    psf = psf / psf.norm()
    batch = torch.rand((n, 1, h, w, C), generator=gen)
    return psf, batch


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    import lenslesspicam_b200 as L  # noqa: F401
    from lenslesspicam_b200 import dist as lpd

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    lpd.init(backend="nccl", device=dev)
    psf, batch = _problem()
    psf_dev = psf.to(dev) if rank == 0 else torch.zeros_like(psf).to(dev)  # only rank 0 holds the PSF before the broadcast
    full, mine = lpd.reconstruct_sharded(batch, psf_dev, n_iter=12, algo="admm", chunk=2, gather=True, device=dev)
    local, mine2 = lpd.reconstruct_sharded(batch, psf_dev, n_iter=12, algo="admm", chunk=1, gather=False)
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), full=full.cpu().numpy(), local=local.numpy(), mine=np.array(list(mine)))
    lpd.finalize()


def test_sharded_batch_equals_single_gpu_bit_for_bit(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import lenslesspicam_b200 as L

    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    psf, batch = _problem()
    rec = L.ADMM(psf.cuda(), n_iter=12)
    single = torch.cat([rec.forward(batch[i : i + 1].cuda()) for i in range(batch.shape[0])]).cpu().numpy()
    seen = []
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(d["full"], single), f"rank {r}: gathered result differs from the single-GPU run"
        assert np.array_equal(d["local"], single[d["mine"]])
        seen += list(d["mine"])
    assert seen == list(range(batch.shape[0]))


def test_engine_calls_leave_the_current_device_alone():
    """every C-ABI entry point runs on the plan's device and restores the caller's: a solver living on
    cuda:1 must not re-target torch's current device"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import lenslesspicam_b200 as L

    torch.cuda.set_device(0)
    psf, batch = _problem(n=1)
    rec = L.ADMM(psf.to("cuda:1"), n_iter=5)
    rec.set_data(batch[0].to("cuda:1"))
    out = rec.apply(disp_iter=None, plot=False)
    assert out.device == torch.device("cuda", 1)
    assert torch.cuda.current_device() == 0
    ref = L.ADMM(psf.to("cuda:0"), n_iter=5)
    ref.set_data(batch[0].to("cuda:0"))
    assert torch.equal(ref.apply(disp_iter=None, plot=False).cpu(), out.cpu())
    assert torch.zeros(1, device="cuda").device.index == 0
