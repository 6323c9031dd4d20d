"""World-size-2 test of the multi-GPU host logic on CPU (gloo): PSF broadcast, batch sharding,
max-over-ranks timing.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    from lenslesspicam_b200 import dist as lpd

    w, r = lpd.init(backend="gloo")
    assert (w, r) == (world, rank)
    gen = torch.Generator().manual_seed(0)
    psf = torch.rand((1, 8, 8, 3), generator=gen)
    if rank != 0:
        psf.zero_()
    lpd.broadcast_psf(psf, src=0)
    mine = list(lpd.shard_batch(7, world, rank))
    slow = lpd.max_over_ranks(10.0 + rank)
    lpd.barrier()
    out.put((rank, psf.sum().item(), mine, slow))
    lpd.finalize()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gen = torch.Generator().manual_seed(0)
    expect = torch.rand((1, 8, 8, 3), generator=gen).sum().item()
    assert all(abs(r[1] - expect) < 1e-6 for r in res), "every rank must hold rank 0's PSF"
    assert res[0][2] == [0, 1, 2, 3] and res[1][2] == [4, 5, 6]  # contiguous, disjoint, complete
    assert all(r[3] == 11.0 for r in res)  # MAX over ranks


def _worker_sharded(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    from lenslesspicam_b200 import dist as lpd

    lpd.init(backend="gloo")
    gen = torch.Generator().manual_seed(3)
    batch = torch.rand((7, 1, 6, 5, 3), generator=gen)  # the same host batch on every rank
    psf = torch.rand((1, 6, 5, 3), generator=gen) if rank == 0 else torch.zeros((1, 6, 5, 3))
    calls = []

    def fake_engine(ins, outs):  # stand-in for apply_host_stream: out = 2*in + sum(psf)
        calls.append([tuple(i.shape) for i in ins])
        for i, o in zip(ins, outs):
            o.copy_(2 * i + psf.sum())

    full, mine = lpd.reconstruct_sharded(batch, psf, n_iter=3, chunk=2, gather=True, reconstruct_fn=fake_engine)
    local, mine2 = lpd.reconstruct_sharded(batch, psf, n_iter=3, chunk=3, gather=False, reconstruct_fn=fake_engine)
    # the rank holds only its own shard of the global batch
    shard = batch[mine[0] : mine[-1] + 1].clone()
    local2, mine3 = lpd.reconstruct_sharded(shard, psf, n_iter=3, chunk=3, reconstruct_fn=fake_engine, total=7)
    assert list(mine3) == list(mine) and torch.equal(local2, local)
    calls.pop()
    out.put((rank, list(mine), full, local, calls, psf.sum().item()))
    lpd.finalize()


def test_reconstruct_sharded_gloo():
    """the sharded-batch call of BASELINE configs[3] under gloo with a stand-in engine: shards are contiguous and
    complete, every rank used rank 0's PSF, chunks cover the shard (ragged last chunk), gather returns the full batch"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gen = torch.Generator().manual_seed(3)
    batch = torch.rand((7, 1, 6, 5, 3), generator=gen)
    psf = torch.rand((1, 6, 5, 3), generator=gen)
    expect = 2 * batch + psf.sum()
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 6]
    for rank, mine, full, local, calls, psum in res:
        assert abs(psum - psf.sum().item()) < 1e-6
        assert torch.allclose(full, expect)
        assert torch.allclose(local, expect[mine[0] : mine[-1] + 1])
    assert res[0][4][0] == [(2, 1, 6, 5, 3), (2, 1, 6, 5, 3)]      # rank 0: 4 items in chunks of 2
    assert res[1][4][0] == [(2, 1, 6, 5, 3), (1, 1, 6, 5, 3)]      # rank 1: 3 items, ragged last chunk
    assert res[1][4][1] == [(3, 1, 6, 5, 3)]


def test_dist_imports_without_the_cuda_library():
    """lenslesspicam_b200.dist is torch.distributed plumbing only: importing it must not load liblpc.so"""
    import subprocess
    import sys

    code = ("import sys; import lenslesspicam_b200.dist as d; import lenslesspicam_b200._lib as l; "
            "assert l._lib is None, 'liblpc.so was loaded by importing dist'; print(len(d.shard_batch(5, 2, 0)))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "3", r.stderr


def test_shard_batch_properties():
    from lenslesspicam_b200.dist import shard_batch

    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 4, 8):
            parts = [list(shard_batch(n, world, r)) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
