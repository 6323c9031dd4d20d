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


def test_shard_batch_properties():
    from lenslesspicam_b200.dist import shard_batch

    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 4, 8):
            parts = [list(shard_batch(n, world, r)) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
