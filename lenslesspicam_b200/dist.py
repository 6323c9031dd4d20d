"""
Multi-GPU plumbing for the batch-sharded path: one process per GPU, the batch axis
split contiguously across ranks, the PSF broadcast ONCE at init (NCCL over NVLink/NVSwitch; gloo on
CPU in the tests), and no collective inside the iteration.  Timing is reduced with a MAX over ranks.

The reference has no multi-GPU construct for this path (its only one is torch.nn.DataParallel for
training learned models, utils/dataset.py:2285-2290).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(world_size, rank, local_rank) from the torchrun environment (1, 0, 0 when launched plainly)."""
    return (
        int(os.environ.get("WORLD_SIZE", "1")),
        int(os.environ.get("RANK", "0")),
        int(os.environ.get("LOCAL_RANK", "0")),
    )


def init(backend=None, device=None):
    world, rank, _ = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world, rank


def shard_batch(n_items, world, rank):
    """Contiguous split of range(n_items): the first n_items % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def broadcast_psf(psf, src=0):
    """Make every rank hold rank `src`'s PSF (in place).  The only data-path collective of a job."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(psf, src=src)
    return psf


def max_over_ranks(value, device=None):
    """MAX-reduce a python float (a device time in ms) over all ranks."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device=None):
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()


def reconstruct_sharded(host_batch, psf, n_iter=100, algo="admm", chunk=1, gather=False, solver=None,
                        reconstruct_fn=None, device=None, out=None, total=None, **solver_kw):
    """BASELINE configs[3] as a call: reconstruct ``host_batch`` ([N, Dd, h, w, C], host memory, the same tensor on
    every rank) with the batch axis sharded contiguously over the ranks of the job.

    Every rank takes ``shard_batch(N, world, rank)``, receives rank 0's PSF (the job's ONLY data-path collective,
    NCCL over NVLink/NVSwitch), builds the solver on its own GPU and streams its shard through
    ``apply_host_stream`` (H2D / iterations / D2H overlapped, ``chunk`` measurements per launch chain).  No collective
    runs inside the iteration.  Returns ``(local_out, index_range)`` -- ``local_out`` is this rank's pinned
    ``[n_local, D, h, w, C]`` -- or, with ``gather=True``, ``(full, index_range)`` where ``full`` is the
    ``[N, D, h, w, C]`` result on every rank (one all_gather of the padded shards at the end).

    ``total=N``: ``host_batch`` holds ONLY this rank's shard (``len(shard_batch(N, world, rank))`` measurements) of a global
    batch of N -- a loader that reads per rank never materialises the whole batch on every rank.
    ``out`` is this rank's pre-allocated (pinned) ``[n_local, D, h, w, C]`` result buffer -- pinning gigabytes of host
    memory takes seconds, a caller that reconstructs batch after batch allocates it once.
    ``solver`` re-uses an existing solver object (its plans and CUDA graphs); ``reconstruct_fn(inputs, outputs)`` replaces
    the engine call (the CPU tests of the host logic run it under gloo with a stand-in)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    N = int(host_batch.shape[0]) if total is None else int(total)
    mine = shard_batch(N, world, rank)
    base = mine.start if total is None else 0  # index of this rank's first measurement inside host_batch
    if total is not None:
        assert int(host_batch.shape[0]) == len(mine), "with total=N, host_batch must hold exactly this rank's shard"
    D = int(psf.shape[0])
    h, w, C = (int(v) for v in host_batch.shape[2:])
    psf = broadcast_psf(psf, src=0)
    n_local = len(mine)
    use_cuda = psf.is_cuda

    def pinned(t):
        return t.pin_memory() if (use_cuda and not t.is_pinned()) else t

    if out is not None:
        assert tuple(out.shape) == (n_local, D, h, w, C) and out.dtype == host_batch.dtype and not out.is_cuda
        local_out = out
    else:
        local_out = pinned(torch.empty((n_local, D, h, w, C), dtype=host_batch.dtype))
    if n_local:
        ins, outs = [], []
        for s in range(0, n_local, max(int(chunk), 1)):
            e = min(s + max(int(chunk), 1), n_local)
            ins.append(pinned(host_batch[base + s : base + e].contiguous()))
            outs.append(local_out[s:e])
        if reconstruct_fn is not None:
            reconstruct_fn(ins, outs)
        else:
            if solver is None:
                import lenslesspicam_b200 as L

                cls = {"admm": L.ADMM, "fista": L.FISTA, "gd": L.GradientDescent, "nesterov": L.NesterovGradientDescent}[algo]
                solver = cls(psf, n_iter=n_iter, **solver_kw)
            # the last chunk may be smaller: it goes through its own call (its own plan of that batch size)
            full = [k for k in range(len(ins)) if ins[k].shape[0] == ins[0].shape[0]]
            solver.apply_host_stream([ins[k] for k in full], [outs[k] for k in full], n_iter=n_iter)
            rest = [k for k in range(len(ins)) if k not in full]
            if rest:
                solver.apply_host_stream([ins[k] for k in rest], [outs[k] for k in rest], n_iter=n_iter)
    if not gather:
        return local_out, mine
    if world == 1:
        return local_out, mine
    # all_gather of equally sized (padded) shards, trimmed afterwards; device tensors under NCCL
    n_max = -(-N // world)
    dev = device if device is not None else (psf.device if use_cuda else torch.device("cpu"))
    pad = torch.zeros((n_max, D, h, w, C), dtype=host_batch.dtype, device=dev)
    pad[:n_local].copy_(local_out, non_blocking=True)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([parts[r][: len(shard_batch(N, world, r))] for r in range(world)], dim=0)
    return full, mine
