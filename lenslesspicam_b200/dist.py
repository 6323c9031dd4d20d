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
