"""Multi-GPU plumbing: one process per GPU, environments sharded as independent contiguous ranges.

The hot path has no data-path collective (environments never interact); the only exchange is ONE all-reduce of the
episode-statistics vector per reporting interval (NCCL on GPUs; gloo in the CPU tests).  SURVEY.md §8(e).
"""
import os

import numpy as np


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(num_envs_total, rank, world):
    """Contiguous env range [lo, hi) owned by `rank` (the first `num_envs_total % world` ranks get one extra env)."""
    base, rem = divmod(int(num_envs_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(base_seed, envs_per_gpu, rank):
    """Weak scaling: every rank owns `envs_per_gpu` envs; env i of rank r is global env r*envs_per_gpu + i and is
    seeded with base_seed + that index, so a given global env is the same world whichever GPU runs it."""
    return int(base_seed) + int(rank) * int(envs_per_gpu) + np.arange(int(envs_per_gpu), dtype=np.int64)


def init_process_group(backend=None, device=None):
    """torch.distributed init from the torchrun environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = rank_world()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def allreduce_stats(stats, device=None):
    """Sums the int64 statistics vector (numpy array or tensor) over all ranks; returns a numpy array.  This is the
    single collective of the path."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(stats, dtype=np.int64)) if not torch.is_tensor(stats) else stats.to(torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        else:
            t = t.cpu().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value, device=None):
    """MAX of a python float over ranks (device-timed intervals are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
