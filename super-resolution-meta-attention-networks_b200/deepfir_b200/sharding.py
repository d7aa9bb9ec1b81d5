"""Multi-GPU inference sharding: images are independent units (no BatchNorm, every pool / attention is per
image), so ranks take contiguous slices of the image list and never exchange activations."""
import os

import torch


def shard_range(n_items: int, rank: int, world: int):
    """contiguous, balanced split: the first n_items % world ranks get one extra item"""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def max_over_ranks(value: float, device=None) -> float:
    """the multi-GPU timing rule: a step takes as long as its slowest rank"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """data-parallel training's one exchange step: average the flat gradient buffer over all ranks, in place (NCCL on
    GPUs, gloo in the CPU tests).  No-op outside torch.distributed or with a single rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat)
    flat.div_(dist.get_world_size())
    return flat


def run_sharded(forward_fn, x, meta, rank: int, world: int):
    """forward_fn(x_slice, meta_slice) on this rank's slice; returns (start, stop, output_slice)"""
    a, b = shard_range(x.shape[0], rank, world)
    if a == b:
        return a, b, None
    return a, b, forward_fn(x[a:b], meta[a:b])
