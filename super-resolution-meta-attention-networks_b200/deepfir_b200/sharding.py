"""Multi-GPU inference sharding: images are independent units (no BatchNorm, every pool / attention is per
image), so ranks take contiguous slices of the image list and never exchange activations.  One LARGE image
shards by LR tile: the reference's four overlapping quadrants (halo = `shave` pixels, handlers.py:99-137) are the units, spread
over the ranks; the only exchange is the assembly of the stitched result (one all-reduce of disjoint parts)."""
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


def quadrant_geometry(height: int, width: int, scale: int, shave: int = 10):
    """The reference's chop (attention_manipulators/handlers.py:99-137): quadrant (far_r, far_c) covers LR rows
    [0, h/2 + shave) or [h - h/2 - shave, h) (same for columns); of its SR result the part nearest to its own image corner
    is kept.  Returns [(lr_rows, lr_cols, dst_rows, dst_cols, src_rows, src_cols)] for the four quadrants, SR slices in
    output pixels."""
    half = (height // 2, width // 2)
    size = (half[0] + shave, half[1] + shave)
    full = (height, width)
    out = []
    for far_r in (0, 1):
        for far_c in (0, 1):
            lr, dst, src = [], [], []
            for axis, far in ((0, far_r), (1, far_c)):
                lr.append(slice(full[axis] - size[axis], full[axis]) if far else slice(0, size[axis]))
                cut, whole, tile = scale * half[axis], scale * full[axis], scale * size[axis]
                dst.append(slice(cut, whole) if far else slice(0, cut))
                src.append(slice(tile - (whole - cut), tile) if far else slice(0, cut))
            out.append((lr[0], lr[1], dst[0], dst[1], src[0], src[1]))
    return out


def run_chopped_sharded(forward_fn, x, meta, scale: int, rank: int, world: int, shave: int = 10, out_channels=None):
    """One (batch of) large LR image(s) evaluated as the reference's four overlapping quadrants, the 4 * B (quadrant, image)
    units spread over `world` ranks (contiguous, balanced).  Each rank runs forward_fn ONCE on its units (they all have the
    same size), writes the kept parts into a zero-initialised full-size result and the ranks add their disjoint parts with one
    all-reduce (NCCL on GPUs, gloo in the CPU tests; nothing to do for world == 1).  Every rank returns the stitched result,
    identical to the single-process `forward_chop`."""
    import torch.distributed as dist
    B, _, H, W = x.shape
    geo = quadrant_geometry(H, W, scale, shave)
    units = [(q, b) for q in range(4) for b in range(B)]
    a, e = shard_range(len(units), rank, world)
    mine = units[a:e]
    out = None
    if mine:
        xs = torch.stack([x[b, :, geo[q][0], geo[q][1]] for q, b in mine])
        ms = torch.stack([meta[b] for _, b in mine]) if meta is not None else None
        sr = forward_fn(xs.contiguous(), ms)
        out = sr.new_zeros(B, sr.shape[1], scale * H, scale * W)
        for i, (q, b) in enumerate(mine):
            _, _, dr, dc, sr_r, sr_c = geo[q]
            out[b, :, dr, dc] = sr[i, :, sr_r, sr_c]
    if world > 1 and dist.is_available() and dist.is_initialized():
        if out is None:
            if out_channels is None:
                raise RuntimeError("run_chopped_sharded: a rank without units needs out_channels to allocate its zero part")
            out = x.new_zeros(B, out_channels, scale * H, scale * W, dtype=torch.float32)
        dist.all_reduce(out)
    return out
