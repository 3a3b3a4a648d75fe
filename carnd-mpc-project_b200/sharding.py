"""Multi-GPU plumbing: independent problems are sharded by batch index, one process per GPU; the only
exchange is the gather of the results after the solve.  torch.distributed is used for the
gather (NCCL on GPUs, gloo in the CPU tests); nothing here computes."""
import numpy as np


def shard_bounds(B, rank, world):
    """Contiguous batch-index range [lo, hi) of `rank`: sizes differ by at most one, ranks in order."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local, B, world, rank, dist=None, device=None):
    """All-gather per-rank arrays whose LAST axis is the local batch shard ([k][B_local] layout of the
    C-ABI) into the full [k][B] array on every rank.  Shards may be uneven (padded to the largest)."""
    import torch
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    if world == 1:
        return t
    sizes = [shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world)]
    mx = max(sizes)
    lead = tuple(t.shape[:-1])
    pad = torch.zeros(lead + (mx,), dtype=t.dtype, device=t.device)
    pad[..., : sizes[rank]] = t
    out = torch.empty((world,) + lead + (mx,), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), pad.contiguous().view(-1))
    return torch.cat([out[r][..., : sizes[r]] for r in range(world)], dim=-1)


def solve_sharded(solve_fn, inputs, B, world, rank, dist=None, device=None):
    """Solve this rank's shard with `solve_fn(**shard_inputs) -> dict of [B_local, ...] arrays` and gather
    every output.  `inputs`: dict of [B, ...] numpy arrays (every rank holds the full batch description,
    e.g. generated from the same seed).  Returns dict of [B, ...] numpy arrays, identical on all ranks."""
    lo, hi = shard_bounds(B, rank, world)
    local = solve_fn(**{k: v[lo:hi] for k, v in inputs.items()})
    out = {}
    for k, v in local.items():
        a = np.asarray(v)
        moved = np.ascontiguousarray(np.moveaxis(a, 0, -1))        # batch axis last, as in the C-ABI
        g = gather_shards(moved, B, world, rank, dist, device)
        out[k] = np.moveaxis(g.cpu().numpy(), -1, 0)
    return out
