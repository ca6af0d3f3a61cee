"""Batch sharding of one sampling call across ranks (one process per GPU).

Each face's chain is independent (GroupNorm and attention are per sample, unet.py:84,113-142),
so the batch is split contiguously, every rank runs its slice with NO per-step communication,
and the only exchange is one final gather of the [B/G,3,R,R] outputs.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous split; the first (batch % world) ranks take one extra sample."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_sample(sample_fn, cond, noise=None, group=None, gather=True):
    """Run `sample_fn(cond_slice, noise_slice)` on this rank's slice of the batch.

    cond: [B,3,R,R] (same on every rank); noise: optional [T,B,3,R,R].
    Returns the full [B,3,R,R] result on every rank when gather=True (one all_gather of the
    final images), else this rank's slice.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(cond, noise)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(cond.shape[0], rank, world)
    local = sample_fn(cond[lo:hi], None if noise is None else noise[:, lo:hi])
    if not gather:
        return local
    sizes = [shard_bounds(cond.shape[0], r, world) for r in range(world)]
    parts = [torch.empty((b - a,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for a, b in sizes]
    if all(p.shape == parts[0].shape for p in parts):
        dist.all_gather(parts, local.contiguous(), group=group)
    else:   # ragged split: pad to the largest slice
        m = max(b - a for a, b in sizes)
        pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        padded = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(padded, pad, group=group)
        parts = [p[: b - a] for p, (a, b) in zip(padded, sizes)]
    return torch.cat(parts, dim=0)
