"""Batch sharding of one sampling call across ranks (one process per GPU of one node).

Each face's chain is independent (GroupNorm and attention are per sample, unet.py:84,113-142),
so the batch is split contiguously, every rank runs its slice with NO per-step communication,
and the only exchange is one final gather of the [B/G,3,R,R] outputs.

Nothing on the data path goes through NCCL. The gather is a device->host concat: all ranks map ONE
pinned host buffer (POSIX shared memory, registered with the CUDA driver), rank g copies its slice
into rows [lo_g, hi_g) of it with an asynchronous D2H copy on its own stream, and nobody waits for
anybody between chains - a rank only synchronises its own stream when it needs the result.
torch.distributed (any backend; `gloo` is enough) is used for rank / world size and for the one
control-plane barrier that tells readers that every slice has landed.
"""
import inspect
import mmap
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous split; the first (batch % world) ranks take one extra sample."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class HostGather:
    """The one final gather: a [B, ...] fp32 host buffer shared by the ranks of a node.

    `put(rows_lo, tensor)` = asynchronous D2H of this rank's slice into its rows (pinned: the segment is
    cudaHostRegister-ed when a CUDA device is in use); `full()` = the concatenation, valid once every
    rank's copy has completed and a barrier has passed (`wait()` does both). CPU tensors work too
    (plain memcpy), which is what the gloo tests exercise.
    """

    def __init__(self, shape, name=None, group=None, pin=None):
        self.group = group
        self.shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(self.shape)) * 4
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        tag = name or "b200sr3_gather_%s_%s" % (os.environ.get("MASTER_PORT", "0"), "x".join(map(str, self.shape)))
        self.path = os.path.join("/dev/shm", tag)
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(max(nbytes, 1))
        self._barrier()                                   # the segment exists before anyone maps it
        self._file = open(self.path, "r+b")
        self._map = mmap.mmap(self._file.fileno(), max(nbytes, 1))
        self.buf = torch.frombuffer(self._map, dtype=torch.float32, count=int(np.prod(self.shape))).view(self.shape)
        self._registered = False
        pin = torch.cuda.is_available() if pin is None else pin
        if pin and nbytes:
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"b200sr3: cudaHostRegister of the gather buffer failed ({rc})")
            self._registered = True
        self._barrier()
        if self.rank == 0:                                # every rank holds a mapping: the name can go
            os.unlink(self.path)

    def _barrier(self):
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)

    def put(self, lo, local):
        self.buf[lo:lo + local.shape[0]].copy_(local, non_blocking=True)

    def wait(self):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        self._barrier()

    def full(self):
        return self.buf

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self._registered = False
        self.buf = None
        try:
            self._map.close()
            self._file.close()
        except (BufferError, ValueError):
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_sample(sample_fn, cond, noise=None, group=None, gather=True, host_gather=None):
    """Run `sample_fn(cond_slice, noise_slice, row_offset)` on this rank's slice of the batch.

    cond: [B,3,R,R] (same on every rank); noise: optional [T,B,3,R,R]. `row_offset` is the slice's
    start row: a sampler that draws its own noise must key it by GLOBAL row (and every rank must use
    the same seed), so that the sharded batch draws exactly the noise of the unsharded one
    (GaussianDiffusion.super_resolution_batched(..., seed=s, row_offset=lo)). A two-argument
    `sample_fn(cond_slice, noise_slice)` is accepted for samplers with injected noise only.
    Returns the full [B,3,R,R] result as a HOST tensor on every rank when gather=True (one D2H concat
    into a shared pinned buffer, no collective library), else this rank's slice where sample_fn left it.
    """
    params = list(inspect.signature(sample_fn).parameters.values())
    takes_row = (any(p.kind == p.VAR_POSITIONAL for p in params)
                 or sum(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in params) >= 3)

    def call(c, z, lo):
        return sample_fn(c, z, lo) if takes_row else sample_fn(c, z)

    if not (dist.is_available() and dist.is_initialized()):
        out = call(cond, noise, 0)
        return out.cpu() if gather else out
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(cond.shape[0], rank, world)
    local = call(cond[lo:hi], None if noise is None else noise[:, lo:hi], lo)
    if not gather:
        return local
    hg = host_gather or HostGather((cond.shape[0],) + tuple(local.shape[1:]), group=group,
                                   pin=local.device.type == "cuda")
    hg.put(lo, local)
    hg.wait()
    out = hg.full().clone()
    if host_gather is None:
        hg._barrier()                                     # everyone has read before the segment is unmapped
        hg.close()
    return out
