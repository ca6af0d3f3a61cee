"""Multi-GPU path on CPU: world_size-2 gloo, contiguous batch split, one final gather (a shared host buffer:
no collective library on the data path; torch.distributed only carries rank/world and the control barrier)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200sr3.sharding import shard_bounds, sharded_sample


def test_shard_bounds_partition():
    for batch in (1, 2, 5, 8, 32, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cond = torch.arange(batch * 3 * 4 * 4, dtype=torch.float32).view(batch, 3, 4, 4)
    noise = torch.ones(5, batch, 3, 4, 4) * torch.arange(batch).view(1, batch, 1, 1, 1)
    seen = []

    def fake_sampler(c, z):                      # stands in for netG.super_resolution_batched
        seen.append(c.shape[0])
        return c * 2 + z[0]

    full = sharded_sample(fake_sampler, cond, noise)
    local = sharded_sample(fake_sampler, cond, noise, gather=False)
    lo, hi = shard_bounds(batch, rank, world)
    ok = torch.equal(full, cond * 2 + noise[0]) and torch.equal(local, (cond * 2 + noise[0])[lo:hi]) and seen[0] == hi - lo
    # a sampler that draws its own noise is told its slice's global start row (the Philox row offset)
    rows = []

    def row_sampler(c, z, row_offset):
        rows.append(row_offset)
        return c + row_offset

    full2 = sharded_sample(row_sampler, cond, None)
    want = torch.cat([cond[a:b] + a for a, b in (shard_bounds(batch, r, world) for r in range(world))])
    ok = ok and rows == [lo] and torch.equal(full2, want)
    # a persistent gather buffer reused over several chains (what bench.py does), no barrier between puts
    from b200sr3.sharding import HostGather
    hg = HostGather((batch, 3, 4, 4), pin=False)
    for k in range(3):
        hg.put(lo, cond[lo:hi] * (k + 1))
    hg.wait()
    ok = ok and torch.equal(hg.full(), cond * 3)
    hg.wait()
    hg.close()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run(batch):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert out == [(0, True), (1, True)]


def test_world2_even_split():
    _run(4)


def test_world2_ragged_split():
    _run(5)
