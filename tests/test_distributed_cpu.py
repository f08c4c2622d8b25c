"""world_size-2 gloo test of the scan sharding / gather plumbing used by bench.py --gpus N (no GPU needed)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from etch_b200 import sharding
    total = 10
    lo, hi = sharding.shard_range(total, rank, world)
    scans = torch.arange(total * 6, dtype=torch.float32).view(total, 2, 3) if rank == 0 else None
    local = sharding.scatter_scans(scans, total, (2, 3), rank, world, torch.device("cpu"))
    assert local.shape[0] == hi - lo
    out = sharding.gather_results(local.sum(dim=(1, 2), keepdim=False).view(-1, 1), total, rank, world)
    t = torch.tensor([float(rank + 1)])
    mx = sharding.max_over_ranks(t)
    # the per-step collectives of bench.py --gpus N: one scatter of the global batch, one gather of packed result rows
    B = 3
    glob = torch.arange(world * B * 4 * 3, dtype=torch.float32).view(world * B, 4, 3) if rank == 0 else None
    mine = sharding.scatter_batch(glob, torch.empty(B, 4, 3), rank, world)
    assert torch.equal(mine, torch.arange(world * B * 12, dtype=torch.float32).view(world * B, 4, 3)[rank * B:(rank + 1) * B])
    fit = dict(vertices=torch.full((B, 6890, 3), float(rank)), params=torch.arange(B * 85, dtype=torch.float32).view(B, 85) + rank,
               joints=torch.full((B, 45, 3), 10.0 + rank))
    rows = sharding.pack_results(fit)
    assert rows.shape == (B, sharding.RESULT_WIDTH)
    allrows = torch.empty(world, B, sharding.RESULT_WIDTH) if rank == 0 else None
    sharding.gather_rows(rows, allrows, rank, world)
    if rank == 0:
        un = sharding.unpack_results(allrows)
        ok = all(bool((un["vertices"][r] == float(r)).all()) and bool((un["joints"][r] == 10.0 + r).all())
                 and torch.equal(un["params"][r], torch.arange(B * 85, dtype=torch.float32).view(B, 85) + r) for r in range(world))
        assert ok
        q.put((out.view(-1).tolist(), mx.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.arange(60, dtype=torch.float32).view(10, 6).sum(1).tolist()
    assert res == expect and mx == 2.0


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from etch_b200 import sharding
    for total in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
