"""Scan sharding across the GPUs of one box: scans are independent, so each rank gets a contiguous slice
of the batch; the only collectives are the input scatter and the result gather (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_scans(scans, total, item_shape, rank, world, device, dtype=torch.float32):
    """rank 0 holds scans [total, *item_shape]; every rank returns its slice on `device`."""
    lo, hi = shard_range(total, rank, world)
    if world == 1:
        return scans[lo:hi].to(device)
    local = torch.empty((hi - lo,) + tuple(item_shape), dtype=dtype, device=device)
    if rank == 0:
        reqs = []
        for r in range(1, world):
            l, h = shard_range(total, r, world)
            if h > l:
                reqs.append(dist.isend(scans[l:h].to(device).contiguous(), dst=r))
        local.copy_(scans[lo:hi])
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(local, src=0)
    return local


def gather_results(local, total, rank, world):
    """inverse of scatter_scans for per-scan results [n_local, ...]; returns the full tensor on rank 0 (None elsewhere)."""
    if world == 1:
        return local
    if rank == 0:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        lo, hi = shard_range(total, 0, world)
        out[lo:hi] = local
        for r in range(1, world):
            l, h = shard_range(total, r, world)
            if h > l:
                dist.recv(out[l:h], src=r)
        return out
    if local.shape[0] > 0:
        dist.send(local.contiguous(), dst=0)
    return None


def max_over_ranks(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
