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


RESULT_WIDTH = 6890 * 3 + 85 + 45 * 3     # floats per scan in a packed result row: vertices | params (orient, pose, betas, transl) | joints


def pack_results(fit):
    """fit dict of one batch (etch_b200.runtime.ScanFitter output) -> [B, RESULT_WIDTH] contiguous rows, so that the result
    gather is ONE message per rank (verts + pose + betas + transl + joints, ~84 KB per scan)."""
    B = fit["vertices"].shape[0]
    return torch.cat([fit["vertices"].reshape(B, -1), fit["params"].reshape(B, -1), fit["joints"].reshape(B, -1)], 1).contiguous()


def unpack_results(rows):
    """inverse of pack_results for [..., RESULT_WIDTH] rows -> dict(vertices [...,6890,3], params [...,85], joints [...,45,3])."""
    lead = rows.shape[:-1]
    nv = 6890 * 3
    return dict(vertices=rows[..., :nv].reshape(*lead, 6890, 3), params=rows[..., nv:nv + 85],
                joints=rows[..., nv + 85:].reshape(*lead, 45, 3))


def scatter_batch(global_batch, local_out, rank, world):
    """One collective: rank 0's [world*B, ...] batch (same device type as local_out) -> every rank's [B, ...] slice.
    NCCL runs it as one grouped send/recv on the current stream; gloo on CPU tensors in the tests."""
    if world == 1:
        local_out.copy_(global_batch)
        return local_out
    chunks = [c.contiguous() for c in global_batch.chunk(world, 0)] if rank == 0 else None
    dist.scatter(local_out, chunks, src=0)
    return local_out


def gather_rows(local_rows, global_rows, rank, world):
    """One collective: every rank's [B, W] rows -> rank 0's [world, B, W] buffer (None elsewhere is fine)."""
    if world == 1:
        global_rows[0].copy_(local_rows)
        return global_rows
    dist.gather(local_rows, [global_rows[r] for r in range(world)] if rank == 0 else None, dst=0)
    return global_rows


def max_over_ranks(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
