"""Mixed-size scan streams (BASELINE.json configs[4]; SURVEY.md §8e "work queue for the mixed-N stream, longest-first").

Scans of a stream have different point counts.  Every kernel of the hot path is static-shape per (batch, N), so the stream is
cut into batches of scans with the SAME point count (one CUDA graph per (batch size, N)), the batches are ordered longest
first and dealt to the ranks with the classic LPT rule (each batch goes to the currently least loaded rank).  Nothing else
crosses ranks: a scan is an independent unit from its points to its fitted mesh.

  plan_stream(sizes, world, batch)  ->  [ [ (N, [scan ids]), ... ]  for each rank ]        (pure host logic, CPU-testable)
  run_stream(fitter, scans, my_batches, device) -> {scan id: {vertices, joints, params, labels}}   (one rank's share)
"""
import torch


def batch_cost(n_scans, n_points):
    """Relative cost model of one batch: the encoder, heads and fit are linear in the points; FPS / kNN / ball query are
    quadratic but stay below 5 % of the work up to 20k points, so they enter with a small weight."""
    return n_scans * (n_points + 2.5e-6 * n_points * n_points)


def plan_stream(sizes, world, batch):
    """sizes[i] = point count of scan i.  Returns per-rank lists of (N, scan ids) batches (ids in stream order within a batch),
    each with at most `batch` scans, assigned longest-processing-time first."""
    if world < 1 or batch < 1:
        raise ValueError("world and batch must be positive")
    by_n = {}
    for i, n in enumerate(sizes):
        if n <= 0:
            raise ValueError("scan %d has no points" % i)
        by_n.setdefault(int(n), []).append(i)
    batches = []
    for n in sorted(by_n, reverse=True):
        ids = by_n[n]
        for k in range(0, len(ids), batch):
            batches.append((n, ids[k:k + batch]))
    batches.sort(key=lambda b: (-batch_cost(len(b[1]), b[0]), b[1][0]))
    load = [0.0] * world
    plan = [[] for _ in range(world)]
    for n, ids in batches:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append((n, ids))
        load[r] += batch_cost(len(ids), n)
    return plan


@torch.no_grad()
def run_stream(fitter, scans, my_batches, device, pad_to=None):
    """scans: {id or index: [N,3] float32 tensor (host or device)}.  Submits this rank's batches back to back (they overlap
    through the fitter's in-flight slots) and returns per-scan results on `device`.  `pad_to` repeats the last scan of a short
    batch so that every batch of a given N replays the same graph (one capture per point count)."""
    out, pending = {}, []

    def drain(item):
        ids, ticket, nreal = item
        res = ticket.result()
        for k, sid in enumerate(ids[:nreal]):
            out[sid] = {key: res[key][k].clone() for key in ("vertices", "joints", "params", "labels")}

    for n, ids in my_batches:
        rows = [scans[i] for i in ids]
        nreal = len(rows)
        if pad_to is not None and nreal < pad_to:
            rows = rows + [rows[-1]] * (pad_to - nreal)
        pts = torch.stack([r.to(device, non_blocking=True) for r in rows], 0).contiguous()
        pending.append((ids, fitter.submit(pts), nreal))
        if len(pending) >= max(1, fitter.in_flight):   # a slot is reused in_flight submits later: consume its result first
            drain(pending.pop(0))
    while pending:
        drain(pending.pop(0))
    torch.cuda.current_stream().synchronize()
    return out
