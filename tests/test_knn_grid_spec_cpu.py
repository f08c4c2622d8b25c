"""CPU restatement of the search rule of etch_knn_grid (csrc/index.cu::knn_grid_query_kernel) against a brute-force kNN.

The CUDA kernel is checked bit for bit against the oracle on the GPU (tests/test_index_gpu.py); this test pins the RULE it
implements -- walk the cells shell by shell, stop when the k-th distance is strictly below the distance to everything
unexplored, hand every query with a tie inside the result or at its boundary to the exact emulation -- so that the three
properties the exactness argument needs hold on random and on lattice (tie-heavy) clouds:
  1. a query that is NOT flagged as a tie returns exactly the brute-force k nearest, ascending;
  2. every query whose k+1 smallest distances contain equal values IS flagged;
  3. queries far outside the candidates' bounding box are still answered exactly.
"""
import numpy as np

F = np.float32


def _grid(pts, k):
    lo, hi = pts.min(0), pts.max(0)
    e = np.maximum(hi - lo, F(1e-6))
    area = e[0] * e[1] + e[1] * e[2] + e[2] * e[0]
    h = F(1.1) * np.sqrt(F(k) * area / (F(np.pi) * F(len(pts))))
    h = max(h, e.max() / F(64))
    while True:
        dims = [min(64, int(e[a] / h) + 1) for a in range(3)]
        if dims[0] * dims[1] * dims[2] <= 65536:
            break
        h *= F(1.26)
    inv_h = F(1.0) / F(h)
    cc = lambda v, a: min(max(int(np.floor((v - lo[a]) * inv_h)), 0), dims[a] - 1)  # noqa: E731
    cells = {}
    for i, p in enumerate(pts):
        cells.setdefault((cc(p[0], 0), cc(p[1], 1), cc(p[2], 2)), []).append(i)
    return lo, F(h), dims, cc, cells


def _query(q, pts, k, grid):
    lo, h, dims, cc, cells = grid
    c = [cc(q[a], a) for a in range(3)]
    found = []          # (distance, index) of every examined candidate
    for r in range(max(dims) + 1):
        rng = [range(max(c[a] - r, 0), min(c[a] + r, dims[a] - 1) + 1) for a in range(3)]
        for z in rng[2]:
            for y in rng[1]:
                for x in rng[0]:
                    if max(abs(x - c[0]), abs(y - c[1]), abs(z - c[2])) != r:
                        continue
                    for i in cells.get((x, y, z), []):
                        d = pts[i] - q
                        found.append((float(F(d[0] * d[0]) + F(d[1] * d[1]) + F(d[2] * d[2])), i))
        if len(found) >= k:
            found.sort()
            dout = np.inf
            for a in range(3):
                if c[a] - r > 0:
                    dout = min(dout, q[a] - (lo[a] + (c[a] - r) * h))
                if c[a] + r + 1 < dims[a]:
                    dout = min(dout, lo[a] + (c[a] + r + 1) * h - q[a])
            if np.isinf(dout):
                break
            dout -= 1e-3 * h
            if dout > 0 and found[k - 1][0] < dout * dout:
                break
    found.sort()
    best = found[:k]
    rej = found[k][0] if len(found) > k else np.inf
    tie = rej <= best[-1][0] or any(best[i][0] == best[i - 1][0] for i in range(1, k))
    return [i for _, i in best], tie


def _brute(q, pts, k):
    d = pts - q
    d2 = (d[:, 0] * d[:, 0]).astype(F) + (d[:, 1] * d[:, 1]).astype(F) + (d[:, 2] * d[:, 2]).astype(F)
    order = np.argsort(d2, kind="stable")
    tie = any(d2[order[i]] == d2[order[i - 1]] for i in range(1, min(k + 1, len(pts))))
    return list(order[:k]), tie


def test_grid_walk_rule_is_exact_or_flags_a_tie():
    rng = np.random.default_rng(0)
    surface = rng.normal(size=(400, 3)).astype(F)
    surface /= np.linalg.norm(surface, axis=1, keepdims=True)
    surface *= np.array([0.3, 0.9, 0.2], dtype=F)                      # an elongated shell: surface-like sampling
    lattice = (rng.integers(-4, 5, size=(300, 3)).astype(F) * F(0.25))  # many exactly equal distances
    for pts in (surface, lattice):
        for k in (3, 8):
            grid = _grid(pts, k)
            queries = list(pts[::7]) + [pts[3] + F(5.0), pts[5] * F(0.5), np.array([9, -9, 9], dtype=F)]
            for q in queries:
                got, flagged = _query(q, pts, k, grid)
                ref, ref_tie = _brute(q, pts, k)
                if ref_tie:
                    assert flagged                                      # property 2
                if not flagged:
                    assert got == ref                                   # properties 1 and 3
