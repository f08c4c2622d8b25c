"""GPU parity of the index kernels (through the drop-in extension modules -> C ABI) against the C oracle.

Bit-exact: FPS (both variants, including the tie rule and the vgtk origin skip), ball query (ordered first-K with
cyclic padding and the nsample-1 quirk), gather, kNN (indices and squared distances, heap tie order)."""
import numpy as np
import pytest
import torch

from oracle import index_ops as O

pytestmark = pytest.mark.gpu


def _ext():
    from etch_b200.ext import epn_grouping, epn_gathering, pointops_cuda
    return epn_grouping, epn_gathering, pointops_cuda


def _scan_bcn(B, n, seed):
    from etch_b200 import synth
    return np.ascontiguousarray(synth.sample_scans(B, n, seed).transpose(0, 2, 1))


def _tie_cloud(B, n, seed):
    """lattice points + exact duplicates: many equal distances -> exercises the arg-max / heap tie rules."""
    rng = np.random.default_rng(seed)
    g = rng.integers(-6, 7, size=(B, n, 3)).astype(np.float32) * 0.125
    g[:, n // 2:] = g[:, : n - n // 2]  # duplicates
    g[:, 5] = 0.0  # a point at the origin (vgtk skip)
    g[:, 7] = [0.01, 0.02, 0.01]  # |p|^2 = 6e-4 <= 1e-3 -> skipped too
    return np.ascontiguousarray(g.transpose(0, 2, 1))


@pytest.mark.parametrize("B,n", [(1, 5000), (8, 5000), (2, 10000), (2, 20000), (3, 777), (2, 64), (1, 1500), (2, 8192), (3, 8193), (1, 28672)])
def test_fps_bcn_matches_oracle(cuda, B, n):
    g, _, _ = _ext()
    x = _scan_bcn(B, n, 11)
    m = (n + 1) // 2
    got = g.furthest_point_sampling(torch.from_numpy(x).to(cuda), m).cpu().numpy()
    np.testing.assert_array_equal(got, O.fps_bcn(x, m))


@pytest.mark.parametrize("n", [96, 1000, 2048, 3000])
def test_fps_bcn_tie_rule_and_origin_skip(cuda, n):
    g, _, _ = _ext()
    x = _tie_cloud(2, n, 5)
    m = n // 2
    got = g.furthest_point_sampling(torch.from_numpy(x).to(cuda), m).cpu().numpy()
    np.testing.assert_array_equal(got, O.fps_bcn(x, m))


@pytest.mark.parametrize("n,radius,nn", [(5000, 0.08, 64), (2500, 0.11313708498984763, 32), (2500, 0.16, 64), (1250, 0.16, 32),
                                          (20000, 0.08, 64)])
def test_ball_query_matches_oracle(cuda, n, radius, nn):
    g, ga, _ = _ext()
    x = _scan_bcn(2, n, 3)
    m = n // 2 if nn == 64 else n
    q = np.ascontiguousarray(x[:, :, :m])
    got = g.ball_query(torch.from_numpy(q).to(cuda), torch.from_numpy(x).to(cuda), radius, nn).cpu().numpy()
    np.testing.assert_array_equal(got, O.ball_query_bcn(q, x, radius, nn))


def test_ball_query_padding_quirks(cuda):
    """cnt == nsample-1 leaves a zero in the last slot; cnt < nsample-1 repeats cyclically (grouping_cuda_kernel.cu:96-102)."""
    g, _, _ = _ext()
    rng = np.random.default_rng(0)
    x = (rng.random((1, 3, 400)).astype(np.float32) - 0.5)
    for nn in (2, 3, 5, 8, 33):
        for r in (0.05, 0.12, 0.3):
            got = g.ball_query(torch.from_numpy(x).to(cuda), torch.from_numpy(x).to(cuda), r, nn).cpu().numpy()
            np.testing.assert_array_equal(got, O.ball_query_bcn(x, x, r, nn))


def test_gather_matches_oracle(cuda):
    _, ga, _ = _ext()
    rng = np.random.default_rng(1)
    p = rng.normal(size=(3, 7, 501)).astype(np.float32)
    idx = rng.integers(0, 501, size=(3, 1234)).astype(np.int32)
    got = ga.gather_points_forward(torch.from_numpy(p).to(cuda), torch.from_numpy(idx).to(cuda)).cpu().numpy()
    np.testing.assert_array_equal(got, O.gather_bcn(p, idx))


def _packed(B, n, seed, ties=False):
    if ties:
        x = _tie_cloud(B, n, seed).transpose(0, 2, 1)
    else:
        from etch_b200 import synth
        x = synth.sample_scans(B, n, seed)
    return np.ascontiguousarray(x.reshape(-1, 3)), (np.arange(1, B + 1) * n).astype(np.int32)


@pytest.mark.parametrize("k", [3, 8, 16, 5])
@pytest.mark.parametrize("ties", [False, True])
def test_knn_self_matches_oracle(cuda, k, ties):
    _, _, p = _ext()
    xyz, off = _packed(3, 1000, 2, ties)
    m = xyz.shape[0]
    idx = torch.zeros(m, k, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(m, k, dtype=torch.float32, device=cuda)
    t = torch.from_numpy(xyz).to(cuda)
    o = torch.from_numpy(off).to(cuda)
    p.knnquery_cuda(m, k, t, t, o, o, idx, d2)
    ri, rd = O.knn_packed(k, xyz, xyz, off, off)
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(d2.cpu().numpy(), rd)


def test_knn_cross_levels_and_short_segments(cuda):
    """queries from a coarser level; a segment shorter than k keeps (1e10, start) fillers like the reference."""
    _, _, p = _ext()
    rng = np.random.default_rng(4)
    seg = [700, 9, 300]
    xyz = rng.normal(size=(sum(seg), 3)).astype(np.float32)
    off = np.cumsum(seg).astype(np.int32)
    nseg = [s // 4 for s in seg]
    noff = np.cumsum(nseg).astype(np.int32)
    q = np.concatenate([xyz[s0:s0 + c] for s0, c in zip(np.concatenate([[0], off[:-1]]), nseg)], 0)
    for k in (3, 16):
        idx = torch.zeros(len(q), k, dtype=torch.int32, device=cuda)
        d2 = torch.zeros(len(q), k, dtype=torch.float32, device=cuda)
        p.knnquery_cuda(len(q), k, torch.from_numpy(xyz).to(cuda), torch.from_numpy(q).to(cuda),
                        torch.from_numpy(off).to(cuda), torch.from_numpy(noff).to(cuda), idx, d2)
        ri, rd = O.knn_packed(k, xyz, q, off, noff)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri)
        np.testing.assert_array_equal(d2.cpu().numpy(), rd)


@pytest.mark.parametrize("segs,ties", [([5000] * 4, False), ([1250, 1250], False), ([312, 312, 312], True), ([78, 40, 19, 300], True),
                                        ([20000, 20000], False), ([10000, 333, 20000], False), ([9000, 12000], True)])
def test_fps_packed_matches_oracle(cuda, segs, ties):
    _, _, p = _ext()
    rng = np.random.default_rng(9)
    if ties:
        xyz = (rng.integers(-5, 6, size=(sum(segs), 3)).astype(np.float32) * 0.25)
    else:
        from etch_b200 import synth
        xyz = np.concatenate([synth.sample_scan(s, i) for i, s in enumerate(segs)], 0)
    off = np.cumsum(segs).astype(np.int32)
    noff = np.cumsum([s // 4 for s in segs]).astype(np.int32)
    idx = torch.zeros(int(noff[-1]), dtype=torch.int32, device=cuda)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=cuda)
    p.furthestsampling_cuda(len(segs), max(segs), torch.from_numpy(xyz).to(cuda), torch.from_numpy(off).to(cuda),
                            torch.from_numpy(noff).to(cuda), tmp, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), O.fps_packed(xyz, off, noff))


def _knn_grid(cuda, k, xyz, q, off, noff):
    import ctypes
    from etch_b200 import _lib as L
    n, m = xyz.shape[0], q.shape[0]
    fn = L.lib().etch_knn_grid_scratch_bytes
    fn.restype = ctypes.c_longlong
    scratch = torch.empty(int(fn(n, len(off))), dtype=torch.uint8, device=cuda)
    idx = torch.zeros(m, k, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(m, k, dtype=torch.float32, device=cuda)
    tx, tq, to, tn = (torch.from_numpy(a).to(cuda) for a in (xyz, q, off, noff))   # keep the device copies alive across the call
    L.call("knn_grid", m, k, L.ptr(tx), n, L.ptr(tq), L.ptr(to), L.ptr(tn), len(off), L.ptr(idx), L.ptr(d2), L.ptr(scratch))
    torch.cuda.synchronize()
    del tx, tq, to, tn
    return idx.cpu().numpy(), d2.cpu().numpy()


@pytest.mark.parametrize("k", [3, 8, 16])
@pytest.mark.parametrize("B,n,ties", [(3, 1000, False), (3, 1000, True), (2, 5000, False), (1, 20000, False)])
def test_knn_grid_matches_oracle(cuda, k, B, n, ties):
    """etch_knn_grid (uniform-grid search + exact fallback on ties) == the reference's brute-force heap scan, bit for bit."""
    xyz, off = _packed(B, n, 2, ties)
    gi, gd = _knn_grid(cuda, k, xyz, xyz, off, off)
    ri, rd = O.knn_packed(k, xyz, xyz, off, off)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(gd, rd)


def test_knn_grid_cross_levels_short_segments_and_outliers(cuda):
    """queries from another level, a segment shorter than k, queries far outside the candidates' bounding box"""
    rng = np.random.default_rng(4)
    seg = [700, 9, 300]
    xyz = rng.normal(size=(sum(seg), 3)).astype(np.float32)
    off = np.cumsum(seg).astype(np.int32)
    nseg = [s // 4 for s in seg]
    noff = np.cumsum(nseg).astype(np.int32)
    q = np.concatenate([xyz[s0:s0 + c] for s0, c in zip(np.concatenate([[0], off[:-1]]), nseg)], 0).copy()
    q[::7] += np.float32(9.0)      # far outside the grid
    q[3::11] *= np.float32(0.5)
    for k in (3, 8, 16):
        gi, gd = _knn_grid(cuda, k, xyz, q, off, noff)
        ri, rd = O.knn_packed(k, xyz, q, off, noff)
        np.testing.assert_array_equal(gi, ri)
        np.testing.assert_array_equal(gd, rd)
