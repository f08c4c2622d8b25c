"""Pins the index oracle AND the product against the REFERENCE's own CUDA kernels (oracle/_ref/libetch_ref_kernels.so:
external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:67-113,351-466, gathering_cuda_kernel.cu:42-68,
external/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-108, sampling/sampling_cuda_kernel.cu:15-129, compiled
unmodified for sm_100a by oracle/build_ref.sh).  Three-way, bit-exact: reference kernel == oracle/etch_oracle.c == etch_*.
"""
import numpy as np
import pytest
import torch

from oracle import index_ops as O
from oracle import ref_kernels as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference at build time)")]


def _ext():
    from etch_b200.ext import epn_grouping, epn_gathering, pointops_cuda
    return epn_grouping, epn_gathering, pointops_cuda


def _scan_bcn(B, n, seed, real=False):
    from etch_b200 import synth
    x = synth.sample_real_scans(B, n, seed) if real else synth.sample_scans(B, n, seed)
    return np.ascontiguousarray(x.transpose(0, 2, 1))


def _tie_cloud(B, n, seed):
    rng = np.random.default_rng(seed)
    g = rng.integers(-6, 7, size=(B, n, 3)).astype(np.float32) * 0.125
    g[:, n // 2:] = g[:, : n - n // 2]
    g[:, 5] = 0.0
    g[:, 7] = [0.01, 0.02, 0.01]
    return np.ascontiguousarray(g.transpose(0, 2, 1))


@pytest.mark.parametrize("B,n,kind", [(2, 5000, "synth"), (2, 5000, "real"), (2, 10000, "synth"), (1, 20000, "real"), (3, 777, "synth"),
                                      (2, 2048, "ties"), (2, 3000, "ties")])
def test_vgtk_fps_three_way(cuda, B, n, kind):
    g, _, _ = _ext()
    x = _tie_cloud(B, n, 5) if kind == "ties" else _scan_bcn(B, n, 11, kind == "real")
    m = (n + 1) // 2
    t = torch.from_numpy(x).to(cuda)
    ref = R.furthest_point_sampling(t, m).cpu().numpy()
    np.testing.assert_array_equal(O.fps_bcn(x, m), ref)                                    # C oracle == reference kernel
    np.testing.assert_array_equal(g.furthest_point_sampling(t, m).cpu().numpy(), ref)      # product == reference kernel


@pytest.mark.parametrize("n,radius,nn,kind", [(5000, 0.08, 64, "synth"), (5000, 0.08, 64, "real"), (2500, 0.11313708498984763, 32, "real"),
                                              (2500, 0.16, 64, "synth"), (1250, 0.16, 32, "real"), (10000, 0.08, 64, "real"),
                                              (20000, 0.08, 64, "synth"), (400, 0.12, 5, "ties"), (400, 0.3, 33, "ties")])
def test_vgtk_ball_query_three_way(cuda, n, radius, nn, kind):
    g, _, _ = _ext()
    x = _tie_cloud(2, n, 3) if kind == "ties" else _scan_bcn(2, n, 3, kind == "real")
    m = n // 2 if nn == 64 else n
    q = np.ascontiguousarray(x[:, :, :m])
    tq, tx = torch.from_numpy(q).to(cuda), torch.from_numpy(x).to(cuda)
    ref = R.ball_query(tq, tx, radius, nn).cpu().numpy()
    np.testing.assert_array_equal(O.ball_query_bcn(q, x, radius, nn), ref)
    np.testing.assert_array_equal(g.ball_query(tq, tx, radius, nn).cpu().numpy(), ref)


def test_vgtk_gather_three_way(cuda):
    _, ga, _ = _ext()
    rng = np.random.default_rng(1)
    p = rng.normal(size=(3, 7, 501)).astype(np.float32)
    idx = rng.integers(0, 501, size=(3, 1234)).astype(np.int32)
    tp, ti = torch.from_numpy(p).to(cuda), torch.from_numpy(idx).to(cuda)
    ref = R.gather_points_forward(tp, ti).cpu().numpy()
    np.testing.assert_array_equal(O.gather_bcn(p, idx), ref)
    np.testing.assert_array_equal(ga.gather_points_forward(tp, ti).cpu().numpy(), ref)


def _packed(segs, seed, kind):
    from etch_b200 import synth
    rng = np.random.default_rng(seed)
    if kind == "ties":
        xyz = rng.integers(-5, 6, size=(sum(segs), 3)).astype(np.float32) * 0.25
    elif kind == "real":
        xyz = np.concatenate([synth.sample_real_scans(1, s, seed + i)[0] for i, s in enumerate(segs)], 0)
    else:
        xyz = np.concatenate([synth.sample_scan(s, seed + i) for i, s in enumerate(segs)], 0)
    return np.ascontiguousarray(xyz, np.float32), np.cumsum(segs).astype(np.int32)


@pytest.mark.parametrize("segs,kind", [([5000, 5000], "synth"), ([5000, 5000], "real"), ([1250, 1250], "real"), ([312, 312, 312], "ties"),
                                       ([78, 40, 19, 300], "ties"), ([10000, 333, 20000], "synth"), ([9000, 12000], "ties")])
def test_pointops_fps_three_way(cuda, segs, kind):
    _, _, p = _ext()
    xyz, off = _packed(segs, 9, kind)
    noff = np.cumsum([s // 4 for s in segs]).astype(np.int32)
    tx, to, tn = (torch.from_numpy(a).to(cuda) for a in (xyz, off, noff))
    ref = R.furthestsampling(tx, to, tn).cpu().numpy()
    np.testing.assert_array_equal(O.fps_packed(xyz, off, noff), ref)
    idx = torch.zeros(int(noff[-1]), dtype=torch.int32, device=cuda)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=cuda)
    p.furthestsampling_cuda(len(segs), max(segs), tx, to, tn, tmp, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref)


def _knn_grid(cuda, k, tx, tq, to, tn):
    import ctypes
    from etch_b200 import _lib as L
    n, m = tx.shape[0], tq.shape[0]
    fn = L.lib().etch_knn_grid_scratch_bytes
    fn.restype = ctypes.c_longlong
    scratch = torch.empty(int(fn(n, to.shape[0])), dtype=torch.uint8, device=cuda)
    idx = torch.zeros(m, k, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(m, k, dtype=torch.float32, device=cuda)
    L.call("knn_grid", m, k, L.ptr(tx), n, L.ptr(tq), L.ptr(to), L.ptr(tn), to.shape[0], L.ptr(idx), L.ptr(d2), L.ptr(scratch))
    torch.cuda.synchronize()
    return idx.cpu().numpy(), d2.cpu().numpy()


@pytest.mark.parametrize("k", [3, 8, 16])
@pytest.mark.parametrize("segs,kind", [([1000, 1000, 1000], "synth"), ([1000, 1000, 1000], "ties"), ([5000, 5000], "real"), ([10000], "real"),
                                       ([20000], "synth")])
def test_pointops_knn_three_way(cuda, k, segs, kind):
    """self-kNN: reference kernel == C oracle == etch_knn_packed (the binding's replacement) == etch_knn_grid (the product path)"""
    _, _, p = _ext()
    xyz, off = _packed(segs, 2, kind)
    tx, to = torch.from_numpy(xyz).to(cuda), torch.from_numpy(off).to(cuda)
    ri, rd = R.knnquery(k, tx, tx, to, to)
    ri, rd = ri.cpu().numpy(), rd.cpu().numpy()
    if sum(segs) <= 10000:   # the scalar C emulation is quadratic; keep the CPU side in seconds
        oi, od = O.knn_packed(k, xyz, xyz, off, off)
        np.testing.assert_array_equal(oi, ri)
        np.testing.assert_array_equal(od, rd)
    m = xyz.shape[0]
    idx = torch.zeros(m, k, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(m, k, dtype=torch.float32, device=cuda)
    from etch_b200 import _lib as L
    L.call("knn_packed", m, k, L.ptr(tx), L.ptr(tx), L.ptr(to), L.ptr(to), int(to.shape[0]), L.ptr(idx), L.ptr(d2))     # brute-force kernel
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(d2.cpu().numpy(), rd)
    idx.zero_(); d2.zero_()
    p.knnquery_cuda(m, k, tx, tx, to, to, idx, d2)                                                                       # the drop-in binding
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(d2.cpu().numpy(), rd)
    gi, gd = _knn_grid(cuda, k, tx, tx, to, to)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(gd, rd)


def test_pointops_knn_cross_level_three_way(cuda):
    """queries from the next-coarser level (TransitionDown / TransitionUp graphs), incl. a segment shorter than k"""
    _, _, p = _ext()
    rng = np.random.default_rng(4)
    seg = [700, 9, 300]
    xyz = rng.normal(size=(sum(seg), 3)).astype(np.float32)
    off = np.cumsum(seg).astype(np.int32)
    nseg = [s // 4 for s in seg]
    noff = np.cumsum(nseg).astype(np.int32)
    q = np.concatenate([xyz[s0:s0 + c] for s0, c in zip(np.concatenate([[0], off[:-1]]), nseg)], 0)
    tx, tq, to, tn = (torch.from_numpy(np.ascontiguousarray(a)).to(cuda) for a in (xyz, q, off, noff))
    for k in (3, 16):
        ri, rd = R.knnquery(k, tx, tq, to, tn)
        ri, rd = ri.cpu().numpy(), rd.cpu().numpy()
        oi, od = O.knn_packed(k, xyz, q, off, noff)
        np.testing.assert_array_equal(oi, ri)
        np.testing.assert_array_equal(od, rd)
        gi, gd = _knn_grid(cuda, k, tx, tq, to, tn)
        np.testing.assert_array_equal(gi, ri)
        np.testing.assert_array_equal(gd, rd)
