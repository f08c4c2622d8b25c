"""Edge cases of the boundary (what the reference kernels do at the corners of their domain), checked
against the C oracle / the reference's own kernels where a reference result exists, and the error behaviour of the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import index_ops as O

pytestmark = pytest.mark.gpu


def _ext():
    from etch_b200.ext import epn_grouping, epn_gathering, pointops_cuda
    return epn_grouping, epn_gathering, pointops_cuda


def test_minimal_sizes_index_kernels(cuda):
    g, ga, p = _ext()
    rng = np.random.default_rng(0)
    # FPS picking a single point / every point / from 2 points
    for n, m in ((2, 1), (2, 2), (33, 33), (64, 1)):
        x = rng.normal(size=(2, 3, n)).astype(np.float32)
        got = g.furthest_point_sampling(torch.from_numpy(x).to(cuda), m).cpu().numpy()
        np.testing.assert_array_equal(got, O.fps_bcn(x, m))
    # ball query: one support point, nsample 1, radius that contains nothing (index 0 fill) and everything
    x = rng.normal(size=(1, 3, 1)).astype(np.float32)
    for nn, r in ((1, 0.5), (4, 0.5), (4, 1e-9)):
        got = g.ball_query(torch.from_numpy(x).to(cuda), torch.from_numpy(x).to(cuda), r, nn).cpu().numpy()
        np.testing.assert_array_equal(got, O.ball_query_bcn(x, x, r, nn))
    x = rng.normal(size=(2, 3, 50)).astype(np.float32)
    got = g.ball_query(torch.from_numpy(x[:, :, :7].copy()).to(cuda), torch.from_numpy(x).to(cuda), 1e-6, 8).cpu().numpy()
    np.testing.assert_array_equal(got, O.ball_query_bcn(x[:, :, :7].copy(), x, 1e-6, 8))      # only the point itself: cyclic fill
    got = g.ball_query(torch.from_numpy(x).to(cuda), torch.from_numpy(x).to(cuda), 100.0, 8).cpu().numpy()
    assert (got == np.arange(8)).all()                                                          # everything in range: first 8 indices
    # kNN with a one-point segment and k larger than every segment
    xyz = rng.normal(size=(1 + 5 + 2, 3)).astype(np.float32)
    off = np.array([1, 6, 8], np.int32)
    for k in (3, 8, 16, 5):
        idx = torch.zeros(8, k, dtype=torch.int32, device=cuda)
        d2 = torch.zeros(8, k, dtype=torch.float32, device=cuda)
        t, o = torch.from_numpy(xyz).to(cuda), torch.from_numpy(off).to(cuda)
        p.knnquery_cuda(8, k, t, t, o, o, idx, d2)
        ri, rd = O.knn_packed(k, xyz, xyz, off, off)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri)
        np.testing.assert_array_equal(d2.cpu().numpy(), rd)


def test_fps_with_all_points_near_the_origin(cuda):
    """vgtk FPS skips points with |p|^2 <= 1e-3 (grouping_cuda_kernel.cu:384-387): a cloud that lies entirely inside that ball keeps
    picking index 0, exactly as the reference does."""
    g, _, _ = _ext()
    rng = np.random.default_rng(1)
    x = (rng.normal(size=(1, 3, 300)) * 0.005).astype(np.float32)
    got = g.furthest_point_sampling(torch.from_numpy(x).to(cuda), 40).cpu().numpy()
    np.testing.assert_array_equal(got, O.fps_bcn(x, 40))
    assert (got == 0).all()


def test_largest_supported_cloud(cuda):
    """28672 points per scan is the documented maximum of etch_fps_bcn (include/etch_b200.h); one more is refused, not mis-computed."""
    from etch_b200 import _lib as L, synth
    g, _, _ = _ext()
    x = np.ascontiguousarray(synth.sample_real_scans(1, 28672, 3).transpose(0, 2, 1))
    got = g.furthest_point_sampling(torch.from_numpy(x).to(cuda), 64).cpu().numpy()
    np.testing.assert_array_equal(got, O.fps_bcn(x, 64))
    big = torch.zeros(1, 3, 28673, device=cuda)
    with pytest.raises(RuntimeError, match="invalid argument"):
        g.furthest_point_sampling(big, 8)


def test_c_abi_rejects_bad_arguments(cuda):
    from etch_b200 import _lib as L
    x = torch.zeros(1, 3, 16, device=cuda)
    idx = torch.zeros(1, 4, dtype=torch.int32, device=cuda)
    lib = L.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.etch_fps_bcn(None, 1, 16, 4, ctypes.c_void_p(idx.data_ptr()), st) == -1                  # null pointer
    assert lib.etch_fps_bcn(ctypes.c_void_p(x.data_ptr()), 0, 16, 4, ctypes.c_void_p(idx.data_ptr()), st) == -1   # empty batch
    assert lib.etch_knn_grid(4, 5, ctypes.c_void_p(x.data_ptr()), 16, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(idx.data_ptr()),
                             ctypes.c_void_p(idx.data_ptr()), 1, ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                             ctypes.c_void_p(x.data_ptr()), st) == -1                                     # nsample outside {3, 8, 16}
    with pytest.raises(RuntimeError):
        L.ptr(torch.zeros(3))                                                                             # CPU tensor
    with pytest.raises(RuntimeError):
        L.ptr(torch.zeros(4, 4, device=cuda).t())                                                         # non-contiguous
    torch.cuda.synchronize()


def test_marker_with_nan_confidence_gives_nan_not_a_fault(cuda):
    """a NaN confidence among the points of a label used to index out of bounds; the reference's torch.topk ranks NaN
    first and yields a NaN marker (fit_SMPL.py:38-51)."""
    import types
    from etch_b200.models import fit_SMPL
    rng = np.random.default_rng(2)
    N, M = 200, 86
    inner = torch.from_numpy(rng.normal(size=(1, N, 3)).astype(np.float32)).to(cuda)
    labels = torch.from_numpy(rng.integers(0, M, size=(1, N))).to(cuda)
    conf = torch.from_numpy((rng.random((1, N, 1)) * 0.5 + 0.4).astype(np.float32)).to(cuda)
    lab0 = int(labels[0, 0])
    conf[0, (labels[0] == lab0).nonzero()[:, 0], 0] = float("nan")          # every confidence of one label is NaN
    args = types.SimpleNamespace(markerset={str(i): i for i in range(M)})
    markers, valid = fit_SMPL.get_markers(args, inner, labels, conf)
    torch.cuda.synchronize()
    assert bool(valid[0, lab0]) and torch.isnan(markers[0, lab0]).all()
    others = [l for l in range(M) if l != lab0 and bool(valid[0, l])]
    assert torch.isfinite(markers[0, others]).all()


def test_small_scan_through_the_whole_path(cuda):
    """the smallest cloud the PointTransformer hierarchy supports (256 points: one point at the deepest level)"""
    import json
    import os
    import types
    from etch_b200 import smpl_model, synth
    from etch_b200.models.models_pointcloud import GT_network_equiv
    from etch_b200.runtime import ScanFitter
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ms = json.load(open(os.path.join(root, "etch_b200", "data", "superset_smpl.json")))
    net = GT_network_equiv(types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=ms))
    net.load_state_dict(synth.make_state_dict(1))
    net = net.to(cuda).eval()
    args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
    fit = ScanFitter(net, args, use_graph=False)(torch.from_numpy(synth.sample_real_scans(1, 256, 9)).to(cuda))
    torch.cuda.synchronize()
    assert fit["vertices"].shape == (1, 6890, 3) and fit["labels"].shape == (1, 256)
