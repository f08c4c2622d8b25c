"""Mesh -> point cloud (src/inference_demo.py:19-39, src/data_utils/GT_dataloader.py:100-102):
the CUDA path through the C ABI against the numpy restatement in oracle/mesh_sample.py -- bit-exact (float64 arithmetic,
same roundings in the same order; face indices are integer work), on the reference's in-tree SMPL body mesh, on a synthetic
mesh with degenerate / tiny faces, and at 80k faces (the size of the in-tree scan)."""
import os

import numpy as np
import pytest
import torch

from oracle import mesh_sample as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _smpl_mesh():
    g = np.load(os.path.join(ROOT, "tests", "golden", "mesh_smpl_00122.npz"))
    return g["vertices"], g["faces"]


def _random_mesh(V, F, seed, degenerate=False):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(V, 3)) * np.array([0.3, 0.9, 0.2])
    f = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    if degenerate:
        f[::7, 1] = f[::7, 0]                 # zero-area faces (repeated vertex): never picked unless pick == cum exactly
        v[f[3]] = v[f[3, 0]] + 1e-9 * rng.normal(size=(3, 3))    # a tiny face
    return v, f


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["smpl", "random", "degenerate", "80k"])
def test_sample_surface_bit_exact(cuda, case):
    from etch_b200 import mesh
    if case == "smpl":
        v, f = _smpl_mesh()
    elif case == "80k":
        v, f = _random_mesh(40002, 80000, 3)
    else:
        v, f = _random_mesh(500, 1500, 1, degenerate=(case == "degenerate"))
    count = 5000 if case != "80k" else 20000
    for seed in (15, 16):            # GT_dataloader.py:102 uses seed = self.seed + 15
        u = O.draws(count, seed)
        ref_p, ref_f = O.sample_surface(v, f, *u)
        tv, tf = torch.from_numpy(v).to(cuda), torch.from_numpy(f).to(cuda)
        p, fi, p32 = mesh.sample_surface(tv, tf, count, seed=seed, want_float32=True)      # draws its own numbers from the same seed
        np.testing.assert_array_equal(fi.cpu().numpy(), ref_f)
        np.testing.assert_array_equal(p.cpu().numpy(), ref_p)                              # float64, bit for bit
        np.testing.assert_array_equal(p32.cpu().numpy(), ref_p.astype(np.float32))         # what torch.from_numpy(points).float() gives


@pytest.mark.gpu
def test_preprocess_scan_bit_exact(cuda):
    from etch_b200 import mesh
    v, _ = _smpl_mesh()
    v = v + np.array([0.3, -1.1, 2.0])
    ref_v, ref_c = O.preprocess_scan(v)
    cv, c = mesh.preprocess_scan(v, cuda)
    np.testing.assert_array_equal(c, ref_c)
    np.testing.assert_array_equal(cv.cpu().numpy(), ref_v)


@pytest.mark.gpu
def test_scan_to_points_feeds_the_network(cuda, tmp_path):
    """OBJ file -> centred, sampled float32 cloud on the device -> network forward, without a CPU hop for the geometry."""
    from etch_b200 import mesh
    v, f = _smpl_mesh()
    path = tmp_path / "scan.obj"
    with open(path, "w") as fh:
        fh.write("# test\n")
        for p in v:
            fh.write("v %.8f %.8f %.8f\n" % tuple(p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))
    lv, lf = mesh.load_obj(str(path))
    np.testing.assert_array_equal(lf, f)
    assert np.abs(lv - v).max() < 1e-8
    pts, centre = mesh.scan_to_points(str(path), 5000, cuda, seed=3)
    ref_v, ref_c = O.preprocess_scan(lv)
    ref_p, _ = O.sample_surface(ref_v, lf, *O.draws(5000, 3))
    np.testing.assert_array_equal(centre, ref_c)
    np.testing.assert_array_equal(pts[0].cpu().numpy(), ref_p.astype(np.float32))
    assert pts.shape == (1, 5000, 3) and pts.dtype == torch.float32


def test_oracle_sampling_properties():
    """CPU: the restated sampler puts every point on its face (barycentric coordinates in [0,1], summing to 1) and picks faces
    in proportion to their area."""
    v, f = _smpl_mesh()
    u = O.draws(20000, 1)
    p, fi = O.sample_surface(v, f, *u)
    a, b, c = v[f[fi, 0]], v[f[fi, 1]], v[f[fi, 2]]
    T = np.stack([b - a, c - a], -1)                               # [n,3,2]
    sol = np.stack([np.linalg.lstsq(T[i], p[i] - a[i], rcond=None)[0] for i in range(0, 20000, 40)])
    assert (sol > -1e-9).all() and (sol.sum(1) < 1 + 1e-9).all()
    area = O.face_areas(v, f)
    big = area > np.quantile(area, 0.5)
    frac = big[fi].mean()
    assert abs(frac - area[big].sum() / area.sum()) < 0.02
    cv, centre = O.preprocess_scan(v)
    assert np.allclose(cv.min(0) + cv.max(0), 0, atol=1e-12) and np.allclose(cv + centre, v)


@pytest.mark.gpu
def test_gt_vectors_block_bit_exact(cuda):
    """GT_dataloader.py:104-124 (nearest info point, closest point on the SMPL mesh, vector assembly) on the in-tree SMPL body
    mesh: the CUDA path == the numpy restatement, bit for bit, and the closest points are true minima."""
    from etch_b200 import mesh
    v, f = _smpl_mesh()
    rng = np.random.default_rng(5)
    n = 600
    u = O.draws(n, 4)
    surf, _ = O.sample_surface(v, f, *u)
    pts = surf + rng.normal(size=(n, 3)) * 0.02              # a clothed-scan-like cloud around the body
    pts[:5] = surf[:5]                                        # points exactly on the surface
    pts[5] = v[f[10, 0]]                                      # a mesh vertex (shared by several faces: tie -> same closest point)
    info_points = pts[::2] + rng.normal(size=(len(pts[::2]), 3)) * 0.006      # ray-cast info points lie near every other sample
    info_vectors = rng.normal(size=info_points.shape) * 0.03
    tv, tf = torch.from_numpy(v).to(cuda), torch.from_numpy(f).to(cuda)
    tp = torch.from_numpy(pts).to(cuda)
    ref_c, ref_d, ref_f = O.closest_point(v, f, pts)
    c, d, fi = mesh.closest_point(tv, tf, tp)
    np.testing.assert_array_equal(c.cpu().numpy(), ref_c)
    np.testing.assert_array_equal(d.cpu().numpy(), ref_d)
    same_face = fi.cpu().numpy() == ref_f
    assert same_face.mean() > 0.99                            # equal distances through adjacent faces may pick either; the point is identical
    assert ref_d[:6].max() < 1e-12
    rd, ri = O.nearest_point(info_points, pts)
    gd, gi = mesh.nearest_point(torch.from_numpy(info_points).to(cuda), tp)
    np.testing.assert_array_equal(gi.cpu().numpy(), ri)
    np.testing.assert_array_equal(gd.cpu().numpy(), rd)
    ref_v = O.gt_vectors(pts, info_points, info_vectors, v, f)
    got_v = mesh.gt_vectors(tp, torch.from_numpy(info_points).to(cuda), torch.from_numpy(info_vectors).to(cuda), tv, tf)
    np.testing.assert_array_equal(got_v.cpu().numpy(), ref_v)
    assert 0.05 < (rd < 0.01).mean() < 0.95                   # both branches of the assembly are exercised


def test_oracle_closest_point_is_the_true_minimum():
    """CPU: the restated projection is no farther than 20000 random barycentric samples of every nearby face, and matches scipy's
    exact KD-tree for the nearest info point."""
    from scipy.spatial import cKDTree
    v, f = _smpl_mesh()
    rng = np.random.default_rng(2)
    pts = v[rng.integers(0, len(v), 40)] + rng.normal(size=(40, 3)) * 0.03
    c, d, fi = O.closest_point(v, f, pts)
    bary = rng.dirichlet([1, 1, 1], size=300)
    tri = v[f]                                                 # [F,3,3]
    for i in range(40):
        near = np.argsort(np.linalg.norm(tri.mean(1) - pts[i], axis=1))[:60]
        cand = np.einsum("sk,fkc->fsc", bary, tri[near]).reshape(-1, 3)
        assert d[i] <= np.linalg.norm(cand - pts[i], axis=1).min() + 1e-12
        assert abs(np.linalg.norm(c[i] - pts[i]) - d[i]) < 1e-12
    ref = v[::5]
    dd, ii = O.nearest_point(ref, pts)
    sd, si = cKDTree(ref).query(pts, k=1)
    assert (ii == si).all() and np.allclose(dd, sd, rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_results_to_host_is_one_packed_copy_of_the_same_values(cuda):
    """etch_b200.io.results_to_host (the per-batch D2H eval.py needs before its writers) returns exactly the device values."""
    import json
    import types
    from etch_b200 import io, smpl_model, synth
    from etch_b200.models.models_pointcloud import GT_network_equiv
    from etch_b200.runtime import ScanFitter
    ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
    net = GT_network_equiv(types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=ms))
    net.load_state_dict(synth.make_state_dict(1))
    net = net.to(cuda).eval()
    args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
    fit = ScanFitter(net, args, use_graph=False)(torch.from_numpy(synth.sample_real_scans(2, 1024, 4)).to(cuda))
    host = io.results_to_host(fit)
    for k in ("vertices", "joints", "params", "tightness", "inner", "markers", "confidences"):
        np.testing.assert_array_equal(host[k], fit[k].cpu().numpy())
    np.testing.assert_array_equal(host["labels"], fit["labels"].cpu().numpy())
    np.testing.assert_array_equal(host["valid"], fit["valid"].cpu().numpy())
    assert host["labels"].dtype == np.int64 and host["valid"].dtype == bool
