"""eval.py output writer: etch_write_points_vector_ply against (a) the bytes the UNMODIFIED reference
function utils.GT_utils.save_points_with_vector wrote (tests/golden/golden_points_vector.ply, tools/gen_golden.py) and (b) a
restatement of its Python loop under this interpreter's numpy; plus the float formatter against numpy / Python on random bit
patterns.  Host-only code: runs in the CPU suite (no compute kernels involved)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _reference_loop(hit_points, vectors, file_path):
    """src/utils/GT_utils.py:22-55, restated (the checker)."""
    n = len(hit_points)
    end = hit_points - vectors
    with open(file_path, "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write(f"element vertex {n * 2}\n")
        f.write("property float x\nproperty float y\nproperty float z\n")
        f.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        f.write(f"element edge {n}\n")
        f.write("property int vertex1\nproperty int vertex2\nend_header\n")
        for p in hit_points:
            f.write(f"{p[0]} {p[1]} {p[2]} 255 0 0\n")
        for v in end:
            f.write(f"{v[0]} {v[1]} {v[2]} 0 0 255\n")
        for i in range(n):
            f.write(f"{i} {n + i}\n")


def test_vector_ply_matches_the_reference_functions_bytes(tmp_path):
    from etch_b200 import io
    g = np.load(os.path.join(GOLD, "golden_points_vector_in.npz"))
    style = 0 if int(str(g["numpy_version"]).split(".")[0]) >= 2 else 1
    out = tmp_path / "v.ply"
    io.save_points_with_vector(g["hit"], g["vec"], str(out), style=style)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "golden_points_vector.ply"), "rb").read()


def test_vector_ply_matches_the_python_loop_in_this_interpreter(tmp_path):
    from etch_b200 import io
    rng = np.random.default_rng(3)
    for n in (0, 1, 777):
        hit = rng.normal(size=(n, 3)).astype(np.float32)
        vec = (rng.normal(size=(n, 3)) * 0.05).astype(np.float32)
        a, b = tmp_path / ("a%d.ply" % n), tmp_path / ("b%d.ply" % n)
        io.save_points_with_vector(hit, vec, str(a))
        _reference_loop(hit, vec, str(b))
        assert open(a, "rb").read() == open(b, "rb").read()


def test_float_formatter_matches_numpy_and_python_on_random_bit_patterns():
    from etch_b200 import io
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.integers(-2 ** 31, 2 ** 31, size=200000).astype(np.uint32).view(np.float32),
                        rng.normal(size=50000).astype(np.float32), (rng.normal(size=20000) * 1e-5).astype(np.float32),
                        np.array([0, -0.0, 1, 1e5, 1e6, 999999.9, 1e16, 1e-4, 9.99e-5, np.inf, -np.inf, np.nan, 1e-45, 3.4e38], np.float32)])
    assert io.format_np_float32(x, 1) == [str(v) for v in x]                 # numpy 1.x f-string form = str(np.float32)
    assert io.format_np_float32(x, 0) == [repr(float(v)) for v in x]         # numpy >= 2 f-string form = repr of the widened double


def test_npz_and_score_writers_follow_eval_py(tmp_path):
    """src/eval.py:136-145,233-262 restated next to the wrappers: same file names, keys, slices and text lines."""
    import torch
    from etch_b200 import io
    rng = np.random.default_rng(3)
    K = 50
    hit, pv, gv = (rng.standard_normal((K, 3)).astype(np.float32) for _ in range(3))
    pl, gl = rng.integers(0, 86, K), rng.integers(0, 86, K)
    pc, gc = rng.random((K, 1)).astype(np.float32), rng.random((K, 1)).astype(np.float32)
    out = str(tmp_path)
    p = io.save_tightness_vectors_info(out, "00122_x", torch.from_numpy(hit), torch.from_numpy(pv), torch.from_numpy(pl), torch.from_numpy(pc),
                                       gv, gl, gc)
    assert p == os.path.join(out, "00122_x", "tightness_vectors_info_00122_x.npz")
    z = np.load(p)
    assert sorted(z.files) == sorted(["hitpts", "pred_vectors", "pred_part_labels", "pred_confidences", "gt_vectors", "gt_labels", "gt_confidences"])
    assert np.array_equal(z["hitpts"], hit) and np.array_equal(z["pred_part_labels"], pl) and np.array_equal(z["gt_confidences"], gc)

    info = [rng.standard_normal((2, 23, 3)).astype(np.float32), rng.standard_normal((2, 10)).astype(np.float32),
            rng.standard_normal((2, 3)).astype(np.float32), rng.standard_normal((2, 3)).astype(np.float32),
            rng.standard_normal((2, 45, 3)).astype(np.float32)]
    z = np.load(io.save_output_smpl_info(out, "00122_x", info, 1))
    assert z["body_pose"].shape == (21, 3) and z["hand_pose"].shape == (2, 3) and z["joints"].shape == (45, 3)
    assert np.array_equal(z["body_pose"], info[0][1][:21]) and np.array_equal(z["hand_pose"], info[0][1][21:23])
    assert np.array_equal(z["betas"], info[1][1]) and np.array_equal(z["global_orient"], info[2][1]) and np.array_equal(z["transl"], info[3][1])

    gt, pr = rng.standard_normal((6890, 3)).astype(np.float32), rng.standard_normal((6890, 3)).astype(np.float32)
    v2v = io.v2v_score(gt, torch.from_numpy(pr))
    ref = np.mean(np.linalg.norm(gt.astype(np.float64) - pr.astype(np.float64), axis=1))      # eval.py:233-236 on trimesh's float64 arrays
    assert v2v == ref and isinstance(v2v, np.float64)
    full, part = torch.ones(86, dtype=torch.bool), torch.ones(86, dtype=torch.bool)
    part[5] = False
    io.append_v2v_score(out, "a", v2v, full)
    io.append_v2v_score(out, "b", v2v, part)
    io.append_v2v_summary(out, v2v * 2, 2)
    lines = open(os.path.join(out, "v2v_score.txt")).read().split("\n")
    assert lines[0] == f"a: {ref}" and lines[1] == f"b: {ref}  attention, the valid mask is not full"
    assert lines[2:6] == ["==========", f"average v2v: {ref * 2 / 2}", f"total v2v: {ref * 2}", "sample num: 2"]
