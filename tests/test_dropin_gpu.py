"""Boundary B1/B2 on the GPU: the reference callers' own call sequence, through the drop-in module NAMES
(`models.models_pointcloud`, `models.fit_SMPL`, `pointops_cuda`, `epn_grouping`, `epn_gathering`), in a fresh interpreter
whose PYTHONPATH holds only etch_b200/dropin and etch_b200/ext -- exactly what INTEGRATION.md tells a maintainer to do.

The script below restates src/inference_demo.py:12-17 (load_model), :41-66 (predict_smpl), :108-127 (un-centre, export, npz) and
the wrapper calls of src/models/pointops.py:10-45 and vgtk/pc/sample.py:58-91 line for line (the reference tree itself does
not travel to the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import argparse, json, os, sys, tempfile
import numpy as np
import torch
from models.models_pointcloud import GT_network_equiv          # src/inference_demo.py:5
from models.fit_SMPL import fit_smpl                            # src/inference_demo.py:10
import pointops_cuda, epn_grouping, epn_gathering, epn_zpconv   # src/models/pointops.py:5, vgtk/pc/sample.py:5-6, vgtk/zpconv/functional.py:17

repo = sys.argv[1]
sys.path.append(repo)      # only for the synthetic checkpoint / body / scan generators (a maintainer has real files instead)
from etch_b200 import smpl_model, synth

args = argparse.Namespace(gender="neutral", output_folder=tempfile.mkdtemp(), num_point=5000, scale_magnitude=10,
                          EPN_input_radius=0.4, EPN_layer_num=2)
args.cuda = torch.cuda.is_available()
args.device = torch.device("cuda" if args.cuda else "cpu")
with open(os.path.join(repo, "etch_b200", "data", "superset_smpl.json")) as f:
    args.markerset = json.load(f)
args.model_path = os.path.join(args.output_folder, "ckpt.pth")
torch.save(synth.make_state_dict(1), args.model_path)
args.smpl_model = smpl_model.synthetic_body(0)      # stands in for datafolder/body_models/smpl/... (licensed, not redistributable)

# ---- load_model (src/inference_demo.py:12-17)
model = GT_network_equiv(option=args).to(args.device)
model.load_state_dict(torch.load(args.model_path))
model.eval()
assert os.path.exists(os.path.join(args.output_folder, "EPN_model_setting_json"))

# ---- preprocess_scan + sample_points_from_mesh (:19-39) replaced by the committed real-scan cloud; centre as :25-28
points = synth.sample_real_scan(args.num_point, 5, rotate=False).astype(np.float64)
original_center = np.array([0.1, -0.2, 0.3])

# ---- predict_smpl (:41-66), verbatim
with torch.no_grad():
    points_tensor = torch.from_numpy(points).float().unsqueeze(0).to(args.device)
    PRED_ITEMS = ["confidence", "direction", "magnitude"]
    results, selected_indexs = model(points_tensor, pred_items=PRED_ITEMS, direction_mode="standard_vector")
    pred_part_labels = results["part_labels"]
    _, pred_part_labels = torch.max(pred_part_labels, -1)
    pred_confidences = results["confidences"]
    pred_directions = results["direction"]
    pred_magnitudes = results["magnitude"]
    pred_vectors = pred_directions * pred_magnitudes / args.scale_magnitude
    pred_inner_points = points_tensor - pred_vectors
    final_mesh_list, _, _, output_smpl_info = fit_smpl(args, pred_inner_points, pred_part_labels, pred_confidences, args.gender)
pred_smpl_mesh, smpl_info = final_mesh_list[0], output_smpl_info

# ---- main (:107-127)
pred_smpl_vertices = pred_smpl_mesh.vertices + original_center
final_smpl_mesh = pred_smpl_mesh.copy()
final_smpl_mesh.vertices = pred_smpl_vertices
final_smpl_mesh.export(os.path.join(args.output_folder, "scan_pred_smpl.obj"))
np.savez(os.path.join(args.output_folder, "scan_output_smpl_info.npz"), body_pose=smpl_info[0][0, :21, :], hand_pose=smpl_info[0][0, 21:23, :],
         betas=smpl_info[1][0], global_orient=smpl_info[2][0], transl=smpl_info[3][0], joints=smpl_info[4][0])
assert [d.shape for d in smpl_info] == [(1, 23, 3), (1, 10), (1, 3), (1, 3), (1, 45, 3)]
assert selected_indexs.shape == (1, 5000, 3) and results["part_labels"].shape == (1, 5000, 86)
assert np.asarray(pred_smpl_mesh.vertices).shape == (6890, 3) and np.isfinite(pred_smpl_vertices).all()
try:
    fit_smpl(args, pred_inner_points, pred_part_labels, pred_confidences, "robot")
except ValueError:
    pass
else:
    raise AssertionError("unknown gender must raise ValueError (fit_SMPL.py:98-99)")

# ---- the reference's wrapper calls over the B2 names: src/models/pointops.py:10-27 (furthestsampling) and :30-45 (knnquery)
xyz = points_tensor[0].contiguous()
offset = torch.cuda.IntTensor([5000]); new_offset = torch.cuda.IntTensor([1250])
n, b, n_max = xyz.shape[0], offset.shape[0], offset[0]
for i in range(1, b):
    n_max = max(offset[i] - offset[i - 1], n_max)
idx = torch.cuda.IntTensor(new_offset[b - 1].item()).zero_()
tmp = torch.cuda.FloatTensor(n).fill_(1e10)
pointops_cuda.furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx)
new_xyz = xyz[idx.long(), :]
m = new_xyz.shape[0]
kidx = torch.cuda.IntTensor(m, 16).zero_(); dist2 = torch.cuda.FloatTensor(m, 16).zero_()
pointops_cuda.knnquery_cuda(m, 16, xyz, new_xyz, offset, new_offset, kidx, dist2)
assert idx[0].item() == 0 and len(set(idx.tolist())) == 1250 and (kidx[:, 0] == idx).all() and (dist2[:, 1:] >= dist2[:, :-1]).all()
# vgtk/pc/sample.py:58-91: furthest_sample -> gather -> ball_query
pc = points_tensor.permute(0, 2, 1).contiguous()
sidx = epn_grouping.furthest_point_sampling(pc, 2500)
new_pc = epn_gathering.gather_points_forward(pc, sidx)
bidx = epn_grouping.ball_query(new_pc, pc, 0.08, 64)
assert new_pc.shape == (1, 3, 2500) and bidx.shape == (1, 2500, 64) and bidx.dtype == torch.int32
torch.cuda.synchronize()
print("dropin-gpu-ok")
'''


def test_inference_demo_call_sequence_through_dropin_names(cuda, tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "etch_b200", "dropin"), os.path.join(ROOT, "etch_b200", "ext")])
    script = tmp_path / "demo_sequence.py"
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dropin-gpu-ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


EVAL_SCRIPT = r'''
import argparse, json, os, sys, tempfile
import numpy as np
import torch
from models.models_pointcloud import GT_network_equiv           # src/eval.py:10
from models.fit_SMPL import fit_smpl                             # src/eval.py:14
repo = sys.argv[1]
sys.path.append(repo)
from etch_b200 import smpl_model, synth
from etch_b200.io import save_points_with_vector                 # drop-in for utils.GT_utils.save_points_with_vector (src/eval.py:12)

args = argparse.Namespace(output_folder=tempfile.mkdtemp(), scale_magnitude=10, EPN_input_radius=0.4, EPN_layer_num=2, batch_size=3)
assert torch.cuda.is_available()                                  # src/eval.py:294
args.device = torch.device("cuda")
args.markerset = json.load(open(os.path.join(repo, "etch_b200", "data", "superset_smpl.json")))
args.smpl_model = smpl_model.synthetic_body(0)
model = GT_network_equiv(option=args).to(args.device)             # src/eval.py:308-310
model.load_state_dict(synth.make_state_dict(1))
model.eval()
B, N = 3, 5000
batch = {"hitpts": torch.from_numpy(synth.sample_real_scans(B, N, 11)), "gender": ["male", "male", "female"], "id": ["a", "b", "c"]}
PRED_ITEMS = ["confidence", "direction", "magnitude"]
with torch.inference_mode():                                      # src/eval.py:88-116, verbatim
    hitpts = batch["hitpts"].to(args.device)
    results, selected_indexs = model(hitpts, PRED_ITEMS, direction_mode="standard_vector")
    pred_part_labels = results["part_labels"]
    _, pred_part_labels = torch.max(pred_part_labels, -1)
    pred_confidences = results["confidences"]
    pred_directions = results["direction"]
    pred_magnitudes = results["magnitude"]
    pred_vectors = pred_directions * pred_magnitudes / args.scale_magnitude
    hitpts_k = torch.gather(hitpts, 1, selected_indexs)            # :119
    for j in range(B):                                            # :126-149 (the writer that is ours)
        os.makedirs(os.path.join(args.output_folder, batch["id"][j]), exist_ok=True)
        save_points_with_vector(hitpts_k[j].clone().detach().cpu().numpy(), pred_vectors[j].clone().detach().cpu().numpy(),
                                os.path.join(args.output_folder, batch["id"][j], "hitpts_pred_vectors.ply"))
with torch.inference_mode():
    pred_inner_points = hitpts_k - pred_vectors                   # :183
gender_list = batch["gender"]                                     # :186-209, different genders -> per-sample calls
final_mesh_list, pred_markers_position, valid_mask, output_smpl_info = [], [], [], [[], [], [], [], []]
for l, gender in enumerate(gender_list):
    m_, p_, v_, info_ = fit_smpl(args, pred_inner_points[l].unsqueeze(0), pred_part_labels[l].unsqueeze(0), pred_confidences[l].unsqueeze(0), gender)
    final_mesh_list.append(m_[0]); pred_markers_position.append(p_[0]); valid_mask.append(v_[0])
    for i in range(len(info_)):
        output_smpl_info[i].append(info_[i][0])
pred_markers_position = torch.stack(pred_markers_position, dim=0)
valid_mask = torch.stack(valid_mask, dim=0)
output_smpl_info = [np.stack(info, axis=0) for info in output_smpl_info]
# the same batch in ONE call (same gender branch, :188-189) must give the same meshes as the per-sample calls
m_all, p_all, v_all, info_all = fit_smpl(args, pred_inner_points, pred_part_labels, pred_confidences, "male")
for j in range(B):
    v2v = np.mean(np.linalg.norm(np.asarray(m_all[j].vertices) - np.asarray(final_mesh_list[j].vertices), axis=1))    # :235-237
    assert v2v < 1e-6, v2v
    final_mesh_list[j].export(os.path.join(args.output_folder, batch["id"][j], "forwarded_smpl_mesh_on_pred.obj"))     # :230
    np.savez(os.path.join(args.output_folder, batch["id"][j], "output_smpl_info.npz"), body_pose=output_smpl_info[0][j][:21, :],
             hand_pose=output_smpl_info[0][j][21:23, :], betas=output_smpl_info[1][j], global_orient=output_smpl_info[2][j],
             transl=output_smpl_info[3][j], joints=output_smpl_info[4][j])                                             # :241-247
assert pred_markers_position.shape == (B, 86, 3) and valid_mask.shape == (B, 86) and valid_mask.dtype == torch.bool
assert int(valid_mask[0].sum()) > 40
ply = open(os.path.join(args.output_folder, "a", "hitpts_pred_vectors.ply")).read().split("\n")
assert ply[2] == "element vertex %d" % (2 * N) and len(ply) == 13 + 3 * N + 1
print("eval-sequence-ok")
'''


def test_eval_call_sequence_through_dropin_names(cuda, tmp_path):
    """src/eval.py:88-116,183-209,230-247 (batch forward, per-sample and whole-batch fit, the writers that are ours) in a fresh
    interpreter with only the drop-in directories on PYTHONPATH."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "etch_b200", "dropin"), os.path.join(ROOT, "etch_b200", "ext")])
    script = tmp_path / "eval_sequence.py"
    script.write_text(EVAL_SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "eval-sequence-ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_tensors_on_another_device_than_the_current_one(cuda):
    """launches follow the tensors' device (and mixed devices raise) instead of the thread's current device."""
    import torch
    from etch_b200 import _lib as L
    from etch_b200.ext import epn_grouping
    x = torch.randn(1, 3, 256, device=cuda)
    a = L.ptr(x)
    assert a.device_index == 0
    if torch.cuda.device_count() > 1:
        y = torch.randn(1, 3, 256, device="cuda:1")
        with pytest.raises(RuntimeError):
            L.call("ball_query_bcn", L.ptr(x), L.ptr(y), 1, 256, 256, L.f32(0.1), 4, L.ptr(torch.empty(1, 256, 4, dtype=torch.int32, device=cuda)))
        got = epn_grouping.furthest_point_sampling(y, 64)          # current device is cuda:0, tensors on cuda:1
        ref = epn_grouping.furthest_point_sampling(y.to(cuda), 64)
        assert got.device.index == 1 and (got.cpu() == ref.cpu()).all()
    idx = epn_grouping.furthest_point_sampling(x, 64)
    assert idx.shape == (1, 64)


def test_fit_smpl_loads_the_body_model_from_the_reference_paths(cuda, tmp_path, monkeypatch):
    """fit_smpl without args.smpl_model reads datafolder/body_models/smpl/<gender>/SMPL_*.pkl relative to the working directory,
    exactly where the reference looks (fit_SMPL.py:92-99); here the pickle is a synthetic SMPL-shaped model written in the SMPL
    pickle layout (v_template, shapedirs, posedirs [V,3,207], J_regressor, kintree_table, weights, f)."""
    import pickle
    import types
    import numpy as np
    import torch
    from etch_b200 import smpl_model
    from etch_b200.models import fit_SMPL
    body = smpl_model.synthetic_body(0)
    V = body["v_template"].shape[0]
    kin = np.stack([np.where(body["parents"] < 0, 2 ** 32 - 1, body["parents"]).astype(np.int64), np.arange(24)], 0)
    pkl = {"v_template": body["v_template"].astype(np.float64), "shapedirs": body["shapedirs"].astype(np.float64),
           "posedirs": body["posedirs"].T.reshape(V, 3, 207).astype(np.float64), "J_regressor": body["J_regressor"].astype(np.float64),
           "kintree_table": kin, "weights": body["lbs_weights"].astype(np.float64), "f": body["faces"].astype(np.uint32)}
    for rel in ("neutral/SMPL_NEUTRAL_10pc_rmchumpy.pkl", "male/SMPL_MALE_10pc.pkl"):
        path = tmp_path / "datafolder" / "body_models" / "smpl" / rel
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "wb") as fh:
            pickle.dump(pkl, fh)
    monkeypatch.chdir(tmp_path)
    ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
    rng = np.random.default_rng(0)
    N = 2000
    inner = torch.from_numpy(body["v_template"][rng.integers(0, V, N)][None].astype(np.float32)).to(cuda)
    labels = torch.from_numpy(rng.integers(0, 86, (1, N))).to(cuda)
    conf = torch.from_numpy((0.4 + 0.5 * rng.random((1, N, 1))).astype(np.float32)).to(cuda)
    a_file = types.SimpleNamespace(markerset=ms, device="cuda:0")
    a_dict = types.SimpleNamespace(markerset=ms, device="cuda:0", smpl_model=body)
    for gender in ("neutral", "male"):
        m1, p1, v1, i1 = fit_SMPL.fit_smpl(a_file, inner, labels, conf, gender)
        m2, p2, v2, i2 = fit_SMPL.fit_smpl(a_dict, inner, labels, conf, gender)
        assert np.abs(np.asarray(m1[0].vertices) - np.asarray(m2[0].vertices)).max() < 1e-6
        assert (np.asarray(m1[0].faces) == body["faces"]).all()
    with pytest.raises(FileNotFoundError):
        fit_SMPL.fit_smpl(a_file, inner, labels, conf, "female")      # no such file under the reference's path
