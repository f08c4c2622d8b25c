"""Generates etch_b200/data/synth_calibration.npz: data-calibrated BatchNorm statistics (+ a confidence offset) for the seeded
synthetic checkpoint etch_b200.synth.make_state_dict(1).

Why: a purely random PointTransformer collapses -- with BatchNorm running statistics that do not describe the activations,
almost all points get the same 3 marker labels and confidences of ~0.01 (conf**20 underflows, markers turn NaN), so the
marker extraction and the LM fit downstream would be tested and benchmarked on a degenerate problem (3 of 86 markers valid).
A trained checkpoint's running statistics ARE its activation statistics; this script gives the synthetic checkpoint the same
property by one forward pass of the UNMODIFIED reference network (src/models/models_pointcloud.py, imported through
tools/ref_shim.py) on twelve clouds of the in-tree 4D-Dress sample, in which every BatchNorm1d first sets its running statistics to
the statistics of its own input (forward pre-hook, eval mode), and reads the running statistics back.  `confi.2.bias` is shifted so that confidences sit around 0.5 and the last magnitude layer is rescaled to
0.3 +- 0.1 (3 cm tightness), like a trained network's.  Only the statistics travel (small); the weights stay the seeded ones.  Re-run: python tools/gen_calibration.py
(then tools/gen_golden.py, whose network golden depends on the checkpoint).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()
from etch_b200 import synth  # noqa: E402
from models.models_pointcloud import GT_network_equiv  # noqa: E402  (the reference's)

markerset = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
torch.set_num_threads(os.cpu_count() or 1)
net = GT_network_equiv(ref_shim.make_option(markerset)).eval()
sd = synth.make_state_dict(1, calibrated=False)
net.load_state_dict(sd, strict=True)
bns = {n: m for n, m in net.named_modules() if isinstance(m, torch.nn.BatchNorm1d)}


def _calibrate(mod, inp):
    # eval-mode forward pre-hook: this layer's running statistics := the statistics of the activations it is about to see, so
    # that the calibrated eval-mode network is exactly self-consistent (every BatchNorm output is zero-mean / unit-variance on
    # the calibration clouds), also at the deep levels where a training-mode pass (unbiased variance of 38 points) is not
    x = inp[0]
    dims = [0] if x.dim() == 2 else [0, 2]
    mod.running_mean.copy_(x.mean(dims))
    mod.running_var.copy_(x.var(dims, unbiased=False))


hooks = [m.register_forward_pre_hook(_calibrate) for m in bns.values()]
# 12 clouds (6 of the real scan, 6 of the body under it, each with its own rotation / subsample / jitter): at the deep levels
# (19 points per scan) the variance ACROSS scans is what a training run would have accumulated, not the within-scan variance
NCAL = 12
from models.pointnet2_utils import PointFeatPropagation  # noqa: E402  (the reference's)


def inv_features(p):
    """the encoder half of the reference forward (models_pointcloud.py:154-184), one scan at a time (its intermediates are ~4 GB
    per 5000-point scan on the CPU); it has no BatchNorm, so it needs no calibration"""
    feats = []
    for b in range(p.shape[0]):
        q = p[b:b + 1]
        r, _ = net.encode(q)
        S = r.xyz.shape[-1]
        equiv = r.feats.permute(0, 1, 3, 2).reshape(1, -1, S)
        pe = PointFeatPropagation(xyz1=q.permute(0, 2, 1), xyz2=r.xyz, points2=equiv).reshape(1, q.shape[1], -1, 60)
        feats.append(pe.mean(-1))
    return torch.cat(feats, 0)


def heads(p, inv):
    labels, conf = net.decode_confidence(inv, p)
    return {"part_labels": labels, "confidences": conf, "magnitude": net.decode_magnitude(inv, p)}


pts = torch.from_numpy(synth.sample_real_scans(NCAL, 5000, 0))
with torch.no_grad():
    inv = inv_features(pts)
    heads(pts, inv)
for h in hooks:
    h.remove()
out = {}
for n, m in bns.items():
    out[n + ".running_mean"] = m.running_mean.detach().numpy().astype(np.float32)
    out[n + ".running_var"] = m.running_var.detach().numpy().astype(np.float32)
with torch.no_grad():
    res = heads(pts, inv)
conf = res["confidences"]
shift = float(0.5 - conf.mean())
cur = net.state_dict()
bias = cur["confidence_encoder.confi.2.bias"].numpy() + shift
out["confidence_encoder.confi.2.bias"] = bias.astype(np.float32)
# magnitudes of a trained network are tightness x 10 in metres: ~0.3 +- 0.1 (3 cm of cloth); rescale the last linear layer
mag = res["magnitude"]
gain = 0.1 / float(mag.std())
w3 = cur["magnitude_encoder.final_layer.3.weight"].numpy() * gain
b3 = (cur["magnitude_encoder.final_layer.3.bias"].numpy() - float(mag.mean())) * gain + 0.3
out["magnitude_encoder.final_layer.3.weight"] = w3.astype(np.float32)
out["magnitude_encoder.final_layer.3.bias"] = b3.astype(np.float32)
net.load_state_dict({**cur, **{k: torch.from_numpy(out[k]) for k in ("confidence_encoder.confi.2.bias",
                                                                      "magnitude_encoder.final_layer.3.weight",
                                                                      "magnitude_encoder.final_layer.3.bias")}})
with torch.no_grad():
    res = heads(pts, inv)
held = torch.from_numpy(synth.sample_real_scans(2, 5000, 7))     # held-out clouds (the ones tests/golden uses)
with torch.no_grad():
    hinv = inv_features(held)
    rh = heads(held, hinv)
    r1 = heads(held[:1], hinv[:1])
print("held-out, alone (B=1): labels used", int(torch.unique(r1["part_labels"].argmax(-1)).numel()))
print("held-out: labels used", [int(torch.unique(l).numel()) for l in rh["part_labels"].argmax(-1)], "conf %.3f .. %.3f" % (
    rh["confidences"].min().item(), rh["confidences"].max().item()), "mag %.3f .. %.3f" % (rh["magnitude"].min().item(), rh["magnitude"].max().item()))
lab = res["part_labels"].argmax(-1)
print("labels used per scan:", [int(torch.unique(l).numel()) for l in lab], "| confidence range %.3f .. %.3f" % (
    res["confidences"].min().item(), res["confidences"].max().item()), "| magnitude range %.3f .. %.3f" % (
    res["magnitude"].min().item(), res["magnitude"].max().item()))
path = os.path.join(ROOT, "etch_b200", "data", "synth_calibration.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
