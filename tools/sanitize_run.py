"""One small eager pass of the whole hot path (every product kernel once, no CUDA graph) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers the index kernels (both FPS variants incl. the cluster one, ball query, kNN grid + brute force), the encoder
(TMA / mbarrier / TMEM kernels), the heads and the fit.  Sizes are small because the tools slow kernels down 10-100x."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from etch_b200 import smpl_model, synth  # noqa: E402
from etch_b200.ext import epn_grouping, pointops_cuda  # noqa: E402
from etch_b200.models.models_pointcloud import GT_network_equiv  # noqa: E402
from etch_b200.runtime import ScanFitter  # noqa: E402

N = int(os.environ.get("SAN_POINTS", "640"))
dev = torch.device("cuda:0")
ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
opt = types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=ms)
net = GT_network_equiv(opt)
net.load_state_dict(synth.make_state_dict(1))
net = net.to(dev).eval()
args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
fit = ScanFitter(net, args, use_graph=False)(torch.from_numpy(synth.sample_real_scans(2, N, 3)).to(dev))
torch.cuda.synchronize()
print("pipeline ok, finite scans:", int(torch.isfinite(fit["vertices"]).all(-1).all(-1).sum()))
if os.environ.get("SAN_EXTRA", "1") == "1":
    # the cluster FPS path (> 8192 points) and the brute-force kNN binding
    big = torch.from_numpy(np.ascontiguousarray(synth.sample_real_scans(1, 9000, 4).transpose(0, 2, 1))).to(dev)
    epn_grouping.furthest_point_sampling(big, 600)
    xyz = torch.from_numpy(synth.sample_real_scans(1, 1500, 5)[0]).to(dev)
    off = torch.tensor([1500], dtype=torch.int32, device=dev)
    idx = torch.zeros(1500, 16, dtype=torch.int32, device=dev)
    d2 = torch.zeros(1500, 16, dtype=torch.float32, device=dev)
    pointops_cuda.knnquery_cuda(1500, 16, xyz, xyz, off, off, idx, d2)
    torch.cuda.synchronize()
    print("extra kernels ok")
