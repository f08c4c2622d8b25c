"""Latency microbenchmarks (SM cycles) of the tcgen05 / TMA primitives and a per-phase breakdown of the LM solver."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from etch_b200 import _lib as L, smpl_model  # noqa: E402
from etch_b200.models import fit_SMPL as F  # noqa: E402

dev = torch.device("cuda:0")
src = torch.randn(64 * 1024, device=dev)
out = torch.zeros(64, dtype=torch.int64, device=dev)
L.call("umma_latency", L.ptr(src), L.ptr(out))
torch.cuda.synchronize()
o = out.cpu().tolist()
names = ["bulk8K", "bulk16K", "bulk32K", "bulk64K", "mma24_N32_issue", "mma24_N32_total", "mma24_N64_issue", "mma24_N64_total", "fence+sync"]
for rep in range(3):
    print("rep", rep, {n: o[rep * 9 + i] for i, n in enumerate(names)})

ms = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "etch_b200", "data", "superset_smpl.json")))
args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
T = F.body_tables(args, "neutral", dev)
B = 8
g = torch.Generator().manual_seed(0)
markers = (0.3 * torch.randn(B, 86, 3, generator=g)).to(dev)
valid = torch.ones(B, 86, dtype=torch.uint8, device=dev)
params = torch.empty(B, 85, device=dev); iters = torch.empty(B, 2, dtype=torch.int32, device=dev); errs = torch.empty(B, 2, device=dev)
prof = torch.zeros(B, 2, 6, dtype=torch.int64, device=dev)
for _ in range(2):
    L.call("lm_fit_profile", L.ptr(markers), L.ptr(valid), L.ptr(T.Tm), L.ptr(T.Sm), L.ptr(T.Pm), L.ptr(T.Wm), L.ptr(T.Jt), L.ptr(T.Js),
           L.ptr(T.parents), L.ptr(T.ancmask), B, T.M, 30, 50, L.f32(0.5), L.f32(0.2), L.f32(0.01), L.f32(1e-3), L.ptr(params), L.ptr(iters),
           L.ptr(errs), L.ptr(prof))
torch.cuda.synchronize()
p = prof.cpu().numpy().astype(np.float64)
it = iters.cpu().numpy()
print("LM iters", it[0].tolist(), "cycles per iteration [eval, jacobian, solve, JtJ, factor, backsub] stage0:", (p[0, 0] / it[0, 0]).round().tolist(),
      "stage1:", (p[0, 1] / it[0, 1]).round().tolist())
