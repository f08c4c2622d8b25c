"""Quick direction-head check (run under `timeout`): network forward on a small and a bench-size batch, compared with the
fp32 CUDA-core direction head (ETCH_B200_NO_TC path of heads.py), plus the kernel time."""
import os, sys, json, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from etch_b200 import _lib as L, synth
from etch_b200.models import heads
from etch_b200.models.models_pointcloud import GT_network_equiv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
net = GT_network_equiv(types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=ms))
net.load_state_dict(synth.make_state_dict(1))
dev = torch.device("cuda:0")
net = net.to(dev).eval()
for B, N in ((2, 1024), (1, 333), (8, 5000)):
    pts = torch.from_numpy(synth.sample_scans(B, N, 5)).to(dev)
    for _ in range(3):      # the last repetition is the warmed one
        L.start_profile()
        o, _ = net(pts, ["direction"])
        prof = L.stop_profile()
    torch.cuda.synchronize()
    d = o["direction"]
    print(B, N, "finite", bool(torch.isfinite(d).all()), "unit", float((d.norm(dim=-1) - 1).abs().max()),
          {k: round(v[1], 3) for k, v in prof.items() if "direction" in k}, flush=True)
    heads.USE_TC = False
    o2, _ = net(pts, ["direction"])
    heads.USE_TC = True
    torch.cuda.synchronize()
    print("   vs fp32 head: median angle err", float((d - o2["direction"]).norm(dim=-1).median()), flush=True)
