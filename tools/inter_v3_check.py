"""A/B check of the InterSO3Conv variants on the GPU: v3 (one point per tile) against v2 (2-point slab kernel) and the fp32
CUDA-core kernel, layer by layer on a seeded scan; then per-layer timings at the bench shape.

  python tools/inter_v3_check.py parity [B N sm_budget]
  python tools/inter_v3_check.py time   [B N]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from etch_b200 import _lib as L, synth  # noqa: E402
from etch_b200.models import encoder  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
dev = torch.device("cuda:0")
sd = synth.make_state_dict(1)
plan = encoder.EncoderPlan(sd, dev)


def run(variant, pts, use_tc=True):
    encoder.INTER_VARIANT = variant
    encoder.USE_TC = use_tc
    tr = []
    encoder.run_encoder(plan, pts, tr)
    torch.cuda.synchronize()
    return tr


if mode == "parity":
    B, N, budget = (int(x) for x in (sys.argv[2:5] + ["2", "1024", "8"][len(sys.argv) - 2:]))
    L.lib().etch_set_sm_budget(budget)
    pts = torch.from_numpy(synth.sample_scans(B, N, 7)).permute(0, 2, 1).contiguous().to(dev)
    ref = run("v2", pts, use_tc=False)
    v2 = run("v2", pts)
    v3 = run("v3", pts)
    ok = True
    for li in range(len(ref)):
        for key in ("inter_z", "out"):
            r = ref[li][key]
            scale = r.abs().max().item()
            e2 = (v2[li][key] - r).abs().max().item() / scale
            e3 = (v3[li][key] - r).abs().max().item() / scale
            print("layer %d %-8s scale %.3e  rel.err v2 %.2e  v3 %.2e" % (li, key, scale, e2, e3), flush=True)
            ok = ok and e3 < 1e-4
    print("PARITY", "OK" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
else:
    B, N = (int(x) for x in (sys.argv[2:4] + ["8", "5000"][len(sys.argv) - 2:]))
    pts = torch.from_numpy(synth.sample_scans(B, N, 50)).permute(0, 2, 1).contiguous().to(dev)
    for variant in ("v2", "v3"):
        run(variant, pts)
        L.start_profile()
        run(variant, pts)
        prof = L.stop_profile()
        print(variant, {k: (c, round(t, 3)) for k, (c, t) in prof.items() if "inter" in k}, flush=True)
        # per-launch times of the inter conv
        for k, v in prof.items():
            pass
