#!/bin/bash
# A/B of an experimental InterSO3Conv source against the shipping one: builds a scratch copy of the library in which
# etch_b200/csrc/so3conv_v3.cu is replaced by <variant.cu>, checks it against the fp32 CUDA-core kernel and times the three launches at
# the bench shape.   bash tools/v3_ab.sh tools/experiments/so3conv_v3_w3.cu
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
python -m etch_b200.build > /dev/null
for src in etch_b200/csrc/so3conv_v3.cu "$@"; do
  tag=$(basename ${src%.cu})
  o=/tmp/ab_$tag.o
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I include -I etch_b200/csrc -c $src -o $o 2>/dev/null
  objs=""
  for f in etch_b200/csrc/*.cu; do
    if [ "$(basename $f)" = "so3conv_v3.cu" ]; then objs="$objs $o"; else objs="$objs etch_b200/build/$(basename ${f%.cu}).o"; fi
  done
  nvcc -shared -o /tmp/libetch_ab_$tag.so $objs -lcudart
  ETCH_B200_LIB=/tmp/libetch_ab_$tag.so python - <<PY
import sys
sys.path.insert(0, "$ROOT")
import torch
from etch_b200 import _lib as L, synth
from etch_b200.models import encoder
dev = torch.device("cuda:0")
plan = encoder.EncoderPlan(synth.make_state_dict(1), dev)
def run(variant, pts, use_tc=True):
    encoder.INTER_VARIANT = variant; encoder.USE_TC = use_tc
    tr = []; encoder.run_encoder(plan, pts, tr); torch.cuda.synchronize(); return tr
small = torch.from_numpy(synth.sample_real_scans(2, 1024, 7)).permute(0, 2, 1).contiguous().to(dev)
ref = run("v2", small, use_tc=False); got = run("v3", small)
err = max(((got[l]["inter_z"] - ref[l]["inter_z"]).abs().max() / ref[l]["inter_z"].abs().max()).item() for l in range(len(ref)))
pts = torch.from_numpy(synth.sample_real_scans(8, 5000, 50)).permute(0, 2, 1).contiguous().to(dev)
best = [1e9] * 3
for rep in range(5):
    L.start_profile(); run("v3", pts); L.stop_profile() if False else None
    prof = L._profile
    torch.cuda.synchronize()
    evs = prof["so3_inter_conv_v3"]
    t = [a.elapsed_time(b) for a, b in evs]
    best = [min(x, y) for x, y in zip(best, t)]
    L._profile = None
print("$tag: parity rel.err %.2e; launches %s ms, total %.3f ms" % (err, ["%.3f" % x for x in best], sum(best)), flush=True)
PY
done
