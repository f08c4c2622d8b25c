#!/bin/bash
# A/B of the control-warp poll interval of the inter conv: builds scratch copies of the library with -DETCH_V3_SLEEP_NS=<ns> and times
# the three launches at the bench shape.   bash tools/v3_sleep_ab.sh 64 128 512
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
for ns in "$@"; do
  out=/tmp/libetch_sleep_$ns.so
  objs=""
  for f in etch_b200/csrc/*.cu; do
    o=/tmp/sleep_${ns}_$(basename ${f%.cu}).o
    if [ "$(basename $f)" = "so3conv_v3.cu" ]; then
      nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DETCH_V3_SLEEP_NS=$ns -I include -c $f -o $o
    else
      o=etch_b200/build/$(basename ${f%.cu}).o
    fi
    objs="$objs $o"
  done
  nvcc -shared -o $out $objs -lcudart
  ETCH_B200_LIB=$out python - <<PY
import os, sys
sys.path.insert(0, "$ROOT")
import torch
from etch_b200 import _lib as L, synth
from etch_b200.models import encoder
dev = torch.device("cuda:0")
plan = encoder.EncoderPlan(synth.make_state_dict(1), dev)
pts = torch.from_numpy(synth.sample_real_scans(8, 5000, 50)).permute(0, 2, 1).contiguous().to(dev)
best = 1e9
for rep in range(4):
    L.start_profile(); encoder.run_encoder(plan, pts); prof = L.stop_profile()
    best = min(best, prof["so3_inter_conv_v3"][1])
print("sleep $ns ns: inter_conv_v3 3 launches %.3f ms" % best, flush=True)
PY
done
