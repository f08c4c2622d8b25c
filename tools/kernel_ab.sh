#!/bin/bash
# A/B of an experimental CUDA source against the shipping one of the same name: builds a scratch copy of the library in which
# etch_b200/csrc/<name>.cu is replaced by <variant.cu> and prints the per-kernel CUDA-event times of one eager step at the bench shape.
#   bash tools/kernel_ab.sh pt_tc /tmp/pt_tc_variant.cu pt_attention_tc [more_variants.cu ...]   (the kernel key may be a comma-separated list)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
NAME=$1; VARIANT=$2; KEY=$3; shift 3
python -m etch_b200.build > /dev/null
for src in etch_b200/csrc/$NAME.cu "$VARIANT" "$@"; do
  tag=$(basename ${src%.cu})_$(echo $src | md5sum | cut -c1-6)
  o=/tmp/ab_$tag.o
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I include -I etch_b200/csrc -c $src -o $o 2>/dev/null
  objs=""
  for f in etch_b200/csrc/*.cu; do
    if [ "$(basename $f)" = "$NAME.cu" ]; then objs="$objs $o"; else objs="$objs etch_b200/build/$(basename ${f%.cu}).o"; fi
  done
  nvcc -shared -o /tmp/libetch_ab_$tag.so $objs -lcudart
  ETCH_B200_LIB=/tmp/libetch_ab_$tag.so python - <<PY
import sys
sys.path.insert(0, "$ROOT")
import torch
import bench
from etch_b200 import _lib as L, synth
dev = torch.device("cuda:0")
pipe = bench.Pipeline(dev, use_graph=False)
pts = torch.from_numpy(synth.sample_real_scans(8, 5000, 50)).to(dev)
ref = None
best = {}
for rep in range(4):
    L.start_profile(); out = pipe.eager(pts); prof = L.stop_profile()
    for k, (c, t) in prof.items():
        best[k] = min(best.get(k, 1e9), t)
print("$src: %s ms; step sum %.2f ms; checksum %.6f" % (" ".join("%s %.3f" % (k, best[k]) for k in "$KEY".split(",")), sum(best.values()), out["vertices"][torch.isfinite(out["vertices"]).all(-1).all(-1)].double().abs().mean().item()), flush=True)
PY
done
