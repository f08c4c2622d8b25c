#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds oracle/_ref/libetch_ref_kernels.so from the reference's OWN CUDA sources, compiled unmodified
# from where they lie under $ETCH_REFERENCE (default /root/reference) -- nothing is copied into the repo.  The two ATen
# headers they include are replaced by the 40-line stand-ins in oracle/refshim (raw pointer + sizes), so torch is not linked.
#   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu     ball_query_cuda (:67-113,471-490), furthest_point_sampling_cuda (:351-466,625-739)
#   external/vgtk/vgtk/cuda/gathering_cuda_kernel.cu    gather_points_forward_cuda (:42-68,102-128)
#   external/pointops/src/knnquery/knnquery_cuda_kernel.cu     knnquery_cuda_launcher (:65-116)
#   external/pointops/src/sampling/sampling_cuda_kernel.cu     furthestsampling_cuda_launcher (:15-171)
# Flags follow the reference build (nvcc -O2: external/vgtk/vgtk/setup.py:13-30, external/pointops/setup.py:16-31) plus the
# sm_100a target.  Output is git-ignored but travels to the GPU box with gpurun.  Exits 0 without doing anything when the
# reference tree is absent (the GPU box uses the prebuilt file).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${ETCH_REFERENCE:-/root/reference}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/_ref"
if [ ! -d "$REF/external/pointops/src" ]; then
  echo "build_ref.sh: $REF not present -- keeping any prebuilt $OUT/libetch_ref_kernels.so"; exit 0
fi
mkdir -p "$OUT"
SRCS=(
  "$REF/external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu"
  "$REF/external/vgtk/vgtk/cuda/gathering_cuda_kernel.cu"
  "$REF/external/pointops/src/knnquery/knnquery_cuda_kernel.cu"
  "$REF/external/pointops/src/sampling/sampling_cuda_kernel.cu"
  "$HERE/ref_launch.cu"
)
STAMP="$OUT/.stamp"
SIG="$(cat "${SRCS[@]}" "$HERE"/refshim/ATen/ATen.h | sha1sum | cut -d' ' -f1)"
if [ -f "$OUT/libetch_ref_kernels.so" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$SIG" ]; then exit 0; fi
OBJS=()
for s in "${SRCS[@]}"; do
  o="$OUT/$(basename "${s%.cu}").o"
  "$NVCC" -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -w -I "$HERE/refshim" -c "$s" -o "$o"
  OBJS+=("$o")
done
"$NVCC" -shared -o "$OUT/libetch_ref_kernels.so" "${OBJS[@]}" -lcudart
rm -f "${OBJS[@]}"
echo "$SIG" > "$STAMP"
echo "built $OUT/libetch_ref_kernels.so"
