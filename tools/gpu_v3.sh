#!/bin/bash
# one gpurun call: probes + A/B parity/timing of the inter-conv variants (each step under its own timeout)
OUT=gpurun_out/${1:-v3}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 180 python tools/inter_v3_check.py parity 2 1024 8 > $OUT/parity_small.log 2>&1; echo "parity small exit $?"; tail -12 $OUT/parity_small.log
timeout 180 python tools/inter_v3_check.py parity 1 1531 148 > $OUT/parity_odd.log 2>&1; echo "parity odd exit $?"; tail -12 $OUT/parity_odd.log
timeout 180 python tools/inter_v3_check.py time 8 5000 > $OUT/time.log 2>&1; echo "time exit $?"; tail -4 $OUT/time.log
