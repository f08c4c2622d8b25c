#!/bin/bash
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 900 python tools/v3_variants.py 0 1 3 7 > $OUT/v3_variants.log 2>&1; cat $OUT/v3_variants.log
timeout 900 python -m pytest tests/test_mesh_gpu.py tests/test_ref_kernels_gpu.py tests/test_index_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_run.py > $OUT/synccheck.log 2>&1; echo "synccheck exit $?"; grep -v "^=========     " $OUT/synccheck.log | tail -5
SAN_POINTS=384 SAN_EXTRA=0 timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_run.py > $OUT/initcheck.log 2>&1; echo "initcheck exit $?"; grep -v "^=========     " $OUT/initcheck.log | tail -5
