#!/bin/bash
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|FAILED|Error" $OUT/pytest.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:inter_conv_v3 -c 3 -f -o $OUT/full_v3 python tools/profile_step.py > $OUT/ncu_full_v3.log 2>&1
ls -la $OUT
