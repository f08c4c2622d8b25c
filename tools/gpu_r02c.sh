#!/bin/bash
OUT=gpurun_out/r02c
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_ref_kernels_gpu.py tests/test_index_gpu.py tests/test_parity_full_gpu.py tests/test_dropin_gpu.py -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|FAILED|Error" $OUT/pytest.log | tail -20
for cfg in "148 5" "148 1" "148 2" "148 3" "140 5" "132 5" "148 8"; do
  set -- $cfg
  ETCH_SM_BUDGET=$1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --in-flight $2 > $OUT/bench_$1_$2.json 2>$OUT/bench_$1_$2.err
  python -c "
import json;d=json.loads(open('$OUT/bench_$1_$2.json').read().strip().splitlines()[-1])
print('budget $1 in_flight $2:', round(d['value'],1), 'scans/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1), 'sum kernels', round(sum(v['ms'] for v in d['kernels_ms'].values()),2))"
done
