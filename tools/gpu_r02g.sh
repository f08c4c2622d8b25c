#!/bin/bash
OUT=gpurun_out/r02g
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
for cfg in "148 5" "148 8" "140 8"; do
  set -- $cfg
  ETCH_SM_BUDGET=$1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight $2 > $OUT/bench_$1_$2.json 2>$OUT/bench_$1_$2.err
  python -c "
import json;d=json.loads(open('$OUT/bench_$1_$2.json').read().strip().splitlines()[-1])
print('budget $1 in_flight $2:', round(d['value'],1), 'scans/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1))"
done
timeout 600 python tools/bench_mixed.py --scans 48 --passes 2 > $OUT/bench_mixed.json 2> $OUT/bench_mixed.err; tail -c 700 $OUT/bench_mixed.json
