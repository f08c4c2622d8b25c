#!/bin/bash
OUT=gpurun_out/r02e
mkdir -p $OUT
timeout 900 python tools/ref_kernel_bench.py $OUT/ref_kernel_bench.md > $OUT/ref_kernel_bench.log 2>&1; echo "refbench exit $?"; tail -25 $OUT/ref_kernel_bench.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; grep -v "^=========     " $OUT/racecheck.log | tail -8
for c in 2 3; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_config$c.json 2> $OUT/bench_config$c.err; python -c "
import json;d=json.loads(open('$OUT/bench_config$c.json').read().strip().splitlines()[-1]); print('config $c', round(d['value'],1), d['ms_per_step'], d['e2e']['value'], d['inflight_check'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1; echo "ncu list exit $?"
ls -la $OUT
