#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list + full captures of the heavy kernels.  Outputs under gpurun_out/<tag>/.
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 600 $OUT/bench.err
python -c "
import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'], d['cpu_baseline'])
for k,v in d['kernels_ms'].items(): print(k,v)
"
if [ -n "$NCU_LIST" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/ncu_list.log 2>&1
fi
if [ -n "$NCU_FULL" ]; then
  for k in $NCU_FULL; do
    timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c ${NCU_COUNT:-1} -f -o $OUT/full_$k python tools/profile_step.py > $OUT/ncu_full_$k.log 2>&1
    ncu -i $OUT/full_$k.ncu-rep --page raw --csv > $OUT/full_$k.raw.csv 2>/dev/null
  done
fi
ls -la $OUT
