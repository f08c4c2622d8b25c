#!/bin/bash
OUT=gpurun_out/r02i
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 exit $?"
python -c "
import json;d=json.loads(open('$OUT/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']); print(d['inflight_check']); print(d['parity']); print(d['cpu_baseline'])"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 exit $?"
python -c "
import json;d=json.loads(open('$OUT/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['fp32_simt_frac']); print(d['parity']); print(d['cpu_baseline'])"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/ref_n1.json 2> $OUT/ref_n1.err; cut -c1-300 $OUT/ref_n1.json
