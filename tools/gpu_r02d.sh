#!/bin/bash
# 2-GPU call: the NCCL scatter/gather e2e path of bench.py, plus (on GPU 0) the mesh tests
OUT=gpurun_out/r02d
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 exit $?"
tail -c 1200 $OUT/bench_n2.err
python -c "
import json;d=json.loads(open('$OUT/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']); print(d['inflight_check'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $OUT/ref_n2.json 2> $OUT/ref_n2.err; echo "ref n2 exit $?"; cut -c1-400 $OUT/ref_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_mesh_gpu.py tests/test_dropin_gpu.py -m gpu -q > $OUT/pytest_mesh.log 2>&1; tail -3 $OUT/pytest_mesh.log
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -k another_device > $OUT/pytest_dev.log 2>&1; tail -3 $OUT/pytest_dev.log
