#!/bin/bash
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s -x --deselect tests/test_variants_gpu.py > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 1500 $OUT/bench.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_run.py > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"
tail -5 $OUT/memcheck.log
