#!/bin/bash
# round-2 second pass: full GPU test suite, timeline, bench (batch 32 / 256)
set -x
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 120 python tools/timeline.py > $O/timeline_b32.txt 2>&1
BATCH=256 timeout 120 python tools/timeline.py > $O/timeline_b256.txt 2>&1
timeout 400 python bench.py --steps 2000 --warmup 50 > $O/bench_b32.json 2> $O/bench_b32.err
timeout 300 python bench.py --batch 256 --steps 500 --warmup 20 --no-cpu > $O/bench_b256.json 2> $O/bench_b256.err
echo done
