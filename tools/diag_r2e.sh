#!/bin/bash
set -x
O=gpurun_out/r2e; mkdir -p $O
timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
B200DQN_FUSE_SAMPLE=0 timeout -s KILL 120 python tools/timeline.py > $O/timeline_nofuse.txt 2>&1
timeout -s KILL 400 python bench.py --steps 2000 --warmup 50 > $O/bench_b32.json 2> $O/bench_b32.err
B200DQN_FUSE_SAMPLE=0 B200DQN_CONV1=ldg timeout -s KILL 400 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_b32_ldg.json 2> $O/bench_b32_ldg.err
echo done
