#!/bin/bash
# round-2 first diagnostic pass: timelines at batch 32/256/1024, in-kernel traces, bench at larger batches
set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
for B in 32 256 1024; do BATCH=$B timeout 120 python tools/timeline.py > $O/timeline_b$B.txt 2>&1; done
for L in conv1_fwd conv2_fwd conv3_fwd fc1_fwd fc1_dgrad conv3_dgrad conv2_dgrad; do
  B200DQN_TRACE_LABEL=$L timeout 60 python tools/trace.py > $O/trace_$L.txt 2>&1
done
timeout 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_b32.json 2> $O/bench_b32.err
timeout 300 python bench.py --batch 256 --steps 500 --warmup 20 --no-cpu > $O/bench_b256.json 2> $O/bench_b256.err
timeout 300 python bench.py --batch 1024 --steps 200 --warmup 10 --no-cpu --replay 200000 > $O/bench_b1024.json 2> $O/bench_b1024.err
echo done
