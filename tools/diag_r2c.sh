#!/bin/bash
# round-2 third pass: (a) everything with the round-1 kernel set (split-K off, conv1 by LDG), (b) + conv1 by TMA,
# (c) + cluster split-K, (d) both (the default) with the full suite
set -x
O=gpurun_out/r2c; mkdir -p $O
Q="tests/test_gpu_net.py -k predict_parity or train_step_parity or fused_ring or trajectory"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -s KILL 400 python -m pytest tests/test_gpu_net.py -m gpu -q --maxfail=20 -k "predict_parity or train_step_parity or fused_ring or trajectory" > $O/pytest_$name.log 2>&1
  echo "rc=$?" >> $O/pytest_$name.log
  env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1
}
env B200DQN_SPLITK=0 B200DQN_CONV1=ldg timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_full_base.log 2>&1; echo "rc=$?" >> $O/pytest_full_base.log
env B200DQN_SPLITK=0 B200DQN_CONV1=ldg timeout -s KILL 120 python tools/timeline.py > $O/timeline_base.txt 2>&1
run tma B200DQN_SPLITK=0 B200DQN_CONV1=tma
run splitk B200DQN_SPLITK=1 B200DQN_CONV1=ldg
run both B200DQN_SPLITK=1 B200DQN_CONV1=tma
timeout -s KILL 400 python bench.py --steps 2000 --warmup 50 > $O/bench_b32.json 2> $O/bench_b32.err
BATCH=256 timeout -s KILL 120 python tools/timeline.py > $O/timeline_both_b256.txt 2>&1
echo done
