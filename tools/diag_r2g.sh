#!/bin/bash
set -x
O=gpurun_out/r2g; mkdir -p $O
B200DQN_OPT_FC1_ONEPASS=1 timeout -s KILL 400 python -m pytest tests/test_gpu_net.py tests/test_gpu_optimizers.py tests/test_gpu_checkpoint.py -m gpu -q --maxfail=20 > $O/pytest_onepass.log 2>&1; echo "rc=$?" >> $O/pytest_onepass.log
B200DQN_OPT_FC1_ONEPASS=1 timeout -s KILL 120 python tools/timeline.py > $O/timeline_onepass.txt 2>&1
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
timeout -s KILL 300 python -m pytest tests/test_gpu_agent_loop.py tests/test_gpu_replay.py -m gpu -q --maxfail=20 -k "step_host or stale" > $O/pytest_new.log 2>&1; echo "rc=$?" >> $O/pytest_new.log
B200DQN_OPT_FC1_ONEPASS=1 timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_onepass.json 2> $O/bench_onepass.err
echo done
