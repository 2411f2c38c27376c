#!/bin/bash
# one fc1 image (forward reads the row-oriented image MN-major, no pack_fc1f): parity + timelines
O=gpurun_out/r2m; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_optimizers.py tests/test_gpu_checkpoint.py -m gpu -q -x > $O/pytest_net.log 2>&1; echo "rc=$?" >> $O/pytest_net.log
run() { name=$1; shift; env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1; }
D=conv2_dgrad,conv3_dgrad,fc1_dgrad
run base A=1
run et B200DQN_EARLY_TRIGGER=$D
run et_c3 B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3
run c3 B200DQN_OPT_FC1_CTAS=3
run et32 B200DQN_EARLY_TRIGGER=conv2_dgrad,conv3_dgrad
run et2f B200DQN_EARLY_TRIGGER=conv2_dgrad,fc1_dgrad
run et2 B200DQN_EARLY_TRIGGER=conv2_dgrad
run et_sidehi B200DQN_EARLY_TRIGGER=$D B200DQN_SIDE_PRIO=hi
run et_c1w B200DQN_EARLY_TRIGGER=$D,conv1_wgrad
run base2 A=1
for f in $O/timeline_*.txt; do echo "$f $(tail -1 $f)"; done
timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 > $O/bench_base.json 2> $O/bench_base.err
B200DQN_EARLY_TRIGGER=$D timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_et.json 2> $O/bench_et.err
echo done
