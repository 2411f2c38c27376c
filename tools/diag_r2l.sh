#!/bin/bash
# scheduling experiments, third batch: early dependency release on the dgrad chain + a faster fc1 optimizer
O=gpurun_out/r2l; mkdir -p $O
run() { name=$1; shift; env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1; }
D=conv2_dgrad,conv3_dgrad,fc1_dgrad
run base A=1
run et B200DQN_EARLY_TRIGGER=$D
run et_c3 B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3
run et_c4 B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=4
run et_sidehi B200DQN_EARLY_TRIGGER=$D B200DQN_SIDE_PRIO=hi
run et_sidehi_c3 B200DQN_EARLY_TRIGGER=$D B200DQN_SIDE_PRIO=hi B200DQN_OPT_FC1_CTAS=3
run et_mainhi_c3 B200DQN_EARLY_TRIGGER=$D B200DQN_STREAM_PRIO=hi B200DQN_SIDE_PRIO=lo B200DQN_OPT_FC1_CTAS=3
run et32_c3 B200DQN_EARLY_TRIGGER=conv2_dgrad,conv3_dgrad B200DQN_OPT_FC1_CTAS=3
run et2f_c3 B200DQN_EARLY_TRIGGER=conv2_dgrad,fc1_dgrad B200DQN_OPT_FC1_CTAS=3
run etall_c3 B200DQN_EARLY_TRIGGER=conv2_fwd,conv3_fwd,fc1_fwd,$D,conv1_wgrad B200DQN_OPT_FC1_CTAS=3
run et_c1w_c3 B200DQN_EARLY_TRIGGER=$D,conv1_wgrad B200DQN_OPT_FC1_CTAS=3
run et_onepass B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_ONEPASS=1
run et_c3_onestream B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3 B200DQN_WGRAD_ONE_STREAM=1
run et_c3_nopdl_c1w B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3 B200DQN_NOPDL_OPS=0x40
run et_c3_again B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3
for f in $O/timeline_*.txt; do echo "$f $(tail -1 $f)"; done
# the same through bench.py (2000 steps, not the 16-step trace)
B200DQN_EARLY_TRIGGER=$D B200DQN_OPT_FC1_CTAS=3 timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 > $O/bench_et_c3.json 2> $O/bench_et_c3.err
B200DQN_EARLY_TRIGGER=$D timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 > $O/bench_et.json 2> $O/bench_et.err
echo done
