#!/bin/bash
# scheduling experiments on one GPU: where do the late side-branch launches come from?
set -x
O=gpurun_out/r2j; mkdir -p $O
run() { name=$1; shift; env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1; }
run base A=1
run nopdl_conv2dgrad B200DQN_NOPDL_OPS=0x20
run nopdl_dgrads B200DQN_NOPDL_OPS=0x68
run nopdl_conv1wgrad B200DQN_NOPDL_OPS=0x40
run side_hi B200DQN_SIDE_PRIO=hi
run main_hi B200DQN_STREAM_PRIO=hi B200DQN_SIDE_PRIO=lo
run wgrad_one_stream B200DQN_WGRAD_ONE_STREAM=1
run optfc1_1 B200DQN_OPT_FC1_CTAS=1
run main_hi_optfc1_1 B200DQN_STREAM_PRIO=hi B200DQN_SIDE_PRIO=lo B200DQN_OPT_FC1_CTAS=1
for f in $O/timeline_*.txt; do echo "$f $(tail -1 $f)"; done
echo done
