#!/bin/bash
# scheduling experiments, second batch: what holds back the launches between 47 and 57 us?
O=gpurun_out/r2k; mkdir -p $O
run() { name=$1; shift; env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1; }
run base A=1
run fc1_skip B200DQN_OPT_FC1_WHEN=skip
run fc1_ev2 B200DQN_OPT_FC1_WHEN=ev2
run fc1_ev3 B200DQN_OPT_FC1_WHEN=ev3
run fc1_last B200DQN_OPT_FC1_WHEN=last
run fc1_ctas3 B200DQN_OPT_FC1_CTAS=3
run fc1_ctas4 B200DQN_OPT_FC1_CTAS=4
run et_conv2dgrad B200DQN_EARLY_TRIGGER=conv2_dgrad
run et_dgrads B200DQN_EARLY_TRIGGER=conv2_dgrad,conv3_dgrad,fc1_dgrad
run et_conv2dgrad_nopdl_c1w B200DQN_EARLY_TRIGGER=conv2_dgrad B200DQN_NOPDL_OPS=0x40
run et_all B200DQN_EARLY_TRIGGER=conv2_fwd,conv3_fwd,fc1_fwd,fc1_dgrad,conv3_dgrad,conv2_dgrad,conv1_wgrad
run fc1_skip_et_conv2dgrad B200DQN_OPT_FC1_WHEN=skip B200DQN_EARLY_TRIGGER=conv2_dgrad
run fc1_ev3_onepass B200DQN_OPT_FC1_WHEN=ev3 B200DQN_OPT_FC1_ONEPASS=1
run base2 A=1
for f in $O/timeline_*.txt; do echo "$f $(tail -1 $f)"; done
echo done
