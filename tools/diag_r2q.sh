#!/bin/bash
# one shared-memory carveout for every kernel: period of the schedule variants
O=gpurun_out/r2q; mkdir -p $O
run() { name=$1; shift; echo -n "$name " >> $O/periods.txt; env "$@" timeout -s KILL 120 python tools/period.py 2>&1 | tail -1 >> $O/periods.txt; }
D=conv2_dgrad,conv3_dgrad,fc1_dgrad
run default A=1
run nodefer B200DQN_DEFER_FC1=0
run old_carveout_nodefer B200DQN_CARVEOUT=0 B200DQN_DEFER_FC1=0
run nodefer_c3 B200DQN_DEFER_FC1=0 B200DQN_OPT_FC1_CTAS=3
run nodefer_c4 B200DQN_DEFER_FC1=0 B200DQN_OPT_FC1_CTAS=4
run defer_c3 B200DQN_OPT_FC1_CTAS=3
run defer_c4 B200DQN_OPT_FC1_CTAS=4
run nodefer_s4 B200DQN_DEFER_FC1=0 B200DQN_STAGES2=none
run nodefer_splitk B200DQN_DEFER_FC1=0 B200DQN_SPLITK=1
run nodefer_etD B200DQN_DEFER_FC1=0 B200DQN_EARLY_TRIGGER=$D
run nodefer_fc1splits13 B200DQN_DEFER_FC1=0 B200DQN_FC1_SPLITS=13
run nodefer_fc1fused B200DQN_DEFER_FC1=0 B200DQN_FC1_FUSED=1
run nodefer_s2dw B200DQN_DEFER_FC1=0 B200DQN_STAGES2=conv2_dgrad,conv2_wgrad,conv3_wgrad
run defer_s4 B200DQN_STAGES2=none
run nodefer_again B200DQN_DEFER_FC1=0
cat $O/periods.txt
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
B200DQN_DEFER_FC1=0 timeout -s KILL 120 python tools/timeline.py > $O/timeline_nodefer.txt 2>&1
echo done
