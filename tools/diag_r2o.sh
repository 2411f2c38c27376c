#!/bin/bash
# steady-state PERIOD (not the traced span) of the schedule variants
O=gpurun_out/r2o; mkdir -p $O
run() { name=$1; shift; echo -n "$name " >> $O/periods.txt; env "$@" timeout -s KILL 120 python tools/period.py 2>&1 | tail -1 >> $O/periods.txt; }
E=conv2_dgrad,conv3_dgrad
D=conv2_dgrad,conv3_dgrad,fc1_dgrad
W=conv2_wgrad,conv3_wgrad
run base A=1
run s2d B200DQN_STAGES2=conv2_dgrad
run et32 B200DQN_EARLY_TRIGGER=$E
run et32_s2d B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad
run etD B200DQN_EARLY_TRIGGER=$D
run etD_s2d B200DQN_EARLY_TRIGGER=$D B200DQN_STAGES2=conv2_dgrad
run base_again A=1
run c3 B200DQN_OPT_FC1_CTAS=3
run s2d_c3 B200DQN_STAGES2=conv2_dgrad B200DQN_OPT_FC1_CTAS=3
run et32_s2d_c3 B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad B200DQN_OPT_FC1_CTAS=3
run s2dw B200DQN_STAGES2=conv2_dgrad,$W
run et2 B200DQN_EARLY_TRIGGER=conv2_dgrad
run et2_s2d B200DQN_EARLY_TRIGGER=conv2_dgrad B200DQN_STAGES2=conv2_dgrad
run etall_s2d B200DQN_EARLY_TRIGGER=conv2_fwd,conv3_fwd,fc1_fwd,$D,conv1_wgrad B200DQN_STAGES2=conv2_dgrad
run fc1_skip B200DQN_OPT_FC1_WHEN=skip
run fc1_skip_s2d_et32 B200DQN_OPT_FC1_WHEN=skip B200DQN_STAGES2=conv2_dgrad B200DQN_EARLY_TRIGGER=$E
run side_hi_s2d B200DQN_SIDE_PRIO=hi B200DQN_STAGES2=conv2_dgrad
run main_hi_s2d B200DQN_STREAM_PRIO=hi B200DQN_SIDE_PRIO=lo B200DQN_STAGES2=conv2_dgrad
run base_third A=1
cat $O/periods.txt
