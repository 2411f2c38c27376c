#!/bin/bash
# 2-stage operand rings (two CTAs per SM) for conv2_dgrad and the conv wgrads, with and without early dependency release
O=gpurun_out/r2n; mkdir -p $O
run() { name=$1; shift; env "$@" timeout -s KILL 120 python tools/timeline.py > $O/timeline_$name.txt 2>&1; }
E=conv2_dgrad,conv3_dgrad
D=conv2_dgrad,conv3_dgrad,fc1_dgrad
W=conv2_wgrad,conv3_wgrad
run base A=1
run s2d B200DQN_STAGES2=conv2_dgrad
run s2w B200DQN_STAGES2=$W
run s2dw B200DQN_STAGES2=conv2_dgrad,$W
run et32 B200DQN_EARLY_TRIGGER=$E
run et32_s2d B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad
run et32_s2w B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=$W
run et32_s2dw B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad,$W
run et_s2dw B200DQN_EARLY_TRIGGER=$D B200DQN_STAGES2=conv2_dgrad,$W
run et32_s2dw_c3 B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad,$W B200DQN_OPT_FC1_CTAS=3
run s2dw_c3 B200DQN_STAGES2=conv2_dgrad,$W B200DQN_OPT_FC1_CTAS=3
run et32_s2d_c3 B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad B200DQN_OPT_FC1_CTAS=3
run et32_s2c2 B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad,conv2_wgrad
run et32_c1w_s2dw B200DQN_EARLY_TRIGGER=$E,conv1_wgrad B200DQN_STAGES2=conv2_dgrad,$W
for f in $O/timeline_*.txt; do echo "$f $(tail -1 $f)"; done
B200DQN_STAGES2=conv2_dgrad,$W timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -m gpu -q -x > $O/pytest_s2dw.log 2>&1; echo "rc=$?" >> $O/pytest_s2dw.log
B200DQN_EARLY_TRIGGER=$E B200DQN_STAGES2=conv2_dgrad,$W timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_et32_s2dw.json 2> $O/bench_et32_s2dw.err
B200DQN_EARLY_TRIGGER=$E timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench_et32.json 2> $O/bench_et32.err
echo done
