#!/bin/bash
# software-pipelined fc1 update (multi-step train_fused) + 2-stage conv2_dgrad default: parity, period, timeline
O=gpurun_out/r2p; mkdir -p $O
run() { name=$1; shift; echo -n "$name " >> $O/periods.txt; env "$@" timeout -s KILL 120 python tools/period.py 2>&1 | tail -1 >> $O/periods.txt; }
run default A=1
run nodefer B200DQN_DEFER_FC1=0
run defer_pdl B200DQN_DEFER_PDL=1
run defer_c3 B200DQN_OPT_FC1_CTAS=3
run defer_c4 B200DQN_OPT_FC1_CTAS=4
run defer_s4 B200DQN_STAGES2=none
run nodefer_s4 B200DQN_DEFER_FC1=0 B200DQN_STAGES2=none
cat $O/periods.txt
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
B200DQN_DEFER_PDL=1 timeout -s KILL 120 python tools/timeline.py > $O/timeline_defer_pdl.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_optimizers.py tests/test_gpu_replay.py tests/test_gpu_agent_loop.py tests/test_gpu_checkpoint.py -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout -s KILL 300 python bench.py --steps 2000 --warmup 50 --no-cpu > $O/bench.json 2> $O/bench.err
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_steps20.json 2> $O/bench_steps20.err
echo done
