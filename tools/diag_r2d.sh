#!/bin/bash
set -x
O=gpurun_out/r2d; mkdir -p $O
timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
B200DQN_OPT_FC1_CTAS=1 timeout -s KILL 120 python tools/timeline.py > $O/timeline_optfc1_1.txt 2>&1
B200DQN_OPT_FC1_CTAS=-2 timeout -s KILL 120 python tools/timeline.py > $O/timeline_optfc1_half.txt 2>&1
B200DQN_CONV1=ldg timeout -s KILL 120 python tools/timeline.py > $O/timeline_ldg.txt 2>&1
timeout -s KILL 400 python bench.py --steps 2000 --warmup 50 > $O/bench_b32.json 2> $O/bench_b32.err
timeout -s KILL 300 python bench.py --batch 256 --steps 500 --warmup 20 --no-cpu > $O/bench_b256.json 2> $O/bench_b256.err
timeout -s KILL 300 python bench.py --batch 1024 --steps 200 --warmup 10 --no-cpu --replay 200000 > $O/bench_b1024.json 2> $O/bench_b1024.err
K='regex:k_sample|k_conv1_tma|k_umma|k_head|k_opt|k_cost|k_pack_image'
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 420 -c 40 --csv --log-file $O/launches.csv python tools/ncu_step.py > $O/ncu_list.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "$K" -s 420 -c 20 -o $O/prof_step python tools/ncu_step.py > $O/ncu_full.log 2>&1
ls -la $O
echo done
