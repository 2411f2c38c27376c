#!/bin/bash
# 4-GPU box (charged 4x): sanity of the last single-GPU change, then the W = 4 / W = 2 data-parallel path
set -x
O=gpurun_out/r2h; mkdir -p $O
export B200DQN_TEST_TIMEOUT=150
timeout -s KILL 200 python -m pytest tests/test_gpu_net.py -m gpu -q --maxfail=10 -k "train_step_parity or fused_ring or predict_parity" > $O/pytest_1gpu.log 2>&1; echo "rc=$?" >> $O/pytest_1gpu.log
timeout -s KILL 100 python tools/timeline.py > $O/timeline_1gpu.txt 2>&1
B200DQN_TEST_WORLDS=4 timeout -s KILL 330 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "oracle and p2p-gather" > $O/pytest_w4.log 2>&1; echo "rc=$?" >> $O/pytest_w4.log
B200DQN_TEST_WORLDS=2 timeout -s KILL 170 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "oracle and p2p-gather and not unfused" > $O/pytest_w2.log 2>&1; echo "rc=$?" >> $O/pytest_w2.log
TIMELINE=1 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 tools/mgpu_check.py > $O/timeline_w4.txt 2>&1
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 4 --steps 1000 --warmup 50 > $O/bench_n4.json 2> $O/bench_n4.err
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 1000 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
B200DQN_FUSED_XLL=1 timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29725 bench.py --gpus 4 --steps 1000 --warmup 50 > $O/bench_n4_fusedxll.json 2> $O/bench_n4_fusedxll.err
echo done
