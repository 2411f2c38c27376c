#!/bin/bash
# multi-GPU development pass: $1 = number of GPUs on the box (2 or 8)
set -x
N=${1:-2}
O=gpurun_out/r2n$N; mkdir -p $O
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
if [ "$N" = "2" ]; then
  B200DQN_FC1_FUSED=1 timeout -s KILL 120 python tools/timeline.py > $O/timeline_1gpu_fc1fused.txt 2>&1
  B200DQN_FC1_FUSED=1 timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -m gpu -q --maxfail=20 -k "train_step_parity or fused_ring or trajectory or rmsprop" > $O/pytest_fc1fused.log 2>&1; echo "rc=$?" >> $O/pytest_fc1fused.log
  B200DQN_TEST_WORLDS=2 timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --maxfail=20 -k "oracle or match_nccl" > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
  TIMELINE=1 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/mgpu_check.py > $O/timeline_w2.txt 2>&1
  TIMELINE=1 B200DQN_DZ_LL=0 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tools/mgpu_check.py > $O/timeline_w2_plainpush.txt 2>&1
  TIMELINE=1 B200DQN_FUSED_XLL=0 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 tools/mgpu_check.py > $O/timeline_w2_unfusedxll.txt 2>&1
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus 2 --steps 1000 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
else
  # 8-GPU box: charged 8x — most important first, every step under a hard cap
  export B200DQN_TEST_TIMEOUT=150
  B200DQN_TEST_WORLDS=8 timeout -s KILL 170 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "oracle and p2p-gather and not unfused" > $O/pytest_w8_default.log 2>&1; echo "rc=$?" >> $O/pytest_w8_default.log
  TIMELINE=1 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 tools/mgpu_check.py > $O/timeline_w8.txt 2>&1
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 bench.py --gpus 8 --steps 1000 --warmup 50 > $O/bench_n8.json 2> $O/bench_n8.err
  B200DQN_TEST_WORLDS=4 timeout -s KILL 170 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "oracle and p2p-gather and not unfused" > $O/pytest_w4_default.log 2>&1; echo "rc=$?" >> $O/pytest_w4_default.log
  TIMELINE=1 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 tools/mgpu_check.py > $O/timeline_w4.txt 2>&1
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 4 --steps 1000 --warmup 50 > $O/bench_n4.json 2> $O/bench_n4.err
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 1000 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
  B200DQN_TEST_WORLDS=8 timeout -s KILL 260 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "oracle and (nccl or two-shot)" > $O/pytest_w8_others.log 2>&1; echo "rc=$?" >> $O/pytest_w8_others.log
fi
echo done
