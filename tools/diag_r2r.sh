#!/bin/bash
# 2 x B200: the final build (one fc1 image, 2-stage conv2_dgrad) through the multi-GPU tests and the bench
O=gpurun_out/r2r; mkdir -p $O
B200DQN_TEST_WORLDS=2 timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/pytest_w2.log 2>&1; echo "rc=$?" >> $O/pytest_w2.log
tail -3 $O/pytest_w2.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2000 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 600 $O/bench_n2.json
echo done
