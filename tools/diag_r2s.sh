#!/bin/bash
# 2 x B200: is the fused loop device-bound or host-launch-bound?  (period vs host enqueue cost, two ring depths)
O=gpurun_out/r2s; mkdir -p $O
tr() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/mgpu_check.py; }
tr 29751 > $O/w2_default.txt 2>&1
B200DQN_STAGES2=none tr 29752 > $O/w2_stages4.txt 2>&1
grep -h "us/step" $O/w2_default.txt | head -12; echo ---; grep -h "us/step" $O/w2_stages4.txt | head -12
echo done
