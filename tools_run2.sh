TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^\*|NCCL version|OMP_NUM'
echo "=== gather grads"; GRADS=1 timeout 100 $TR --master-port 29502 tools_mgpu_check.py 2>&1 | grep -E "layer|status|rror" | sort | tail -14
echo "=== gather"; TIMELINE=1 timeout 150 $TR --master-port 29503 tools_mgpu_check.py 2>&1 | grep -vE "$F" | grep -E "rank 0|----|^[a-z_0-9]+ +[0-9]|rror" | tail -44
echo "=== bench"; timeout 200 $TR --master-port 29504 bench.py --gpus 2 --steps 2000 --warmup 50 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_r1k_n2.log | cut -c1-300
