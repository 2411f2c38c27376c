TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^\*|NCCL version|OMP_NUM'
echo "=== gather"; TIMELINE=1 timeout 120 $TR --master-port 29503 tools_mgpu_check.py 2>&1 | grep -vE "$F" | grep -E "rank 0|---- rank|^[a-z_0-9]+ +[0-9]|rror" | head -48 > gpurun_out/r1_timeline_n2.txt; head -42 gpurun_out/r1_timeline_n2.txt
echo "=== bench"; timeout 200 $TR --master-port 29504 bench.py --gpus 2 --steps 3000 --warmup 100 --no-cpu 2>&1 | tail -1 | tee gpurun_out/r1_bench_n2.json | cut -c1-300
