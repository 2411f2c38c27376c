#!/bin/bash
set -x
O=gpurun_out/r2t; mkdir -p $O
timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout -s KILL 120 python tools/timeline.py > $O/timeline_default.txt 2>&1
timeout -s KILL 400 python bench.py --steps 2000 --warmup 50 > $O/bench_b32.json 2> $O/bench_b32.err
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 > $O/bench_b32_driver_style.json 2> $O/bench_b32_driver_style.err
timeout -s KILL 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo done
timeout -s KILL 100 python tools/period.py > gpurun_out/r2t/period.txt 2>&1
