#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/ci.log 2>&1; cat gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv python tools/time_unet.py 32 1 > gpurun_out/final_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
