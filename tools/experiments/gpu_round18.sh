#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/ci.log 2>&1; cat gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err; cut -c1-400 gpurun_out/r18_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r18.csv python tools/time_unet.py 32 1 > gpurun_out/r18_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_gemm -c 700 --csv --log-file gpurun_out/conv_dram_r18.csv python tools/time_unet.py 32 1 > gpurun_out/r18_ncu2.log 2>&1
tail -2 gpurun_out/r18_ncu2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
