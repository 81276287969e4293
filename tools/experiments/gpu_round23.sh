#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"pmv|combine|heun|euler|churn|cg_|dot_partial|rows_|cols_|lincomb|dwt|psf" -c 400 --csv --log-file gpurun_out/guidance_dram.csv python tools/time_guided.py 32 1 > gpurun_out/r23_ncu.log 2>&1
tail -3 gpurun_out/r23_ncu.log
timeout 900 python bench.py > gpurun_out/r23_bench.json 2> gpurun_out/r23_bench.err; cut -c1-300 gpurun_out/r23_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r23.csv python tools/time_unet.py 32 1 > gpurun_out/r23_ncu2.log 2>&1
