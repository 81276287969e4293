#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r34.csv python tools/time_unet.py 32 1 > gpurun_out/r34_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/r34_bench.json 2> gpurun_out/r34_bench.err; cut -c1-260 gpurun_out/r34_bench.json
