#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_cfg5.csv python tools/time_configs.py 32 1 > gpurun_out/cfg5_ncu.log 2>&1
tail -3 gpurun_out/cfg5_ncu.log
