#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_conv_gemm_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r13.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r13.log
timeout 600 python tools/time_guided.py 32 5 2>&1 | grep -v Warning | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_guided.csv python tools/time_guided.py 32 1 > gpurun_out/guided_ncu.log 2>&1
