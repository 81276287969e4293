#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_unet_gpu.py tests/test_guidance_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r22.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t_r22.log
timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r22.csv python tools/time_unet.py 32 1 > gpurun_out/r22_ncu.log 2>&1
