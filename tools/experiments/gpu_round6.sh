#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_layers_gpu.py tests/test_unet_gpu.py tests/test_guidance_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/t_r6.log 2>&1; echo "tests exit $?"; tail -n 5 gpurun_out/t_r6.log
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r6_time.log 2>&1
KDIP_UNFUSED_GNRED=1 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r6_unfused_time.log 2>&1
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r6_time2.log 2>&1
tail -n 1 gpurun_out/r6_time.log gpurun_out/r6_unfused_time.log gpurun_out/r6_time2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r6.csv python tools/time_unet.py 32 1 > gpurun_out/r6_ncu.log 2>&1
