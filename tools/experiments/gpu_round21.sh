#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_unet_gpu.py tests/test_guidance_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r21.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t_r21.log
timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
KDIP_UNFUSED_SKIPADD=1 timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
