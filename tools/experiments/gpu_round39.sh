#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py tests/test_guidance_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r39.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r39.log
timeout 300 python tools/time_unet.py 32 50 2>&1 | tail -1
KDIP_BENCH_SHAPES=10,7,11,12,13,14,4,5 timeout 120 python tools/bench_conv.py 32 30 2>&1 | grep "TF/s"
