#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r30.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t_r30.log
timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
KDIP_BENCH_SHAPES=0,1,2,3,5,6 timeout 200 python tools/bench_conv.py 32 20 2>&1 | tail -6
KDIP_CONV_HALO=1 timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
