#!/bin/bash
# round 2: fused GroupNorm apply on the conv operand path - parity, then A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x -k "fused_groupnorm_apply" -p no:cacheprovider > gpurun_out/xf_conv.log 2>&1; echo "conv xf exit $?"
tail -n 15 gpurun_out/xf_conv.log
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/xf_unet.log 2>&1; echo "conv+unet exit $?"
tail -n 8 gpurun_out/xf_unet.log
KDIP_FUSE_GNAPPLY=0 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
KDIP_FUSE_GNAPPLY=0 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
