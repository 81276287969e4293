#!/bin/bash
# Regression experiment for the transform-warpgroup wait protocol: the repeatability test with the fix, then with the old protocol
# (KDIP_CONV_DBG=16), then the UNet timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -q -m gpu -k "repeatable or fused_groupnorm_apply" -p no:cacheprovider > gpurun_out/xfp_fixed.log 2>&1; echo "fixed rc=$?"; tail -2 gpurun_out/xfp_fixed.log
for i in 1 2 3; do
KDIP_CONV_DBG=16 timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -q -m gpu -k "repeatable" -p no:cacheprovider > gpurun_out/xfp_old$i.log 2>&1; echo "old protocol run $i rc=$?"; tail -3 gpurun_out/xfp_old$i.log | cut -c1-300
done
timeout 120 python tools/time_unet.py 32 5 > gpurun_out/time_unet1.log 2>&1; tail -1 gpurun_out/time_unet1.log
