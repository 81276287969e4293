#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3
export KDIP_BENCH_SHAPES=0,9,1,3,2
for sl in 6 7 8; do
echo "== unfused slots $sl"; KDIP_HALO_SLOTS=$sl timeout 200 python tools/bench_conv.py 32 20
echo "== fused slots $sl"; KDIP_HALO_SLOTS=$sl KDIP_BENCH_XF=1 timeout 200 python tools/bench_conv.py 32 20
done
unset KDIP_BENCH_SHAPES
KDIP_FUSE_GNAPPLY=0 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
KDIP_HALO_SLOTS=7 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
