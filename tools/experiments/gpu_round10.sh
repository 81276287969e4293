#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r10.log 2>&1; echo "tests exit $?"; tail -n 8 gpurun_out/t_r10.log
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r10_time.log 2>&1
KDIP_HALO_PAIR=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r10_nopair_time.log 2>&1
tail -n 1 gpurun_out/r10_time.log gpurun_out/r10_nopair_time.log
KDIP_BENCH_SHAPES=0,1,2,3 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r10_bench_conv.log 2>&1; cat gpurun_out/r10_bench_conv.log
KDIP_HALO_PAIR=0 KDIP_BENCH_SHAPES=0,1,2,3 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r10_bench_conv_nopair.log 2>&1; cat gpurun_out/r10_bench_conv_nopair.log
