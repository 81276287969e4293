#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/halo_probe.py > gpurun_out/halo_probe.log 2>&1; cat gpurun_out/halo_probe.log | head -30
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r8.log 2>&1; echo "tests exit $?"; tail -n 12 gpurun_out/t_r8.log
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r8_time.log 2>&1
KDIP_CONV_HALO=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r8_nohalo_time.log 2>&1
tail -n 1 gpurun_out/r8_time.log gpurun_out/r8_nohalo_time.log
KDIP_BENCH_SHAPES=0,1,2,3 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r8_bench_conv.log 2>&1; cat gpurun_out/r8_bench_conv.log
KDIP_CONV_HALO=0 KDIP_BENCH_SHAPES=0,1,2,3 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r8_bench_conv_nohalo.log 2>&1; cat gpurun_out/r8_bench_conv_nohalo.log
