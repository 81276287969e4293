#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/t_r5.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r5.log
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r5_time.log 2>&1
KDIP_CONV_PAIRMT=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r5_nopairmt_time.log 2>&1
tail -n 1 gpurun_out/r5_time.log gpurun_out/r5_nopairmt_time.log
KDIP_BENCH_SHAPES=0,1,3 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r5_bench_conv.log 2>&1; cat gpurun_out/r5_bench_conv.log
timeout 900 python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; cat gpurun_out/r5_bench.json
