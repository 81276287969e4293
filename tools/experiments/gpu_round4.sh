#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_layers_gpu.py tests/test_unet_gpu.py tests/test_guidance_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/t_r4.log 2>&1; echo "tests exit $?"; tail -n 5 gpurun_out/t_r4.log
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r4_time.log 2>&1
KDIP_CONV_MT=1 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/r4_mt1_time.log 2>&1
tail -n 1 gpurun_out/r4_time.log gpurun_out/r4_mt1_time.log
timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r4_bench_conv.log 2>&1
KDIP_CONV_MT=1 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r4_bench_conv_mt1.log 2>&1
KDIP_CONV_MT=1 KDIP_CONV_PAIR=0 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r4_bench_conv_mt1_nopair.log 2>&1
KDIP_CONV_STAGES=2 timeout 300 python tools/bench_conv.py 32 10 > gpurun_out/r4_bench_conv_st2.log 2>&1
paste gpurun_out/r4_bench_conv.log <(cut -c33- gpurun_out/r4_bench_conv_mt1.log) <(cut -c33- gpurun_out/r4_bench_conv_mt1_nopair.log) <(cut -c33- gpurun_out/r4_bench_conv_st2.log)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r4.csv python tools/time_unet.py 32 1 > gpurun_out/r4_ncu.log 2>&1
