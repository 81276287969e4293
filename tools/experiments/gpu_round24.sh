#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_operators_gpu.py tests/test_guidance_gpu.py tests/test_elementwise_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_r24.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/t_r24.log
timeout 600 python tools/time_guided.py 32 5 2>&1 | grep -E "guidance=|bare"
KDIP_FFT_RADIX2=1 timeout 600 python tools/time_guided.py 32 5 2>&1 | grep -E "guidance=|bare"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"rows_|cols_" -c 120 --csv --log-file gpurun_out/fft256_dram.csv python tools/time_guided.py 32 1 > gpurun_out/r24_ncu.log 2>&1
