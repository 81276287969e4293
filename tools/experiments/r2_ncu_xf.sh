#!/bin/bash
mkdir -p gpurun_out
KDIP_BENCH_XF=1 KDIP_BENCH_SHAPES=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_conv_r2_xf python tools/bench_conv.py 16 2 > gpurun_out/full_conv_r2_xf.log 2>&1
ls -la gpurun_out/full_conv_r2_xf.ncu-rep
