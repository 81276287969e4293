#!/bin/bash
mkdir -p gpurun_out
KDIP_BENCH_SHAPES=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_conv_r1e python tools/bench_conv.py 16 2 > gpurun_out/full_conv_r1e.log 2>&1
KDIP_BENCH_SHAPES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_conv_r1e_res python tools/bench_conv.py 16 2 > gpurun_out/full_conv_r1e_res.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gn_bwd_apply -s 20 -c 1 -f -o gpurun_out/full_gnbwd_r1e python tools/time_unet.py 8 1 > gpurun_out/full_gnbwd_r1e.log 2>&1
ls -la gpurun_out/full_*r1e*.ncu-rep
