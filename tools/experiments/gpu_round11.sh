#!/bin/bash
mkdir -p gpurun_out
export KDIP_BENCH_SHAPES=0
KDIP_HALO_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_halo python tools/bench_conv.py 16 2 > gpurun_out/full_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_halo_pair python tools/bench_conv.py 16 2 > gpurun_out/full_halo_pair.log 2>&1
KDIP_CONV_HALO=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/full_v1mt2 python tools/bench_conv.py 16 2 > gpurun_out/full_v1mt2.log 2>&1
ls -la gpurun_out/*.ncu-rep
