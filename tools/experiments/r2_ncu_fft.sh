#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_fft.py 32 20 > gpurun_out/time_fft.log 2>&1; cat gpurun_out/time_fft.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rows_r2c_256|cols_256|rows_c2r_256" -s 30 -c 3 -f -o gpurun_out/full_fft_r2 python tools/time_fft.py 32 20 > gpurun_out/full_fft_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
