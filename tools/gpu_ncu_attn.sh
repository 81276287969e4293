#!/bin/bash
# One `ncu --set full` capture of each streamed attention kernel (T = 1024, 8 heads, 16 images) -> gpurun_out/full_attn_tcs.ncu-rep
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tcs -s 6 -c 3 -f -o gpurun_out/full_attn_tcs python tools/time_attn.py 16 > gpurun_out/full_attn_tcs.log 2>&1
ncu -i gpurun_out/full_attn_tcs.ncu-rep --page raw --csv > gpurun_out/full_attn_tcs_raw.csv 2>/dev/null
ls -la gpurun_out/full_attn_tcs*
