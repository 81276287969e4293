#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r31.csv python tools/time_unet.py 32 1 > gpurun_out/r31_ncu.log 2>&1
for pair in 1 0; do echo "== KDIP_CONV_PAIR=$pair"; KDIP_CONV_PAIR=$pair KDIP_BENCH_SHAPES=2,4,5,6 timeout 200 python tools/bench_conv.py 32 20 2>&1 | tail -4; done
KDIP_CONV_PAIR=0 timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
timeout 300 python tools/time_unet.py 32 60 2>&1 | tail -1
