#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/time_unet.py 32 50 2>&1 | tail -1; }
run A=0
run KDIP_CONV_PAIRMT=0
run KDIP_HALO_PAIR=0
run KDIP_CONV_PAIR=0
run A=0
bash tools/gpu_ci.sh > gpurun_out/ci.log 2>&1; cat gpurun_out/summary.txt
