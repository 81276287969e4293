#!/bin/bash
# sustained UNet timing with plain polling waits vs hinted (sleeping) waits in conv_gemm_kernel, alternating on one box
for rep in 1 2; do
for h in 0 20000; do
  touch k-diffusion-inverse-problems_b200/csrc/conv_gemm.cu
  KDIP_NVCC_EXTRA="-DKDIP_MBAR_HINT_NS=$h" bash k-diffusion-inverse-problems_b200/csrc/build.sh > /tmp/build.log 2>&1 || { echo build failed; tail -3 /tmp/build.log; continue; }
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 500 > /tmp/clk_$h.csv 2>/dev/null &
  SMI=$!
  timeout 120 python tools/sustain_unet.py 32 10 2>/dev/null | tail -1 | sed "s/^/hint=$h /"
  kill $SMI 2>/dev/null
  sort -t, -k1 -n /tmp/clk_$h.csv | awk -F, 'NR>4{a[NR]=$1; p[NR]=$2} END{print "   clock/power samples under load (sorted, low end):", a[6], p[6], a[10], p[10]}'
done
done
