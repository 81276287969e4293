#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=power.draw,clocks.sm,temperature.gpu --format=csv,noheader
for rep in 1 2; do
for halo in 1 0; do for ws in 0 1; do
  echo "== halo=$halo ws=$ws"; KDIP_CONV_HALO=$halo KDIP_CONV_WS=$ws timeout 300 python tools/time_unet.py 32 80 2>&1 | tail -1
  nvidia-smi --query-gpu=power.draw,clocks.sm,temperature.gpu --format=csv,noheader
done; done; done
