#!/bin/bash
mkdir -p gpurun_out
for halo in 1 0; do for ws in 0 1; do
  echo "== halo=$halo ws=$ws"; KDIP_CONV_HALO=$halo KDIP_CONV_WS=$ws timeout 200 python tools/time_unet.py 32 4 2>&1 | tail -1
done; done
echo "== again halo=1 ws=1"; KDIP_CONV_HALO=1 KDIP_CONV_WS=1 timeout 200 python tools/time_unet.py 32 4 2>&1 | tail -1
KDIP_CONV_DBG=3 KDIP_BENCH_SHAPES=0 timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -1
