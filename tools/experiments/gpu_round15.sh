#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.sm,temperature.gpu --format=csv
export KDIP_BENCH_SHAPES=0,1,3
for rep in 1 2; do
for halo in 1 0; do for ws in 0 1; do
  echo "== rep$rep halo=$halo ws=$ws"; KDIP_CONV_HALO=$halo KDIP_CONV_WS=$ws timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3
done; done; done
for ws in 0 1; do echo "== halo dbg3 ws=$ws"; KDIP_CONV_WS=$ws KDIP_CONV_DBG=3 timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3; done
nvidia-smi --query-gpu=name,power.limit,clocks.sm,temperature.gpu,power.draw --format=csv
