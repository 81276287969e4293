#!/bin/bash
mkdir -p gpurun_out
export KDIP_BENCH_SHAPES=0,3 KDIP_HALO_PAIR=0
for dbg in 0 1 2 3; do echo "== single halo dbg=$dbg"; KDIP_CONV_DBG=$dbg timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -2; done
export KDIP_HALO_PAIR=1
for dbg in 0 1 2 3; do echo "== pair halo dbg=$dbg"; KDIP_CONV_DBG=$dbg timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -2; done
