#!/bin/bash
export KDIP_BENCH_SHAPES=9,0,1
for dbg in 0 2 1 3; do echo "== dbg=$dbg"; KDIP_CONV_DBG=$dbg timeout 120 python tools/bench_conv.py 32 20 2>&1 | tail -3; done
