#!/bin/bash
export KDIP_BENCH_SHAPES=10,7,11,12,13,14,4,5
echo "== default"; KDIP_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py 32 30 2>&1 | grep -E "TF/s|BN=" | grep -v clusters | sed -E 's/smem=[0-9]+//'
for bn in 64 128 256; do for pair in 0 1; do for mt in 1 2; do
  echo "== BN=$bn PAIR=$pair MT=$mt"; KDIP_CONV_BN=$bn KDIP_CONV_PAIR=$pair KDIP_CONV_MT=$mt timeout 120 python tools/bench_conv.py 32 30 2>&1 | grep "TF/s"
done; done; done
