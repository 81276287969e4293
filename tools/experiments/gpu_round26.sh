#!/bin/bash
export KDIP_BENCH_SHAPES=9,0
for st in 8 3 2; do echo "== stages<=$st"; KDIP_CONV_STAGES=$st KDIP_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py 32 30 2>&1 | grep -E "TF/s|stages" | grep -v "pair clusters"; done
