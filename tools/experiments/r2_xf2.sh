#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x -k "fused_groupnorm_apply" -p no:cacheprovider 2>&1 | tail -n 3
export KDIP_BENCH_SHAPES=0,9,1,3,2
echo "== unfused"; timeout 200 python tools/bench_conv.py 32 20
echo "== fused"; KDIP_BENCH_XF=1 timeout 200 python tools/bench_conv.py 32 20
echo "== fused, handshake only (dbg 4)"; KDIP_CONV_DBG=4 KDIP_BENCH_XF=1 timeout 200 python tools/bench_conv.py 32 20
echo "== fused, copy only (dbg 8)"; KDIP_CONV_DBG=8 KDIP_BENCH_XF=1 timeout 200 python tools/bench_conv.py 32 20
