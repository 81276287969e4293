#!/bin/bash
mkdir -p gpurun_out
export KDIP_BENCH_SHAPES=0,1,3
echo "== halo default"; timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3
echo "== halo ws"; KDIP_CONV_WS=1 timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3
echo "== halo ws dbg3"; KDIP_CONV_WS=1 KDIP_CONV_DBG=3 timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3
echo "== v1 mt2 ws"; KDIP_CONV_HALO=0 KDIP_CONV_WS=1 timeout 120 python tools/bench_conv.py 32 10 2>&1 | tail -3
echo "== tests with ws"; KDIP_CONV_WS=1 timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
