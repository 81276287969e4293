#!/bin/bash
# register budgets of the 640-thread conv variant: control / transform / epilogue
for cfg in "72 64 136" "80 64 136" "96 64 128"; do
  set -- $cfg
  touch k-diffusion-inverse-problems_b200/csrc/conv_gemm.cu
  KDIP_NVCC_EXTRA="-DKDIP_XF_REG_CTL=$1 -DKDIP_XF_REG_XF=$2 -DKDIP_XF_REG_EPI=$3" bash k-diffusion-inverse-problems_b200/csrc/build.sh > /tmp/build.log 2>&1 || { echo "build failed $cfg"; tail -3 /tmp/build.log; continue; }
  cuobjdump --dump-resource-usage k-diffusion-inverse-problems_b200/csrc/build/conv_gemm.o 2>&1 | grep -A1 "ILb1ELb1ELb1E" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' '
  echo "ctl/xf/epi = $cfg"
  for d in 0 4; do KDIP_BENCH_XF=1 KDIP_CONV_DBG=$d timeout 100 python tools/bench_conv_gn.py 32 10 256 128 128 0 2>/dev/null | grep "gn_reduce=0" | sed "s/^/   dbg=$d /"; done
done
