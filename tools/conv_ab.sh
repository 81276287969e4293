#!/bin/bash
# A/B of the conv_gemm variants on the FFHQ UNet (B=32): per-launch device times via ncu, summarised by tests/tools/conv_table.py
mkdir -p gpurun_out
KDIP_CONV_DEBUG=1 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/ab_default_time.log 2> gpurun_out/ab_default_debug.log
KDIP_CONV_PAIR=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/ab_nopair_time.log 2>&1
KDIP_CONV_TMAEPI=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/ab_legacy_time.log 2>&1
for cfg in default nopair legacy; do
  case $cfg in
    default) export -n KDIP_CONV_PAIR KDIP_CONV_TMAEPI;;
    nopair) export KDIP_CONV_PAIR=0;;
    legacy) export -n KDIP_CONV_PAIR; export KDIP_CONV_TMAEPI=0;;
  esac
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$cfg.csv \
    python tools/time_unet.py 32 1 > gpurun_out/ab_${cfg}_ncu.log 2>&1
done
tail -1 gpurun_out/ab_*_time.log
