#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_conv.log 2>&1; echo "conv+unet tests exit $?"
timeout 300 python tools/time_unet.py 32 3 > gpurun_out/pairfix_time.log 2>&1
KDIP_CONV_PAIR=0 timeout 300 python tools/time_unet.py 32 3 > gpurun_out/nopair_time.log 2>&1
tail -n 1 gpurun_out/pairfix_time.log gpurun_out/nopair_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_pairfix.csv python tools/time_unet.py 32 1 > gpurun_out/pairfix_ncu.log 2>&1
for k in gn_bwd_reduce gn_bwd_apply gn_apply; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 2 -f -o gpurun_out/full_$k python tools/time_unet.py 8 1 > gpurun_out/full_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 200 -c 3 -f -o gpurun_out/full_conv python tools/time_unet.py 8 1 > gpurun_out/full_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
