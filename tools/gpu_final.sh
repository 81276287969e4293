#!/bin/bash
# End-of-round GPU pass: the driver's own test command, the ncu launch list of one UNet evaluation, the bench line, the
# reference arm, smoke(), and the ImageNet-architecture timing.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/ci.log 2>&1; tail -2 gpurun_out/ci.log
KDIP_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv python tools/time_unet.py 32 1 > gpurun_out/final_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python tools/time_imagenet.py 32 3 > gpurun_out/time_imagenet_final.log 2>&1; tail -2 gpurun_out/time_imagenet_final.log
timeout 120 python tools/sustain_unet.py 32 9 > gpurun_out/time_unet_final.log 2>&1; tail -1 gpurun_out/time_unet_final.log
timeout 200 python tools/time_fft.py 32 20 > gpurun_out/time_fft_final.log 2>&1; cat gpurun_out/time_fft_final.log
