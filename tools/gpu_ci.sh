#!/bin/bash
# Run the GPU parity tests file by file (a trapped kernel poisons one process, not the rest); logs -> gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-tests/test_elementwise_gpu.py tests/test_conv_gemm_gpu.py tests/test_layers_gpu.py tests/test_unet_gpu.py tests/test_operators_gpu.py tests/test_guidance_gpu.py tests/test_frontend_gpu.py}; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -s -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "== $name exit $? ==" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|error" "gpurun_out/$name.log" | tail -n 2
done
