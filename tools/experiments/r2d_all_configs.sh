#!/bin/bash
# bench lines of every BASELINE config with the current code (short runs), logs -> gpurun_out/
mkdir -p gpurun_out
for c in 4 2 0 3; do
  case $c in 4) k=3; w=3;; 2) k=2; w=3;; 0) k=3; w=3;; 3) k=1; w=3;; esac
  timeout 1500 python bench.py --config $c --steps $k --warmup $w > gpurun_out/r2d_bench_cfg$c.json 2> gpurun_out/r2d_bench_cfg$c.err
  echo "cfg$c rc=$? $(cut -c1-160 gpurun_out/r2d_bench_cfg$c.json)"
  grep -a "bench\]" gpurun_out/r2d_bench_cfg$c.err | tail -2
done
