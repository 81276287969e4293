#!/bin/bash
# Build libkdip.so (sm_100a only) in-tree: k-diffusion-inverse-problems_b200/kdip/libkdip.so
# nvcc cross-compiles without a GPU.  Objects are cached under csrc/build/ and rebuilt when the source is newer.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../kdip/libkdip.so"
mkdir -p "$HERE/build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${KDIP_NVCC_EXTRA}"
OBJS=""
pids=""
for src in "$HERE"/*.cu; do
  obj="$HERE/build/$(basename "${src%.cu}").o"
  OBJS="$OBJS $obj"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/kdip_common.cuh" -nt "$obj" ] || [ "$HERE/../../include/kdip.h" -nt "$obj" ] \
     || [ "$HERE/unet_kernels.cuh" -nt "$obj" ] || [ "$HERE/fft.cuh" -nt "$obj" ] || [ "$HERE/attention_tc.cuh" -nt "$obj" ] || [ "$0" -nt "$obj" ]; then
    # a source may ask for extra flags with a line "// NVCC_FLAGS: ..." (e.g. -fmad=false where the reference's
    # separately-rounded mul/add order must be reproduced)
    extra=$(grep -m1 -oP '(?<=^// NVCC_FLAGS: ).*' "$src" || true)
    $NVCC $FLAGS $extra -c "$src" -o "$obj" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
