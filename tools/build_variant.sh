#!/bin/bash
# Development aid: build a variant of libovvc_b200.so into build/variants/<name>.so (for A/B timing with tools/stage_bench.py --lib).
#   tools/build_variant.sh <name> [csrc dir] [extra nvcc flags, e.g. -DMCK_MIN_BLOCKS=6]
set -e
name=$1; src=${2:-openvvc_b200/csrc}; shift; shift || true
out=build/variants; mkdir -p $out/obj_$name
pids=()
for f in $src/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wno-unused-function --use_fast_math "$@" \
    -I$(pwd)/include -c $f -o $out/obj_$name/$(basename $f .cu).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -o $out/$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $out/obj_$name
echo $out/$name.so
