#!/bin/bash
# Developer tool: build kernel-geometry variants of the library for A/B timing on the GPU box.
# usage: tools/build_variants.sh name:"-DTFX_ROWBYTES=512 -DTFX_STAGES=2 -DTFX_WARPS=3" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/variants; mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I$ROOT/include -Xcompiler -fPIC,-fopenmp,-O3 --expt-relaxed-constexpr"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
  d=$OUT/obj_$name; mkdir -p $d
  for f in runtime sos_cascade sos_tile sos_packed sos_tma filterbank fir delay host_stream; do $NV $flags -c $ROOT/torchfx_b200/csrc/$f.cu -o $d/$f.o & done
  for f in sos_plan cpu_twin; do $NV $flags -x cu -c $ROOT/torchfx_b200/csrc/$f.cpp -o $d/$f.o & done
  wait
  $NV -shared -o $OUT/lib_$name.so $d/*.o -Xcompiler -fopenmp -lgomp -cudart static
  echo built $OUT/lib_$name.so
  ) &
done
wait
