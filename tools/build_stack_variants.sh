#!/bin/bash
# Developer tool: A/B variants of ONE translation unit (default bank_stack.cu), linked with the
# objects of the normal build.  usage: tools/build_stack_variants.sh name:"-DTFX_BS_CH=128" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
UNIT=${UNIT:-bank_stack}
OUT=$ROOT/build/variants; mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I$ROOT/include -Xcompiler -fPIC,-fopenmp,-O3 --expt-relaxed-constexpr"
make -C $ROOT/torchfx_b200/csrc -j8 >/dev/null
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
  $NV $flags -c $ROOT/torchfx_b200/csrc/$UNIT.cu -o $OUT/${UNIT}_$name.o
  others=$(ls $ROOT/build/obj/*.o | grep -v "/$UNIT.o")
  $NV -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/lib_$name.so $OUT/${UNIT}_$name.o $others -Xcompiler -fopenmp -lgomp -cudart static
  echo built $OUT/lib_$name.so
  ) &
done
wait
