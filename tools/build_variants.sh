#!/bin/bash
# Build A/B variants of libhdgpu.so (compile-time knobs of kernel_fast6d.cu / kernel_rounds6d.cuh) into
# hyperdeal_b200/lib/variants/; select one at run time with HD_LIBHDGPU=<path>.
# usage: tools/build_variants.sh name "-DHD_R6_UNROLL_TASKS=1" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
python -c "from hyperdeal_b200 import build as B; B.build()"
mkdir -p hyperdeal_b200/lib/variants
OBJS=$(ls hyperdeal_b200/lib/*.o | grep -v kernel_fast6d.o)
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $flags -c hyperdeal_b200/csrc/kernel_fast6d.cu -o /tmp/kf_$name.o 2> hyperdeal_b200/lib/variants/$name.ptxas.log
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o hyperdeal_b200/lib/variants/libhdgpu_$name.so $OBJS /tmp/kf_$name.o
  echo "$name: $(grep -A2 'k_rounds_3d3v_k3ILb0ELb0' hyperdeal_b200/lib/variants/$name.ptxas.log | grep -E 'spill|Used' | tr '\n' ' ')"
done
