#!/bin/bash
# Builds tuning variants of the library (same sources, different -D knobs) into variants/.
# usage: tools/build_variants.sh name1:"-DFLAG ..." name2:"..."
set -e
mkdir -p variants
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr -shared"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( $NV $flags -Xptxas -v -o variants/liblbm_b200_${name}.so lbm_b200/csrc/engine.cu -lcudart 2> variants/${name}.ptxas.log; \
    grep -A2 "sweep_kernelILi19ELb0" variants/${name}.ptxas.log | grep -E "Used|spill" | tr '\n' ' ' | sed "s/^/${name}: /"; echo ) &
done
wait
