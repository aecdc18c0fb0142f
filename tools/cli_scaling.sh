#!/bin/bash
# Run under `gpurun --gpus 8`: weak scaling of the REFERENCE-FACING path -- the command-line driver src/main.cpp
# (config file + scenario XML -> lbm::Domain -> lbm_b200_step_group from one host thread), 512^3 cells per GPU.
# usage: tools/cli_scaling.sh <tag> "<N list>"
TAG=${1:-r02}
mkdir -p gpurun_out build
OUT=gpurun_out/cli_scaling_${TAG}.txt
: > $OUT
g++ -std=c++17 -O2 -fopenmp -ffp-contract=off -Iinclude/lbm -Iinclude/lbm/io src/main.cpp -o build/lbm \
    -Llbm_b200 -llbm_b200 -Wl,-rpath,$PWD/lbm_b200 || exit 1
for N in ${2:-1 2 4 8}; do
  Z=$((512 * N))
  sed "s/zl=\"512\"/zl=\"$Z\"/; s/Cavity512/Cavity512x$N/" scenarios/cavity512.xml > build/cavity512x$N.xml
  line=$(timeout 600 build/lbm configs/cavity512.cfg --scenario-file build/cavity512x$N.xml --gpus $N --timesteps 200 2>&1 | grep -E "interior cell updates|error|Error" | tail -1)
  echo "build/lbm cavity 512x512x$Z, $N GPU(s), 200 steps: $line" | tee -a $OUT
done
