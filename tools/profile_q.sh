#!/bin/bash
# one `ncu --set full` capture of sweep_kernel for a given lattice/size (run under gpurun)
Q=${1:-19}; SIZE=${2:-512}; TAG=${3:-r01}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 \
    -f -o gpurun_out/sweep_q${Q}_${TAG} \
    python bench.py --Q $Q --size ${SIZE} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_q${Q}_${TAG}.log 2>&1
