#!/bin/bash
# Run under gpurun (one GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv   every kernel launch of a short bench run with its device time
#   sweep_<tag>.ncu-rep  one `--set full` capture of sweep_kernel (source-level, -lineinfo)
# usage: tools/profile_gpu.sh <tag> [size-for-full-capture]
TAG=${1:-r01}
SIZE=${2:-512}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 \
    -f -o gpurun_out/sweep_${TAG} \
    python bench.py --size ${SIZE} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_${TAG}.log 2>&1
ls -la gpurun_out
