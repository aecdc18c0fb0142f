#!/bin/bash
# Run under gpurun (one GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv            every kernel launch of a short bench run with its device time
#   sweep_<tag>.raw.csv (+ .source.csv)   one `--set full` capture of sweep_kernel<19> (source-level, -lineinfo)
#   sweep_q27_<tag>.ncu-rep, sweep_q15_<tag>.ncu-rep   the same for the other lattices
#   sweep_q27_channel_<tag>.ncu-rep                     D3Q27 on the 1024x256x256 channel (config 4)
# afterwards, here:  tools/refresh_traffic.py gpurun_out/sweep_<tag>.raw.csv gpurun_out/sweep_q27_<tag>.raw.csv ...
#                    tools/ncu_summary.py full gpurun_out/sweep_<tag>.raw.csv
# usage: tools/profile_gpu.sh <tag> [size-for-full-capture]
TAG=${1:-r02}
SIZE=${2:-512}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 \
    -f -o gpurun_out/sweep_${TAG} \
    python bench.py --size ${SIZE} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-exact > gpurun_out/sweep_${TAG}.log 2>&1
for Q in 27 15; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 \
    -f -o gpurun_out/sweep_q${Q}_${TAG} \
    python bench.py --Q $Q --size ${SIZE} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-exact > gpurun_out/sweep_q${Q}_${TAG}.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 12 -c 1 \
    -f -o gpurun_out/sweep_q27_channel_${TAG} \
    python tools/bench_configs.py channel > gpurun_out/sweep_q27_channel_${TAG}.log 2>&1
# gpurun copies back at most 64 MiB and a capture is 45 MB: keep the raw metric pages (and the source page of the
# D3Q19 capture) as CSV, drop the reports
for rep in gpurun_out/sweep_${TAG} gpurun_out/sweep_q27_${TAG} gpurun_out/sweep_q15_${TAG} gpurun_out/sweep_q27_channel_${TAG}; do
  [ -f $rep.ncu-rep ] || continue
  ncu -i $rep.ncu-rep --page raw --csv > $rep.raw.csv 2>/dev/null
done
ncu -i gpurun_out/sweep_${TAG}.ncu-rep --page source --csv > gpurun_out/sweep_${TAG}.source.csv 2>/dev/null
rm -f gpurun_out/*_${TAG}.ncu-rep
ls -la gpurun_out
