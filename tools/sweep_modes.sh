#!/bin/bash
# Run under gpurun: times the interior-launch modes (LBM_B200_SWEEP_MODE 0 inline / 1 bulk+wall / 2 checked+wall).
# usage: tools/sweep_modes.sh <tag> "<Q list>" [size]
mkdir -p gpurun_out
OUT=gpurun_out/modes_${1:-r02}.txt
: > $OUT
for Q in ${2:-19}; do
for mode in 0 1 2; do
  line=$(LBM_B200_SWEEP_MODE=$mode python bench.py --Q $Q --size ${3:-512} --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-exact 2>&1 | tail -1)
  echo "Q$Q mode=$mode $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]["sm_mhz"])' 2>/dev/null || echo "FAILED $line")" | tee -a $OUT
done
done
