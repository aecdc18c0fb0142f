#!/bin/bash
# Run under gpurun: parity of the TMA-fed sweep, then its speed against the direct sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tma_gpu.py -x -q > gpurun_out/tma_tests.log 2>&1; echo "tma tests rc=$?"; tail -15 gpurun_out/tma_tests.log
OUT=gpurun_out/tma_${1:-r02}.txt
: > $OUT
for Q in ${2:-19 27 15}; do
for tma in 0 1; do
  line=$(LBM_B200_TMA=$tma timeout 300 python bench.py --Q $Q --size ${3:-512} --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-exact 2>&1 | tail -1)
  echo "Q$Q tma=$tma $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]["sm_mhz"])' 2>/dev/null || echo "FAILED $line")" | tee -a $OUT
done
done
