#!/bin/bash
# Run under gpurun: times every variants/liblbm_b200_*.so on the cavity bench.
# usage: tools/sweep_variants.sh <tag> "<Q list>" [size]
mkdir -p gpurun_out
OUT=gpurun_out/variants_${1:-r01}.txt
: > $OUT
for Q in ${2:-19}; do
for so in variants/liblbm_b200_*.so; do
  name=$(basename $so .so); name=${name#liblbm_b200_}
  line=$(LBM_B200_LIB=$PWD/$so python bench.py --Q $Q --size ${3:-512} --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-exact 2>&1 | tail -1)
  echo "Q$Q $name $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])' 2>/dev/null || echo "FAILED $line")" | tee -a $OUT
done
done
