#!/bin/bash
# Run under gpurun: for every variants/liblbm_b200_*.so the cavity bench in FAST and in EXACT (bit-identical) arithmetic
# for D3Q19 / D3Q27, and the other configs (channel, Taylor-Green).  usage: tools/exact_variants.sh <tag>
mkdir -p gpurun_out
OUT=gpurun_out/exact_variants_${1:-r09}.txt
: > $OUT
one() {  # name so Q extra-flag label
  line=$(LBM_B200_LIB=$2 python bench.py --Q $3 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-exact $4 2>&1 | tail -1)
  echo "$1 cavity512 Q$3 $5 $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])' 2>/dev/null || echo "FAILED $line")" | tee -a $OUT
}
for so in variants/liblbm_b200_*.so; do
  name=$(basename $so .so); name=${name#liblbm_b200_}
  for Q in ${QS:-19 27}; do
    one $name $PWD/$so $Q "" fast
    one $name $PWD/$so $Q "--exact" exact
  done
  LBM_B200_LIB=$PWD/$so python tools/bench_configs.py ${WHAT:-big} 2>&1 | sed "s/^/$name /" | tee -a $OUT
done
