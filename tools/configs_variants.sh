#!/bin/bash
# Run under gpurun: the cavity bench and the other configs (channel, Taylor-Green) for every variants/liblbm_b200_*.so
mkdir -p gpurun_out
OUT=gpurun_out/configs_variants_${1:-r02}.txt
: > $OUT
for so in variants/liblbm_b200_*.so; do
  name=$(basename $so .so); name=${name#liblbm_b200_}
  for Q in 19 27; do
    line=$(LBM_B200_LIB=$PWD/$so python bench.py --Q $Q --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-exact 2>&1 | tail -1)
    echo "$name cavity512 Q$Q $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"])' 2>/dev/null || echo FAILED)" | tee -a $OUT
  done
  LBM_B200_LIB=$PWD/$so python tools/bench_configs.py big 2>&1 | sed "s/^/$name /" | tee -a $OUT
done
