#!/bin/bash
# compute-sanitizer memcheck + racecheck on small lattices (run under gpurun); summary -> gpurun_out/sanitizer_<tag>.txt
TAG=${1:-r01}
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_${TAG}.txt
: > $OUT
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool : parity tests (cavity, channel+block, weird, periodic) + multi-slab + x-face guesses" >> $OUT
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_multislab_gpu.py tests/test_round2_gpu.py -m gpu -x -q \
      -k "((cavity16 or channel_block or weird or one_step or cavity_slabs or obstacle or periodic_extension) and not 15 and not fast) or xface" > gpurun_out/sanitizer_${tool}.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}.log | tail -4 >> $OUT
done
cat $OUT
