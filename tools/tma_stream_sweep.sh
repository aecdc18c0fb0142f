#!/bin/bash
# Run under gpurun: TMA streaming rate (HBM -> shared memory) against box shape, ops per stage and ring depth.
mkdir -p gpurun_out
OUT=gpurun_out/tma_stream_${1:-r02}.txt
: > $OUT
S=tools/selftest/tma_stream
for args in "128 2 19 5" "128 2 19 5 14" "128 2 27 4" "128 2 1 64" "128 2 4 24" "128 4 19 2" "128 4 8 6" "128 8 8 3" "128 8 4 6" "128 8 1 24" \
            "128 16 1 12" "128 32 1 6" "256 1 19 5" "256 2 8 6" "256 4 4 6" "256 8 2 6" "256 16 1 6" "64 4 19 5" "32 8 19 5" \
            "128 2 19 5 16 1" "128 2 19 5 16 2" "256 8 2 6 16 2" "128 2 19 2" "128 2 19 3"; do
  timeout 60 $S $args | tee -a $OUT
done
