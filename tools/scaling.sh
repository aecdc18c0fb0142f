#!/bin/bash
# weak-scaling series the way the driver launches it; run under `gpurun --gpus 8`
mkdir -p gpurun_out
OUT=gpurun_out/scaling_${1:-r01}.txt
: > $OUT
for T in p2p nccl; do
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then CMD="python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e";
  else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 10 --transport $T --no-e2e"; fi
  line=$($CMD 2>&1 | tail -1)
  echo "$T N=$N $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "MLUPS", d["ms_per_step"], "ms/step", d["gpu_launches"], "launches", d["clocks"]["sm_mhz"], "MHz", d["clocks"]["reasons"])' 2>/dev/null || echo "FAILED: $line" | cut -c1-300)" | tee -a $OUT
done
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 2>&1 | tail -1 > gpurun_out/BENCH_local_8gpu.json
cut -c1-300 gpurun_out/BENCH_local_8gpu.json
