#!/bin/bash
# isolates what slows the p2p transport as the slab count grows (debug switches, experiments only)
N=${1:-4}
run() { echo "== $1"; env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --transport $3 --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"])'; }
run "nccl" "X=1" nccl
run "p2p" "X=1" p2p
run "p2p nostore" "LBM_B200_DEBUG_NOSTORE=1" p2p
run "p2p nosync" "LBM_B200_DEBUG_NOSYNC=1" p2p
run "p2p nostore nosync" "LBM_B200_DEBUG_NOSTORE=1 LBM_B200_DEBUG_NOSYNC=1" p2p
