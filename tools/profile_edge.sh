#!/bin/bash
# Run under `gpurun --gpus 2`: NVLink / peer-memory counters of the fused sweep + exchange (edge-plane launches).
# One process drives both slabs (tools/edge_probe.py), so this is a single-process ncu session.
TAG=${1:-r02}
mkdir -p gpurun_out
M=gpu__time_duration.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvlrx__bytes.sum,syslts__t_requests_aperture_peer.sum,syslts__t_requests_aperture_peer_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:sweep_kernel -s 8 -c 8 --csv \
    --log-file gpurun_out/edge_sweep_${TAG}.csv python tools/edge_probe.py 512 6 > gpurun_out/edge_sweep_${TAG}.log 2>&1
tail -2 gpurun_out/edge_sweep_${TAG}.log
