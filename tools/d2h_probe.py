#!/usr/bin/env python
"""Run under torchrun with N ranks (gpurun --gpus N): what can the HOST absorb?  Every rank copies a 2 GiB device
buffer to page-locked host memory several times, all ranks at once; prints per-rank and aggregate GB/s for
 (a) torch pinned memory, (b) lbm_b200_host_alloc (first-touch on the NUMA node next to the GPU),
 (c) cudaHostAlloc(cudaHostAllocWriteCombined), and the topology the placement code can see."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from lbm_b200 import capi

N = 1 << 28          # doubles = 2 GiB
src = torch.zeros(N, dtype=torch.float64, device="cuda")


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(name, dst_ptr_or_tensor):
    reps = 4
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst_ptr_or_tensor.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    mine = reps * N * 8 / (time.perf_counter() - t0) / 1e9
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([mine], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        print("%-44s aggregate %7.1f GB/s (sum of per-rank rates %7.1f), %d ranks" % (name, world * reps * N * 8 / dt / 1e9, float(t.item()), world), flush=True)


if rank == 0:
    os.system("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'; ls /sys/devices/system/node/ | tr '\\n' ' '; echo; "
              "for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/vendor)\" = 0x10de ]; then echo $(basename $d) numa_node=$(cat $d/numa_node); fi; done | head -10; "
              "nvidia-smi topo -m | head -14")
barrier()
run("torch pinned memory", torch.empty(N, dtype=torch.float64, pin_memory=True))
hb = capi.HostBuffer(N, device=local)
capi.lib.lbm_b200_bind_host_thread(local)
run("lbm_b200_host_alloc (NUMA first touch)", torch.from_numpy(hb.array))
try:
    from cuda import cudart
    err, p = cudart.cudaHostAlloc(N * 8, cudart.cudaHostAllocWriteCombined)
    if int(err) == 0:
        import ctypes
        ctypes.memset(p, 0, N * 8)
        import numpy as np
        arr = np.ctypeslib.as_array((ctypes.c_double * N).from_address(p))
        run("cudaHostAlloc write-combined", torch.from_numpy(arr))
except Exception as ex:
    if rank == 0:
        print("write-combined probe failed:", ex)
if world > 1:
    dist.destroy_process_group()
