#!/usr/bin/env python
"""Two z-slabs on two GPUs driven from ONE process (LocalSlabStack = what the C++ Domain does for gpus = 2), so
that a single ncu session sees the edge-plane launches of the fused sweep + exchange: every step each slab
launches sweep_kernel over its two edge planes (grid z = 2, the launch that stores the leaving populations into
the neighbour's ghost plane over NVLink) and then over the interior.  usage: edge_probe.py [n] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lbm_b200 import capi  # noqa: E402
from lbm_b200.slabs import LocalSlabStack  # noqa: E402
import _oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
assert capi.device_count() >= 2, "needs two GPUs"
st = LocalSlabStack(19, n, n, 2 * n, 0.6, O.cavity_boxes(n, n, 2 * n), 2, devices=[0, 1])
for _ in range(steps):       # one step at a time: a profiler serialises kernels, so a device-side wait for a
    st.step(1)               # neighbour's NEXT launch could never be satisfied
    st.sync()
print("edge_probe: %d steps of two %d^3 slabs, %d launches on slab 0" % (steps, n, st.slabs[0].launch_count()))
st.close()
