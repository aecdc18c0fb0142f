#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (3: Taylor-Green 256^3 x {15,19,27}; 4: channel D3Q27
1024x256x256) with the library's own CUDA-event timer.  Run under gpurun; prints one line per case."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lbm_b200 import capi  # noqa: E402
import scenario_reader as scenario  # noqa: E402
import cases  # noqa: E402

PEAK = 6539.2


def timed(d, steps, warm=10):
    d.step(warm)
    d.sync()
    d.step(steps)
    d.sync()
    return d.elapsed_ms() / steps


def report(name, Q, cells, ms):
    mlups = cells / (ms * 1e-3) / 1e6
    print("%-44s D3Q%d  %8.3f ms/step  %9.1f MLUPS  %7.1f GB/s algorithmic  %.3f of measured HBM" % (
        name, Q, ms, mlups, mlups * 2 * Q * 8 / 1e3, mlups * 2 * Q * 8 / 1e3 / PEAK), flush=True)


for Q in (15, 19, 27):
    n = 256
    rho, u, _ = cases.taylor_green(n, mode="3d")
    with capi.Domain(Q, n, n, n, 0.6) as d:
        d.set_boxes(cases.periodic_shell_boxes(n, n, n))
        d.init_equilibrium(rho, u)
        report("Taylor-Green 256^3 periodic", Q, n ** 3, timed(d, 200))
d, sc = scenario.domain_from_scenario(os.path.join(ROOT, "scenarios", "channel_d3q27.xml"), 27, 0.6)
report("channel 1024x256x256 inflow/outflow/no-slip", 27, sc["xl"] * sc["yl"] * sc["zl"], timed(d, 100))
d.close()
