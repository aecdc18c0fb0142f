#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (3: Taylor-Green 256^3 x {15,19,27}; 4: channel D3Q27
1024x256x256), of the reference's own small lattices (40^3 scenarios, 64^3 = config 1; with and without CUDA
graphs) and of pipe geometries (the reference's 250x54x54 pipe.vtk and a 512x256x256 pipe with the same 44 %
of solid cells; pulls issued speculatively or only after the bit map was checked) with the library's own
CUDA-event timer.  Run under gpurun; prints one line per case.
usage: tools/bench_configs.py [big] [channel] [small] [pipe]   (default: big small pipe)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lbm_b200 import capi  # noqa: E402
import scenario_reader as scenario  # noqa: E402
import cases  # noqa: E402

import json
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
WHAT = set(sys.argv[1:]) or {"big", "small", "pipe"}


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


def big():
    for Q in (15, 19, 27):
        n = 256
        rho, u, _ = cases.taylor_green(n, mode="3d")
        with capi.Domain(Q, n, n, n, 0.6) as d:
            d.set_boxes(cases.periodic_shell_boxes(n, n, n))
            d.init_equilibrium(rho, u)
            report("Taylor-Green 256^3 periodic", Q, n ** 3, timed(d, 200))
    channel()


def channel():
    d, sc = scenario.domain_from_scenario(os.path.join(ROOT, "scenarios", "channel_d3q27.xml"), 27, 0.6)
    report("channel 1024x256x256 inflow/outflow/no-slip", 27, sc["xl"] * sc["yl"] * sc["zl"], timed(d, 100))
    d.close()


def small():
    """launch-bound lattices: 16 steps per CUDA-graph launch against one launch per step"""
    import _oracle as O
    for n in (40, 64, 128):
        for graphs in (0, 1):
            with capi.Domain(19, n, n, n, 0.6) as d:
                d.set_graphs(graphs)
                d.set_boxes(O.cavity_boxes(n, n, n))
                report("cavity %d^3 %s" % (n, "CUDA graphs of 16 steps" if graphs else "one launch per step"), 19, n ** 3,
                       timed(d, 1600, warm=33))


def pipe_mask(xl, yl, zl):
    zz, yy = np.meshgrid(np.arange(zl), np.arange(yl), indexing="ij")
    disc = ((yy - (yl - 1) / 2) ** 2 / (yl / 2 - 1) ** 2 + (zz - (zl - 1) / 2) ** 2 / (zl / 2 - 1) ** 2) < 0.72   # ~56 % fluid
    return np.ascontiguousarray(np.broadcast_to(disc[:, :, None], (zl, yl, xl))).astype(np.uint8)


def pipe():
    import _oracle as O
    from test_round2_gpu import read_legacy_vtk_mask
    (xl, yl, zl), mask = read_legacy_vtk_mask(os.path.join(ROOT, "tests", "golden", "pipe.vtk"))
    geos = [("reference pipe.vtk 250x54x54", xl, yl, zl, mask), ("pipe 512x256x256", 512, 256, 256, pipe_mask(512, 256, 256))]
    for name, xl, yl, zl, mask in geos:
        fluid = int(mask.sum())
        for checked in (0, 1):
            with capi.Domain(19, xl, yl, zl, 0.6) as d:
                d.set_sweep_engine(tma=0, checked=checked)
                d.set_fluid_mask(mask)
                d.set_boxes(O.channel_boxes(xl, yl, zl))
                d.tag_null_cells()
                ms = timed(d, 400 if xl < 300 else 100, warm=33)
            mlups = fluid / (ms * 1e-3) / 1e6
            print("%-30s %4.1f %% solid, %-28s D3Q19  %8.3f ms/step  %9.1f MLUPS (fluid-cell updates)  %.3f of measured HBM" % (
                name, 100.0 * (1 - fluid / mask.size), "bit checked before the pulls" if checked else "speculative pulls",
                ms, mlups, mlups * 2 * 19 * 8 / 1e3 / PEAK), flush=True)


if "big" in WHAT:
    big()
if "channel" in WHAT:
    channel()
if "small" in WHAT:
    small()
if "pipe" in WHAT:
    pipe()
