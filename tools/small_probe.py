#!/usr/bin/env python
"""Small lattices under ncu: per-launch durations of the sweep at the reference's own sizes (40^3 scenarios, 64^3).
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sweep_kernel --csv python tools/small_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lbm_b200 import capi  # noqa: E402
import _oracle as O  # noqa: E402

for n in (40, 64, 128):
    with capi.Domain(19, n, n, n, 0.6) as d:
        d.set_graphs(0)
        d.set_boxes(O.cavity_boxes(n, n, n))
        d.step(12)
        d.sync()
