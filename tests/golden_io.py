"""Loads tests/golden/*.npz fixtures (made by tests/golden/make_golden.py from the reference's own code)."""
import glob
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")) if "kat_" not in p)


def load(name):
    d = np.load(os.path.join(HERE, name + ".npz"))
    boxes = [(int(r[0]), (r[1], r[2], r[3]), float(r[4]), tuple(int(v) for v in r[5:11])) for r in d["boxes"]]
    c = dict(Q=int(d["Q"]), xl=int(d["xl"]), yl=int(d["yl"]), zl=int(d["zl"]), tau=float(d["tau"]), steps=int(d["steps"]),
             boxes=boxes, periodic=bool(int(d["periodic"])), f=d["f"], rho=d["rho"], u=d["u"], kind=d["kind"])
    c["f_init"] = d["f_init"] if d["f_init"].size else None
    c["fluid_mask"] = d["fluid_mask"] if d["fluid_mask"].size else None
    return c


def kat():
    return dict(np.load(os.path.join(HERE, "kat_survey.npz")))
