"""Generates tests/golden/*.npz from the reference's OWN code (oracle/_ref/libref_lbm.so, built from
/root/reference by `make -C oracle ref`).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Every fixture stores the inputs (lattice, sizes, tau, boxes, steps, optional initial state / mask) and
the reference's outputs (all populations, density, velocity, handler kinds).  The reference ships no
golden vectors of its own; these files pin the oracle and the CUDA path to the
reference's executable behaviour on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402


def boxes_to_array(boxes):
    return np.array([[k, *v, rho, *ext] for (k, v, rho, ext) in boxes], dtype=np.float64).reshape(-1, 11)


def fixtures():
    rng = np.random.default_rng(2026)
    for Q in (15, 19, 27):
        yield "cavity_q%d" % Q, dict(Q=Q, xl=6, yl=5, zl=7, boxes=O.cavity_boxes(6, 5, 7), steps=12)
        ch = O.channel_boxes(10, 5, 4)
        ch.insert(2, (O.NOSLIP, (0.0, 0.0, 0.0), 1.0, (3, 5, 2, 3, 0, 2)))
        yield "channel_q%d" % Q, dict(Q=Q, xl=10, yl=5, zl=4, boxes=ch, steps=15)
        sh = O.face_boxes(4, 4, 6, [("z0", O.PRESSURE, None, 1.005), ("zmax", O.OUTFLOW), ("x0", O.FREESLIP),
                                    ("xmax", O.FREESLIP), ("y0", O.NOSLIP), ("ymax", O.NOSLIP)])
        yield "shear_q%d" % Q, dict(Q=Q, xl=4, yl=4, zl=6, boxes=sh, steps=14)
        f0 = rng.random((6 * 6 * 6, Q)) * 0.1 + 0.05
        yield "periodic_q%d" % Q, dict(Q=Q, xl=4, yl=4, zl=4, boxes=[], steps=9, f_init=f0, periodic=True)
    mask = np.ones((4, 5, 8), dtype=np.uint8)
    mask[1:3, 1:4, 2:5] = 0
    yield "mask_q19", dict(Q=19, xl=8, yl=5, zl=4, boxes=O.channel_boxes(8, 5, 4), steps=11, fluid_mask=mask)


def main():
    ref = O.ref()
    if ref is None:
        raise SystemExit("needs /root/reference (or a prebuilt oracle/_ref/libref_lbm.so)")
    for name, c in fixtures():
        kw = {k: c[k] for k in ("f_init", "fluid_mask", "periodic") if k in c}
        out = ref.run(c["Q"], c["xl"], c["yl"], c["zl"], 0.6, c["boxes"], c["steps"], **kw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), Q=c["Q"], xl=c["xl"], yl=c["yl"], zl=c["zl"], tau=0.6, steps=c["steps"],
            boxes=boxes_to_array(c["boxes"]), periodic=int(bool(c.get("periodic", False))),
            f_init=c.get("f_init", np.zeros(0)), fluid_mask=c.get("fluid_mask", np.zeros(0, dtype=np.uint8)),
            f=out["f"], rho=out["rho"], u=out["u"], kind=out["kind"])
        print("wrote", name, out["f"].shape)
    # known-answer scalars of SURVEY.md 8c (cavity.xml order, tau 0.6, lid 0.05): sequential sum of the
    # interior densities and max |u_x|
    kat = {}
    for Q, n, steps in [(19, 32, 100), (15, 32, 100), (27, 32, 100), (19, 32, 200), (19, 64, 20)]:
        o = ref.run(Q, n, n, n, 0.6, O.cavity_boxes(n, n, n), steps, want=("rho", "u"))
        kat["q%d_n%d_s%d" % (Q, n, steps)] = np.array([np.cumsum(o["rho"].reshape(-1))[-1], np.abs(o["u"][..., 0]).max()])
    np.savez(os.path.join(HERE, "kat_survey.npz"), **kat)
    print("wrote kat_survey")


if __name__ == "__main__":
    main()
