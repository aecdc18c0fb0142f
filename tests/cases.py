"""Scenario builders shared by the CPU and GPU tests (inputs fed to BOTH the
oracle and the CUDA path).  Box lists use the format of tests/_oracle.py."""
import numpy as np

import _oracle as O


def cavity(n, lid=(0.05, 0.0, 0.0)):
    return dict(xl=n, yl=n, zl=n, boxes=O.cavity_boxes(n, n, n, lid))


def channel(xl, yl, zl, block=None, u_in=(0.03, 0.0, 0.0)):
    b = O.channel_boxes(xl, yl, zl, u_in)
    if block is not None:
        b.insert(2, (O.NOSLIP, (0.0, 0.0, 0.0), 1.0, tuple(block)))
    return dict(xl=xl, yl=yl, zl=zl, boxes=b)


def shearflow():
    """build/scenarios/shearflow.xml"""
    return dict(xl=8, yl=8, zl=20, boxes=O.face_boxes(8, 8, 20, [
        ("z0", O.PRESSURE, None, 1.005), ("zmax", O.OUTFLOW), ("x0", O.FREESLIP), ("xmax", O.FREESLIP),
        ("y0", O.NOSLIP), ("ymax", O.NOSLIP)]))


def step_flow():
    """build/scenarios/step.xml"""
    return dict(xl=30, yl=30, zl=50, boxes=O.face_boxes(30, 30, 50, [
        ((17, 31, 0, 31, 0, 0), O.INFLOW, (0.0, 0.0, 0.05)), ((0, 16, 0, 31, 0, 25), O.NOSLIP),
        ("zmax", O.OUTFLOW), ("y0", O.NOSLIP), ("ymax", O.NOSLIP), ("x0", O.NOSLIP), ("xmax", O.NOSLIP)]))


def masked_pipe(xl=24, yl=10, zl=10):
    zz, yy, xx = np.meshgrid(np.arange(zl), np.arange(yl), np.arange(xl), indexing="ij")
    mask = (((yy - (yl - 1) / 2) ** 2 + (zz - (zl - 1) / 2) ** 2) < (min(yl, zl) / 2 - 1) ** 2).astype(np.uint8)
    return dict(xl=xl, yl=yl, zl=zl, boxes=O.channel_boxes(xl, yl, zl), fluid_mask=mask)


def weird(Q, seed=0):
    """uncovered ghost shell, ParallelBoundary face, interior free-slip block, oblique lid, random state"""
    xl, yl, zl = 10, 9, 8
    rng = np.random.default_rng(seed)
    boxes = O.face_boxes(xl, yl, zl, [("z0", O.NOSLIP), ("x0", O.PARALLEL), ((4, 6, 3, 5, 2, 4), O.FREESLIP),
                                      ("ymax", O.MOVINGWALL, (0.01, 0.02, -0.03))])
    f0 = rng.random(((xl + 2) * (yl + 2) * (zl + 2), Q)) * 0.1 + 0.05
    return dict(xl=xl, yl=yl, zl=zl, boxes=boxes, f_init=f0)


def periodic_random(Q, n=8, seed=1):
    rng = np.random.default_rng(seed)
    f0 = rng.random(((n + 2) ** 3, Q)) * 0.1 + 0.05
    return dict(xl=n, yl=n, zl=n, boxes=[], f_init=f0, periodic=True)


def periodic_shell_boxes(xl, yl, zl):
    """the PERIODIC extension of the CUDA path: every ghost-shell cell mirrors its wrapped image"""
    return O.face_boxes(xl, yl, zl, [(e, O.PERIODIC) for e in ("z0", "zmax", "x0", "xmax", "y0", "ymax")])


def interior_index(xl, yl, zl):
    """flat Domain::idx of the interior cells in z,y,x order"""
    z, y, x = np.meshgrid(np.arange(1, zl + 1), np.arange(1, yl + 1), np.arange(1, xl + 1), indexing="ij")
    return (x + (xl + 2) * y + (xl + 2) * (yl + 2) * z).reshape(-1)


def taylor_green(n, U0=0.01, mode="xy"):
    """Taylor-Green initial state on all (n+2)^3 cells (cell centres at i-1/2).

    Returns (rho, u, c) with c the decay constant: kinetic energy ~ exp(-c * nu * k^2 * t), k = 2 pi / n.
      xy/yz/xz  2-D vortex in a coordinate plane, |k|^2 = 2 k^2          -> c = 4
      diag      the 2-D vortex rotated by 45 degrees in the x-y plane, |k|^2 = 4 k^2 -> c = 8
      3d        u = U0 sin kx cos ky cos kz, v = -U0 cos kx sin ky cos kz, w = 0, |k|^2 = 3 k^2 -> c = 6
                (exact only as the initial decay rate; Re = U0/(k nu) is kept small)
    """
    cs2 = 0.57735026919 ** 2
    k = 2 * np.pi / n
    idx = np.arange(n + 2) - 0.5
    Z, Y, X = np.meshgrid(idx, idx, idx, indexing="ij")
    u = np.zeros((n + 2, n + 2, n + 2, 3))
    rho = np.ones_like(X)
    if mode == "3d":
        u[..., 0] = U0 * np.sin(k * X) * np.cos(k * Y) * np.cos(k * Z)
        u[..., 1] = -U0 * np.cos(k * X) * np.sin(k * Y) * np.cos(k * Z)
        c = 6.0
    elif mode == "diag":
        a, b = k * (X + Y), k * (X - Y)
        ua = U0 * np.sin(a) * np.cos(b)
        ub = -U0 * np.cos(a) * np.sin(b)
        u[..., 0] = (ua + ub) / np.sqrt(2.0)
        u[..., 1] = (ua - ub) / np.sqrt(2.0)
        rho = 1 - U0 ** 2 / (4 * cs2) * (np.cos(2 * a) + np.cos(2 * b))
        c = 8.0
    else:
        A, B = {"xy": (X, Y), "yz": (Y, Z), "xz": (X, Z)}[mode]
        ia, ib = {"xy": (0, 1), "yz": (1, 2), "xz": (0, 2)}[mode]
        u[..., ia] = U0 * np.sin(k * A) * np.cos(k * B)
        u[..., ib] = -U0 * np.cos(k * A) * np.sin(k * B)
        rho = 1 - U0 ** 2 / (4 * cs2) * (np.cos(2 * k * A) + np.cos(2 * k * B))
        c = 4.0
    return rho, u, c


def duct_profile(ny, nz, terms=60):
    """Fully developed laminar velocity in a rectangular duct (walls half-way between the first/last fluid
    cell and the wall cell: widths ny, nz lattice units, cell centres at j-1/2), normalised to mean 1."""
    y = np.arange(1, ny + 1) - 0.5 - ny / 2.0      # centred coordinates
    z = np.arange(1, nz + 1) - 0.5
    a, b = ny / 2.0, float(nz)                      # y in [-a, a], z in [0, b]
    Zz, Yy = np.meshgrid(z, y, indexing="ij")
    w = np.zeros_like(Yy)
    for m in range(1, 2 * terms, 2):
        w += (1.0 / m ** 3) * (1 - np.cosh(m * np.pi * Yy / b) / np.cosh(m * np.pi * a / b)) * np.sin(m * np.pi * Zz / b)
    return w / w.mean()


def random_scenario(seed, Q):
    """A random small scenario: random sizes, 3-9 random boxes of every handler kind (faces, slabs, interior
    blocks, overlapping, later ones overwrite earlier ones), random solid mask, random near-equilibrium state.
    Deliberately ugly: exercises last-writer-wins, handlers inside the domain, uncovered shell cells."""
    rng = np.random.default_rng(seed)
    xl, yl, zl = (int(v) for v in rng.integers(3, 12, size=3))
    kinds = [O.NOSLIP, O.MOVINGWALL, O.FREESLIP, O.OUTFLOW, O.INFLOW, O.PRESSURE, O.PARALLEL]
    boxes = []
    faces = ["z0", "zmax", "x0", "xmax", "y0", "ymax"]
    rng.shuffle(faces)
    n_faces = int(rng.integers(2, 7))
    spec = []
    for name in faces[:n_faces]:
        k = int(rng.choice(kinds))
        spec.append((name, k, tuple(rng.uniform(-0.05, 0.05, 3)), float(rng.uniform(0.95, 1.05))))
    boxes += O.face_boxes(xl, yl, zl, spec)
    for _ in range(int(rng.integers(1, 4))):          # interior / straddling blocks
        lo = [int(rng.integers(0, n + 1)) for n in (xl, yl, zl)]
        hi = [int(min(n + 1, l + rng.integers(0, 4))) for n, l in zip((xl, yl, zl), lo)]
        k = int(rng.choice(kinds))
        boxes.append((k, tuple(float(v) for v in rng.uniform(-0.05, 0.05, 3)), float(rng.uniform(0.95, 1.05)),
                      (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])))
    mask = (rng.random((zl, yl, xl)) > 0.1).astype(np.uint8) if rng.random() < 0.5 else None
    _, w = O.oracle().model(Q)
    n_all = (xl + 2) * (yl + 2) * (zl + 2)
    f0 = np.tile(w, (n_all, 1)) * (1 + 0.05 * rng.standard_normal((n_all, Q)))
    case = dict(xl=xl, yl=yl, zl=zl, boxes=boxes, f_init=f0)
    if mask is not None:
        case["fluid_mask"] = mask
    return case
