"""ctypes access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

``oracle``  -> oracle/liblbm_oracle.so   plain-C restatement (oracle/lbm_oracle.c)
``ref``     -> oracle/_ref/libref_lbm.so the reference's own headers, compiled from
               /root/reference by `make -C oracle ref` (oracle/ref_driver.cpp)

Both export the interface of oracle/oracle.h.  The product (lbm_b200/) never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

FLUID, NOSLIP, MOVINGWALL, FREESLIP, OUTFLOW, INFLOW, PRESSURE, NULL, PARALLEL, PERIODIC = range(10)


class Box(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("v", C.c_double * 3), ("rho", C.c_double),
                ("x0", C.c_uint64), ("xE", C.c_uint64), ("y0", C.c_uint64), ("yE", C.c_uint64),
                ("z0", C.c_uint64), ("zE", C.c_uint64)]


class Case(C.Structure):
    _fields_ = [("Q", C.c_int32), ("threads", C.c_int32),
                ("xl", C.c_uint64), ("yl", C.c_uint64), ("zl", C.c_uint64),
                ("tau", C.c_double), ("n_boxes", C.c_int32), ("periodic", C.c_int32),
                ("boxes", C.POINTER(Box)), ("fluid_mask", C.c_void_p), ("f_init", C.c_void_p),
                ("null_opt", C.c_int32), ("mask_literal", C.c_int32), ("steps", C.c_uint64), ("untimed", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("f", C.c_void_p), ("rho", C.c_void_p), ("u", C.c_void_p), ("kind", C.c_void_p),
                ("seconds", C.c_double)]


def face_boxes(xl, yl, zl, spec):
    """Boxes for named extents in the given order, like io/scenario.h:99-116.

    spec: list of (extent, kind, v, rho) with extent in z0,zmax,x0,xmax,y0,ymax or a 6-tuple.
    """
    named = {
        "z0": (0, xl + 1, 0, yl + 1, 0, 0), "zmax": (0, xl + 1, 0, yl + 1, zl + 1, zl + 1),
        "x0": (0, 0, 0, yl + 1, 0, zl + 1), "xmax": (xl + 1, xl + 1, 0, yl + 1, 0, zl + 1),
        "y0": (0, xl + 1, 0, 0, 0, zl + 1), "ymax": (0, xl + 1, yl + 1, yl + 1, 0, zl + 1),
    }
    out = []
    for item in spec:
        extent, kind = item[0], item[1]
        v = item[2] if len(item) > 2 and item[2] is not None else (0.0, 0.0, 0.0)
        rho = item[3] if len(item) > 3 and item[3] is not None else 1.0
        ext = named[extent] if isinstance(extent, str) else tuple(extent)
        out.append((kind, tuple(float(a) for a in v), float(rho), tuple(int(e) for e in ext)))
    return out


def cavity_boxes(xl, yl, zl, lid=(0.05, 0.0, 0.0)):
    """build/scenarios/cavity.xml:6-11 order: z0 noslip, zmax lid, x0, xmax, y0, ymax noslip."""
    return face_boxes(xl, yl, zl, [("z0", NOSLIP), ("zmax", MOVINGWALL, lid), ("x0", NOSLIP),
                                   ("xmax", NOSLIP), ("y0", NOSLIP), ("ymax", NOSLIP)])


def channel_boxes(xl, yl, zl, u_in=(0.03, 0.0, 0.0), rho_ref=1.0):
    """build/scenarios/pipe.xml:5-10 order: z0, zmax noslip, x0 inflow, xmax outflow, y0, ymax noslip."""
    return face_boxes(xl, yl, zl, [("z0", NOSLIP), ("zmax", NOSLIP), ("x0", INFLOW, u_in, rho_ref),
                                   ("xmax", OUTFLOW, None, rho_ref), ("y0", NOSLIP), ("ymax", NOSLIP)])


class Checker:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self._run = getattr(self.lib, prefix + "run")
        self._run.argtypes = [C.POINTER(Case), C.POINTER(Result)]
        self._run.restype = C.c_int
        g = lambda n: getattr(self.lib, prefix + n)
        self._density = g("density"); self._density.restype = C.c_double
        self._density.argtypes = [C.c_int, C.c_void_p]
        self._velocity = g("velocity"); self._velocity.restype = None
        self._velocity.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        self._feq = g("feq"); self._feq.restype = None
        self._feq.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        self._bgk = g("bgk"); self._bgk.restype = None
        self._bgk.argtypes = [C.c_int, C.c_double, C.c_void_p]
        self._model = g("model"); self._model.restype = C.c_int
        self._model.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        self._vi = g("velocity_index"); self._vi.restype = C.c_int
        self._vi.argtypes = [C.c_int] * 4

    # ---- single-cell helpers -------------------------------------------------
    def density(self, Q, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        return self._density(Q, f.ctypes.data)

    def velocity(self, Q, f, rho):
        f = np.ascontiguousarray(f, dtype=np.float64)
        u = np.zeros(3)
        self._velocity(Q, f.ctypes.data, rho, u.ctypes.data)
        return u

    def feq(self, Q, rho, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros(Q)
        self._feq(Q, rho, u.ctypes.data, out.ctypes.data)
        return out

    def bgk(self, Q, tau, f):
        f = np.array(f, dtype=np.float64)
        self._bgk(Q, tau, f.ctypes.data)
        return f

    def model(self, Q):
        c = np.zeros((Q, 3)); w = np.zeros(Q)
        rc = self._model(Q, c.ctypes.data, w.ctypes.data)
        assert rc == 0
        return c, w

    def velocity_index(self, Q, u, v, w):
        return self._vi(Q, u, v, w)

    # ---- whole-lattice run ----------------------------------------------------
    def run(self, Q, xl, yl, zl, tau, boxes, steps, *, f_init=None, fluid_mask=None, periodic=False,
            null_opt=True, mask_literal=False, threads=None, untimed=0, want=("f", "rho", "u", "kind")):
        """Returns dict(f=[ncell,Q] AoS in Domain::idx order, rho=[zl,yl,xl], u=[zl,yl,xl,3],
        kind=[ncell], seconds)."""
        n_all = (xl + 2) * (yl + 2) * (zl + 2)
        n_int = xl * yl * zl
        arr = (Box * max(1, len(boxes)))()
        for i, (kind, v, rho, ext) in enumerate(boxes):
            arr[i].kind = kind
            arr[i].v[0], arr[i].v[1], arr[i].v[2] = v
            arr[i].rho = rho
            arr[i].x0, arr[i].xE, arr[i].y0, arr[i].yE, arr[i].z0, arr[i].zE = ext
        c = Case()
        c.Q, c.threads = Q, int(threads or os.cpu_count() or 1)
        c.xl, c.yl, c.zl, c.tau = xl, yl, zl, tau
        c.n_boxes, c.periodic = len(boxes), int(bool(periodic))
        c.boxes = arr
        keep = []
        if fluid_mask is not None:
            m = np.ascontiguousarray(fluid_mask, dtype=np.uint8).reshape(-1)
            assert m.size == n_int
            keep.append(m); c.fluid_mask = m.ctypes.data
        if f_init is not None:
            fi = np.ascontiguousarray(f_init, dtype=np.float64).reshape(-1)
            assert fi.size == n_all * Q
            keep.append(fi); c.f_init = fi.ctypes.data
        c.null_opt, c.mask_literal, c.steps = int(bool(null_opt)), int(bool(mask_literal)), int(steps)
        c.untimed = int(untimed)
        r = Result()
        out = {}
        if "f" in want:
            out["f"] = np.empty((n_all, Q)); r.f = out["f"].ctypes.data
        if "rho" in want:
            out["rho"] = np.empty((zl, yl, xl)); r.rho = out["rho"].ctypes.data
        if "u" in want:
            out["u"] = np.empty((zl, yl, xl, 3)); r.u = out["u"].ctypes.data
        if "kind" in want:
            out["kind"] = np.empty(n_all, dtype=np.uint8); r.kind = out["kind"].ctypes.data
        rc = self._run(C.byref(c), C.byref(r))
        if rc != 0:
            raise RuntimeError("%srun failed with code %d" % (self.prefix, rc))
        out["seconds"] = r.seconds
        return out


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liblbm_oracle.so"])


def have_ref_sources():
    return os.path.isdir("/root/reference/include")


def build_ref():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_cache = {}


def ref_fast():
    """the reference's own headers again, compiled -O3 -mavx2 -mfma (FMA contraction allowed): the generous
    CPU baseline of BASELINE.md; timing only, not bit-identical to the parity build.  None if not built."""
    if "ref_fast" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_lbm_fast.so")
        if not os.path.exists(path) and have_ref_sources():
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref_fast"])
        _cache["ref_fast"] = Checker(path, "ref_") if os.path.exists(path) else None
    return _cache["ref_fast"]


def oracle():
    if "oracle" not in _cache:
        path = os.path.join(ORACLE_DIR, "liblbm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        _cache["oracle"] = Checker(path, "oracle_")
    return _cache["oracle"]


def ref():
    """The compiled reference, or None when neither the prebuilt .so nor /root/reference exists."""
    if "ref" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_lbm.so")
        if not os.path.exists(path):
            if not have_ref_sources():
                _cache["ref"] = None
                return None
            build_ref()
        _cache["ref"] = Checker(path, "ref_")
    return _cache["ref"]
