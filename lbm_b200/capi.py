"""ctypes binding of the C ABI in include/lbm_b200.h (lbm_b200/liblbm_b200.so).

This is the Python-side stub the tests and bench.py use; the reference-facing
host surface is the C++ one under include/lbm/.  There is no fallback: if the
shared library is missing, import fails; if no CUDA device is present, every
compute entry point raises LbmError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBM_B200_LIB selects a tuning variant built by tools/build_variants.sh (same ABI, same sources)
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")

FLUID, NOSLIP, MOVINGWALL, FREESLIP, OUTFLOW, INFLOW, PRESSURE, NULL, PARALLEL, PERIODIC = range(10)
FAST, EXACT = 0, 1
AOS, SOA = 0, 1
COLLIDE_FIELD, STREAM_FIELD = 0, 1
DOWN, UP = 0, 1
AXIS_Y, AXIS_Z = 1, 2
EXPORT_BYTES = 256

# every symbol include/lbm_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "lbm_b200_last_error", "lbm_b200_abi_version", "lbm_b200_device_count",
    "lbm_b200_model", "lbm_b200_model_inv", "lbm_b200_model_velocity_index",
    "lbm_b200_create", "lbm_b200_create_slab", "lbm_b200_create_slab_axis", "lbm_b200_destroy",
    "lbm_b200_set_arithmetic", "lbm_b200_set_tau", "lbm_b200_set_stream",
    "lbm_b200_set_handlers", "lbm_b200_paint_boxes", "lbm_b200_set_boxes", "lbm_b200_set_geometry",
    "lbm_b200_set_geometry_planes", "lbm_b200_get_geometry_planes", "lbm_b200_get_kind",
    "lbm_b200_set_fluid_mask", "lbm_b200_set_fluid_mask_literal", "lbm_b200_set_fluid_mask_global",
    "lbm_b200_paint_mask", "lbm_b200_tag_null_cells",
    "lbm_b200_upload_populations", "lbm_b200_download_populations", "lbm_b200_init_equilibrium",
    "lbm_b200_save_checkpoint", "lbm_b200_load_checkpoint", "lbm_b200_upload_planes", "lbm_b200_download_planes",
    "lbm_b200_step", "lbm_b200_step_group", "lbm_b200_set_graphs", "lbm_b200_set_sweep_engine", "lbm_b200_sync", "lbm_b200_elapsed_ms", "lbm_b200_launch_count", "lbm_b200_tma_launch_count", "lbm_b200_steps_done",
    "lbm_b200_macroscopic", "lbm_b200_macroscopic_begin", "lbm_b200_macroscopic_end", "lbm_b200_diagnostics",
    "lbm_b200_host_alloc", "lbm_b200_host_free", "lbm_b200_bind_host_thread", "lbm_b200_selftest_division",
    "lbm_b200_halo_layout", "lbm_b200_halo_plane", "lbm_b200_edge_plane", "lbm_b200_dst_buffer",
    "lbm_b200_step_edges", "lbm_b200_step_interior", "lbm_b200_step_finish",
    "lbm_b200_export", "lbm_b200_connect", "lbm_b200_connect_local", "lbm_b200_halo_push_all", "lbm_b200_halo_pushed",
    "lbm_b200_disconnect",
]


class LbmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("lbm_b200 error %d: %s" % (code, message))
        self.code = code


class Bc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("v", C.c_double * 3), ("rho", C.c_double)]


if not os.path.exists(LIB_PATH):
    raise ImportError("%s not found: build it with `make` (python -c 'import __graft_entry__ as g; g.build()'); "
                      "there is no CPU fallback" % LIB_PATH)
lib = C.CDLL(LIB_PATH)

_H = C.c_void_p
lib.lbm_b200_last_error.restype = C.c_char_p
lib.lbm_b200_steps_done.restype = C.c_uint64
lib.lbm_b200_steps_done.argtypes = [_H]
lib.lbm_b200_model.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
lib.lbm_b200_create.argtypes = [C.POINTER(_H), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int]
lib.lbm_b200_create_slab.argtypes = [C.POINTER(_H), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_double, C.c_int]
lib.lbm_b200_create_slab_axis.argtypes = [C.POINTER(_H), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                                          C.c_uint64, C.c_double, C.c_int]
lib.lbm_b200_destroy.argtypes = [_H]
lib.lbm_b200_set_arithmetic.argtypes = [_H, C.c_int]
lib.lbm_b200_set_tau.argtypes = [_H, C.c_double]
lib.lbm_b200_set_stream.argtypes = [_H, C.c_void_p]
lib.lbm_b200_set_geometry.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
lib.lbm_b200_set_boxes.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_int]
lib.lbm_b200_set_handlers.argtypes = [_H, C.c_void_p, C.c_int]
lib.lbm_b200_paint_boxes.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_int]
lib.lbm_b200_set_geometry_planes.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
lib.lbm_b200_get_geometry_planes.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
lib.lbm_b200_set_fluid_mask.argtypes = [_H, C.c_void_p]
lib.lbm_b200_set_fluid_mask_literal.argtypes = [_H, C.c_void_p]
lib.lbm_b200_set_fluid_mask_global.argtypes = [_H, C.c_void_p, C.c_int]
lib.lbm_b200_paint_mask.argtypes = [_H, C.c_void_p, C.c_uint16, C.c_int]
lib.lbm_b200_tag_null_cells.argtypes = [_H, C.c_int, C.POINTER(C.c_uint64)]
lib.lbm_b200_get_kind.argtypes = [_H, C.c_void_p]
lib.lbm_b200_step_group.argtypes = [C.POINTER(_H), C.c_int, C.c_uint64]
lib.lbm_b200_set_graphs.argtypes = [_H, C.c_int]
lib.lbm_b200_set_sweep_engine.argtypes = [_H, C.c_int, C.c_int]
lib.lbm_b200_macroscopic_begin.argtypes = [_H, C.c_void_p, C.c_void_p]
lib.lbm_b200_macroscopic_end.argtypes = [_H]
lib.lbm_b200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]
lib.lbm_b200_host_free.argtypes = [C.c_void_p]
lib.lbm_b200_bind_host_thread.argtypes = [C.c_int]
lib.lbm_b200_selftest_division.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.POINTER(C.c_uint64)]
lib.lbm_b200_upload_populations.argtypes = [_H, C.c_void_p, C.c_int, C.c_int]
lib.lbm_b200_download_populations.argtypes = [_H, C.c_void_p, C.c_int, C.c_int]
lib.lbm_b200_init_equilibrium.argtypes = [_H, C.c_void_p, C.c_void_p]
lib.lbm_b200_upload_planes.argtypes = [_H, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
lib.lbm_b200_download_planes.argtypes = [_H, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
lib.lbm_b200_save_checkpoint.argtypes = [_H, C.c_char_p]
lib.lbm_b200_load_checkpoint.argtypes = [_H, C.c_char_p]
lib.lbm_b200_step.argtypes = [_H, C.c_uint64]
lib.lbm_b200_sync.argtypes = [_H]
lib.lbm_b200_elapsed_ms.argtypes = [_H, C.POINTER(C.c_double)]
lib.lbm_b200_launch_count.argtypes = [_H, C.POINTER(C.c_uint64)]
lib.lbm_b200_tma_launch_count.argtypes = [_H, C.POINTER(C.c_uint64)]
lib.lbm_b200_macroscopic.argtypes = [_H, C.c_void_p, C.c_void_p]
lib.lbm_b200_diagnostics.argtypes = [_H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
lib.lbm_b200_halo_layout.argtypes = [_H, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
lib.lbm_b200_halo_plane.argtypes = [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.lbm_b200_edge_plane.argtypes = [_H, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.lbm_b200_dst_buffer.argtypes = [_H]
lib.lbm_b200_step_edges.argtypes = [_H]
lib.lbm_b200_step_interior.argtypes = [_H]
lib.lbm_b200_step_finish.argtypes = [_H]
lib.lbm_b200_export.argtypes = [_H, C.c_void_p]
lib.lbm_b200_connect.argtypes = [_H, C.c_int, C.c_void_p]
lib.lbm_b200_connect_local.argtypes = [_H, C.c_int, _H]
lib.lbm_b200_halo_push_all.argtypes = [_H]
lib.lbm_b200_halo_pushed.argtypes = [_H]
lib.lbm_b200_disconnect.argtypes = [_H]


def _check(rc):
    if rc != 0:
        raise LbmError(rc, lib.lbm_b200_last_error().decode("utf-8", "replace"))


def device_count():
    return lib.lbm_b200_device_count()


def selftest_division(n, seed=1, tau=0.6):
    """mismatches between the bit-identical mode's reciprocal-based quotients and IEEE division on n operands per divisor"""
    bad = C.c_uint64(0)
    _check(lib.lbm_b200_selftest_division(n, seed, tau, C.byref(bad)))
    return bad.value


def model(Q):
    """(velocities[Q,3], weights[Q]) of model.h's d3q15/d3q19/d3q27."""
    c = np.zeros((Q, 3))
    w = np.zeros(Q)
    _check(lib.lbm_b200_model(Q, c.ctypes.data, w.ctypes.data))
    return c, w


def velocity_index(Q, u, v, w):
    r = lib.lbm_b200_model_velocity_index(Q, u, v, w)
    if r < -1:
        _check(r)
    return r


def _boxes_arrays(boxes):
    """boxes: list of (kind, v3, rho, (x0,xE,y0,yE,z0,zE)) (the box-list format the tests use)."""
    n = len(boxes)
    ext = np.zeros((max(n, 1), 6), dtype=np.uint64)
    tab = (Bc * max(n, 1))()
    for i, (kind, v, rho, e) in enumerate(boxes):
        ext[i] = e
        tab[i].kind = kind
        tab[i].v[0], tab[i].v[1], tab[i].v[2] = v
        tab[i].rho = rho
    return ext, tab


def paint_boxes(xl, yl, zl_local, z_first, boxes):
    """Dense kind / bc-id maps (Domain::idx order, local planes 0..zl_local+1 of a slab whose first interior
    plane is global z_first) and the handler table for a box list -- what Domain::setBoundaryCondition builds
    on the host (domain.hpp:175-194), last writer wins."""
    kind = np.zeros((zl_local + 2, yl + 2, xl + 2), dtype=np.uint8)
    bcid = np.zeros((zl_local + 2, yl + 2, xl + 2), dtype=np.uint16)
    table = []
    off = z_first - 1
    for (k, v, rho, (x0, xE, y0, yE, z0, zE)) in boxes:
        table.append((k, v, rho))
        lz0, lz1 = max(z0 - off, 0), min(zE - off, zl_local + 1)
        if lz0 > lz1:
            continue
        kind[lz0:lz1 + 1, y0:yE + 1, x0:xE + 1] = k
        bcid[lz0:lz1 + 1, y0:yE + 1, x0:xE + 1] = len(table) - 1
    return kind.reshape(-1), bcid.reshape(-1), table


class Domain:
    """Python mirror of lbm::Domain<M> + BGKCollision<M>(tau) backed by one GPU slab.

    Domain(Q, xl, yl, zl, tau)                        whole domain
    Domain(Q, xl, yl, zl_global, tau, z_first=, zl_local=)   one z-slab
    Domain(Q, xl, yl_global, zl, tau, axis=AXIS_Y, z_first=, zl_local=)   one y-slab: rows z_first .. z_first+zl_local-1
    xl, yl, zl of the object are the LOCAL lengths (what its host arrays use); yl_global / zl_global the domain's.
    """

    def __init__(self, Q, xl, yl, zl, tau, device=-1, z_first=None, zl_local=None, exact=False, axis=AXIS_Z):
        self._h = _H()
        self.Q, self.xl, self.yl, self.zl, self.tau, self.axis = Q, xl, yl, zl, tau, axis
        self.yl_global, self.zl_global = yl, zl
        if z_first is None:
            self.z_first = 1
            _check(lib.lbm_b200_create(C.byref(self._h), Q, xl, yl, zl, tau, device))
        else:
            self.z_first = z_first
            if axis == AXIS_Y:
                self.yl = zl_local
            else:
                self.zl = zl_local
            _check(lib.lbm_b200_create_slab_axis(C.byref(self._h), Q, xl, yl, zl, axis, z_first, zl_local, tau, device))
        self.ncell = (xl + 2) * (self.yl + 2) * (self.zl + 2)
        if exact:
            self.set_arithmetic(EXACT)

    # -- life cycle
    def close(self):
        if self._h:
            lib.lbm_b200_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- configuration
    def set_arithmetic(self, mode):
        _check(lib.lbm_b200_set_arithmetic(self._h, mode))

    def set_stream(self, cuda_stream):
        _check(lib.lbm_b200_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_boxes(self, boxes):
        ext, tab = _boxes_arrays(boxes)
        _check(lib.lbm_b200_set_boxes(self._h, ext.ctypes.data, C.cast(tab, C.c_void_p), len(boxes)))

    def set_geometry(self, kind, bc_id=None, table=()):
        kind = np.ascontiguousarray(kind, dtype=np.uint8).reshape(-1)
        assert kind.size == self.ncell
        _, tab = _boxes_arrays([(k, v, rho, (0,) * 6) for (k, v, rho) in table])
        bid = None
        if bc_id is not None:
            bid = np.ascontiguousarray(bc_id, dtype=np.uint16).reshape(-1)
            assert bid.size == self.ncell
        _check(lib.lbm_b200_set_geometry(self._h, kind.ctypes.data, bid.ctypes.data if bid is not None else None,
                                         C.cast(tab, C.c_void_p), len(table)))

    def set_handlers(self, table):
        """table: list of (kind, v3, rho); replaces (extends) the handler table"""
        _, tab = _boxes_arrays([(k, v, rho, (0,) * 6) for (k, v, rho) in table])
        _check(lib.lbm_b200_set_handlers(self._h, C.cast(tab, C.c_void_p), len(table)))

    def paint_boxes(self, extents, ids):
        ext = np.ascontiguousarray(extents, dtype=np.uint64).reshape(-1, 6)
        ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(-1)
        assert ext.shape[0] == ids.size
        _check(lib.lbm_b200_paint_boxes(self._h, ext.ctypes.data, ids.ctypes.data, ids.size))

    def set_geometry_planes(self, kind, bc_id, z_begin, z_count, literal=False):
        kind = np.ascontiguousarray(kind, dtype=np.uint8).reshape(-1)
        bid = np.ascontiguousarray(bc_id, dtype=np.uint16).reshape(-1)
        assert kind.size == bid.size == z_count * (self.yl + 2) * (self.xl + 2)
        _check(lib.lbm_b200_set_geometry_planes(self._h, kind.ctypes.data, bid.ctypes.data, z_begin, z_count, int(literal)))

    def geometry_planes(self, z_begin=0, z_count=None):
        """(kind, handler id) of the collide field as Domain::cell() reports them"""
        if z_count is None:
            z_count = self.zl + 2 - z_begin
        n = z_count * (self.yl + 2) * (self.xl + 2)
        k = np.empty(n, dtype=np.uint8)
        b = np.empty(n, dtype=np.uint16)
        _check(lib.lbm_b200_get_geometry_planes(self._h, k.ctypes.data, b.ctypes.data, z_begin, z_count))
        return k, b

    def set_fluid_mask(self, mask, literal=False):
        """mask of this handle's own interior planes (io/vtk.hpp:137-150); literal: collide field only"""
        m = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        assert m.size == self.xl * self.yl * self.zl
        if literal:
            _check(lib.lbm_b200_set_fluid_mask_literal(self._h, m.ctypes.data))
        else:
            _check(lib.lbm_b200_set_fluid_mask(self._h, m.ctypes.data))

    def set_fluid_mask_global(self, mask, literal=False):
        """mask of the WHOLE domain; a slab picks its planes and its neighbours' edge planes"""
        m = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        assert m.size == self.xl * self.yl_global * self.zl_global
        _check(lib.lbm_b200_set_fluid_mask_global(self._h, m.ctypes.data, int(literal)))

    def tag_null_cells(self, literal=False):
        n = C.c_uint64()
        _check(lib.lbm_b200_tag_null_cells(self._h, int(literal), C.byref(n)))
        return n.value

    def set_graphs(self, mode):
        _check(lib.lbm_b200_set_graphs(self._h, mode))

    def set_sweep_engine(self, tma=-1, checked=-1):
        _check(lib.lbm_b200_set_sweep_engine(self._h, tma, checked))

    def kind(self):
        k = np.empty(self.ncell, dtype=np.uint8)
        _check(lib.lbm_b200_get_kind(self._h, k.ctypes.data))
        return k

    # -- state
    def upload(self, f, layout=AOS, field=COLLIDE_FIELD):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        assert f.size == self.ncell * self.Q
        _check(lib.lbm_b200_upload_populations(self._h, f.ctypes.data, layout, field))

    def download(self, layout=AOS, field=COLLIDE_FIELD, out=None):
        shape = (self.ncell, self.Q) if layout == AOS else (self.Q, self.ncell)
        f = out if out is not None else np.empty(shape)
        _check(lib.lbm_b200_download_populations(self._h, f.ctypes.data, layout, field))
        return f

    def init_equilibrium(self, rho, u):
        rho = np.ascontiguousarray(rho, dtype=np.float64).reshape(-1)
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        assert rho.size == self.ncell and u.size == 3 * self.ncell
        _check(lib.lbm_b200_init_equilibrium(self._h, rho.ctypes.data, u.ctypes.data))

    def download_planes(self, z_begin, z_count, field=COLLIDE_FIELD):
        f = np.empty((z_count * (self.yl + 2) * (self.xl + 2), self.Q))
        _check(lib.lbm_b200_download_planes(self._h, f.ctypes.data, field, z_begin, z_count))
        return f

    def upload_planes(self, f, z_begin, z_count, field=COLLIDE_FIELD):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        assert f.size == z_count * (self.yl + 2) * (self.xl + 2) * self.Q
        _check(lib.lbm_b200_upload_planes(self._h, f.ctypes.data, field, z_begin, z_count))

    def save_checkpoint(self, path):
        _check(lib.lbm_b200_save_checkpoint(self._h, str(path).encode()))

    def load_checkpoint(self, path):
        _check(lib.lbm_b200_load_checkpoint(self._h, str(path).encode()))

    # -- hot path
    def step(self, n=1):
        _check(lib.lbm_b200_step(self._h, n))

    def sync(self):
        _check(lib.lbm_b200_sync(self._h))

    def elapsed_ms(self):
        ms = C.c_double()
        _check(lib.lbm_b200_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        _check(lib.lbm_b200_launch_count(self._h, C.byref(n)))
        return n.value

    def tma_launch_count(self):
        n = C.c_uint64(0)
        _check(lib.lbm_b200_tma_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def steps_done(self):
        return lib.lbm_b200_steps_done(self._h)

    # -- read-out
    def macroscopic(self, rho=None, u=None):
        """(rho[zl,yl,xl], u[zl,yl,xl,3]) of the interior cells, io/vtk.hpp:62-73 order."""
        if rho is None:
            rho = np.empty((self.zl, self.yl, self.xl))
        if u is None:
            u = np.empty((self.zl, self.yl, self.xl, 3))
        _check(lib.lbm_b200_macroscopic(self._h, rho.ctypes.data, u.ctypes.data))
        return rho, u

    def macroscopic_begin(self, rho_ptr, u_ptr):
        """split read-out into caller-owned (pinned) host memory given as raw addresses"""
        _check(lib.lbm_b200_macroscopic_begin(self._h, C.c_void_p(rho_ptr or 0), C.c_void_p(u_ptr or 0)))

    def macroscopic_end(self):
        _check(lib.lbm_b200_macroscopic_end(self._h))

    def diagnostics(self):
        m, k, um = C.c_double(), C.c_double(), C.c_double()
        _check(lib.lbm_b200_diagnostics(self._h, C.byref(m), C.byref(k), C.byref(um)))
        return m.value, k.value, um.value

    # -- slabs
    def halo_layout(self):
        n, b = C.c_int(), C.c_size_t()
        _check(lib.lbm_b200_halo_layout(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    def halo_plane(self, buffer, side, k, recv):
        p = C.c_void_p()
        _check(lib.lbm_b200_halo_plane(self._h, buffer, side, k, int(recv), C.byref(p)))
        return p.value

    def edge_plane(self, side, q, recv):
        p = C.c_void_p()
        _check(lib.lbm_b200_edge_plane(self._h, side, q, int(recv), C.byref(p)))
        return p.value

    def dst_buffer(self):
        return lib.lbm_b200_dst_buffer(self._h)

    def step_edges(self):
        _check(lib.lbm_b200_step_edges(self._h))

    def step_interior(self):
        _check(lib.lbm_b200_step_interior(self._h))

    def step_finish(self):
        _check(lib.lbm_b200_step_finish(self._h))

    def export(self):
        blob = C.create_string_buffer(EXPORT_BYTES)
        _check(lib.lbm_b200_export(self._h, blob))
        return blob.raw

    def connect(self, side, blob):
        _check(lib.lbm_b200_connect(self._h, side, C.c_char_p(blob)))

    def connect_local(self, side, other):
        _check(lib.lbm_b200_connect_local(self._h, side, other._h))

    def disconnect(self):
        _check(lib.lbm_b200_disconnect(self._h))

    def halo_push_all(self):
        _check(lib.lbm_b200_halo_push_all(self._h))

    def halo_pushed(self):
        _check(lib.lbm_b200_halo_pushed(self._h))


def step_group(domains, n):
    """n steps of a stack of connected slabs from one host thread (lbm_b200_step_group)"""
    arr = (_H * len(domains))(*[d._h for d in domains])
    _check(lib.lbm_b200_step_group(arr, len(domains), n))


class HostBuffer:
    """page-locked host array next to a GPU (lbm_b200_host_alloc), exposed as a numpy array"""

    def __init__(self, n, dtype=np.float64, device=-1):
        self.ptr = C.c_void_p()
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        _check(lib.lbm_b200_host_alloc(C.byref(self.ptr), max(self.nbytes, 1), device))
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    @property
    def address(self):
        return self.ptr.value

    def free(self):
        if self.ptr:
            self.array = None
            lib.lbm_b200_host_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
