"""GPU: the C++ host surface (include/lbm, reference class names) driven exactly like the reference's
main/scenario code, compared with the oracle on the same inputs; plus the command-line driver."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import _oracle as O
from test_host_surface import compile_cpp, ROOT

pytestmark = pytest.mark.gpu

SCENARIOS = {
    "step": ("scenarios/step_small.xml", dict(xl=20, yl=12, zl=30, boxes=O.face_boxes(20, 12, 30, [
        ((11, 21, 0, 13, 0, 0), O.INFLOW, (0.0, 0.0, 0.04), 1.0), ((0, 10, 0, 13, 0, 12), O.NOSLIP),
        ("zmax", O.OUTFLOW), ("y0", O.NOSLIP), ("ymax", O.NOSLIP), ("x0", O.NOSLIP), ("xmax", O.NOSLIP)]))),
    "shear": ("scenarios/shear_small.xml", dict(xl=8, yl=8, zl=20, boxes=O.face_boxes(8, 8, 20, [
        ("z0", O.PRESSURE, None, 1.005), ("zmax", O.OUTFLOW), ("x0", O.FREESLIP), ("xmax", O.FREESLIP),
        ("y0", O.NOSLIP), ("ymax", O.NOSLIP)]))),
}


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "host_api_check")
    compile_cpp(os.path.join(ROOT, "tests", "cpp", "host_api_check.cpp"), out)
    return out


def read_dump(path, steps):
    raw = open(path, "rb").read()
    xl, yl, zl, Q = struct.unpack("4i", raw[:16])
    n_all, n_int = (xl + 2) * (yl + 2) * (zl + 2), xl * yl * zl
    off = 16
    def take(count, dtype=np.float64):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    f0 = take(n_all * Q).reshape(n_all, Q)
    f1 = take(n_all * Q).reshape(n_all, Q)
    kinds = take(n_all, np.uint8)
    rho = take(n_int).reshape(zl, yl, xl)
    u = take(3 * n_int).reshape(zl, yl, xl, 3)
    assert off == len(raw)
    return (xl, yl, zl, Q), f0, f1, kinds, rho, u


@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("name", sorted(SCENARIOS))
@pytest.mark.parametrize("gpus", [1, 2])
def test_reference_style_calls_match_oracle(exe, tmp_path, Q, name, gpus):
    from lbm_b200 import capi
    if gpus > capi.device_count():
        pytest.skip("needs %d GPUs" % gpus)
    xml, case = SCENARIOS[name]
    steps = 25
    cfg = tmp_path / "run.cfg"
    cfg.write_text("tau = 0.6\ntimesteps = %d\ntimesteps-per-plot = 0\noutput-dir = %s\nscenario-file = %s\n"
                   % (steps, tmp_path / "vtk", os.path.join(ROOT, xml)))
    env = dict(os.environ, LBM_B200_ARITHMETIC="exact", LBM_B200_GPUS=str(gpus))
    dump = tmp_path / "dump.bin"
    r = subprocess.run([exe, str(Q), str(cfg), str(steps), str(dump)], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "HOST_API_CHECK DONE gpus=%d" % gpus in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    dims, f0, f1, kinds, rho, u = read_dump(dump, steps)
    assert dims == (case["xl"], case["yl"], case["zl"], Q)
    want = O.oracle().run(Q, case["xl"], case["yl"], case["zl"], 0.6, case["boxes"], steps, f_init=f0)
    assert np.array_equal(kinds, want["kind"])
    fluid = want["kind"] == O.FLUID
    assert np.array_equal(f1[fluid], want["f"][fluid])
    if gpus == 1:
        assert np.array_equal(f1, want["f"])          # boundary cells too (materialised on read-back)
        assert np.array_equal(rho, want["rho"]) and np.array_equal(u, want["u"])
    # the .vts file holds the same density / velocity (raw appended Float64 blocks)
    vts = tmp_path / "vtk" / ("%s.%d.vts" % ("Step" if name == "step" else "Shear", steps))
    blob = open(vts, "rb").read()
    start = blob.index(b"<AppendedData encoding=\"raw\">") + len(b"<AppendedData encoding=\"raw\">")
    start = blob.index(b"_", start) + 1
    nv = struct.unpack("<Q", blob[start:start + 8])[0]
    vel = np.frombuffer(blob, dtype=np.float64, count=nv // 8, offset=start + 8).reshape(u.shape)
    nd = struct.unpack("<Q", blob[start + 8 + nv:start + 16 + nv])[0]
    den = np.frombuffer(blob, dtype=np.float64, count=nd // 8, offset=start + 16 + nv).reshape(rho.shape)
    assert np.array_equal(vel, u) and np.array_equal(den, rho)
    assert re.search(rb'WholeExtent="0 %d 0 %d 0 %d"' % (case["xl"] - 1, case["yl"] - 1, case["zl"] - 1), blob)


def test_command_line_driver_runs_a_scenario(tmp_path):
    exe = str(tmp_path / "lbm")
    compile_cpp(os.path.join(ROOT, "src", "main.cpp"), exe)
    cfg = tmp_path / "c.cfg"
    cfg.write_text("collision-model = bgk\ntau = 0.6\ntimesteps = 40\ntimesteps-per-plot = 20\noutput-dir = %s\n"
                   "scenario-file = scenarios/cavity64.xml\nlattice = 19\n" % (tmp_path / "vtk"))
    r = subprocess.run([exe, str(cfg)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MLUPS:" in r.stdout and "Finished!" in r.stdout and "Reading scenario: \"Cavity64\"" in r.stdout
    assert sorted(os.listdir(tmp_path / "vtk")) == ["Cavity64.20.vts", "Cavity64.40.vts"]
