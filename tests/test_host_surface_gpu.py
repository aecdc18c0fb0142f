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
@pytest.mark.parametrize("gpus", [1, 2, "2y"])
def test_reference_style_calls_match_oracle(exe, tmp_path, Q, name, gpus):
    """gpus = "2y": the same two-GPU Domain split along y (x-z planes) instead of z"""
    from lbm_b200 import capi
    axis = "y" if gpus == "2y" else "z"
    gpus = int(str(gpus)[0])
    if gpus > capi.device_count():
        pytest.skip("needs %d GPUs" % gpus)
    xml, case = SCENARIOS[name]
    steps = 25
    cfg = tmp_path / "run.cfg"
    cfg.write_text("tau = 0.6\ntimesteps = %d\ntimesteps-per-plot = 0\noutput-dir = %s\nscenario-file = %s\n"
                   % (steps, tmp_path / "vtk", os.path.join(ROOT, xml)))
    env = dict(os.environ, LBM_B200_ARITHMETIC="exact", LBM_B200_GPUS=str(gpus), LBM_B200_SPLIT_AXIS=axis)
    dump = tmp_path / "dump.bin"
    r = subprocess.run([exe, str(Q), str(cfg), str(steps), str(dump)], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "HOST_API_CHECK DONE gpus=%d" % gpus in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    dims, f0, f1, kinds, rho, u = read_dump(dump, steps)
    assert dims == (case["xl"], case["yl"], case["zl"], Q)
    want = O.oracle().run(Q, case["xl"], case["yl"], case["zl"], 0.6, case["boxes"], steps, f_init=f0)
    assert np.array_equal(kinds, want["kind"])
    fluid = want["kind"] == O.FLUID
    assert np.array_equal(f1[fluid], want["f"][fluid])
    assert np.array_equal(f1, want["f"])              # boundary cells too (materialised on read-back)
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


def test_create_subdomain_like_the_reference(tmp_path):
    """Domain::create_subdomain (domain.hpp:197-248): extents, copied cells, handlers of the cut / outer faces"""
    exe = str(tmp_path / "subdomain_check")
    compile_cpp(os.path.join(ROOT, "tests", "cpp", "subdomain_check.cpp"), exe)
    cfg = tmp_path / "c.cfg"
    cfg.write_text("tau = 0.6\ntimesteps = 1\ntimesteps-per-plot = 0\noutput-dir = %s\nscenario-file = %s\n"
                   % (tmp_path / "vtk", os.path.join(ROOT, "scenarios", "step_small.xml")))
    r = subprocess.run([exe, str(cfg)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and "SUBDOMAIN_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


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


def test_legacy_vtk_fluid_mask_scenario(exe, tmp_path):
    """domain from a legacy-VTK STRUCTURED_POINTS mask (io/vtk.hpp:94-157, pipe.xml style): solid cells
    become no-slip cells (in both lattices, DESIGN.md deviation 1), then the <boundary> nodes apply."""
    Q, steps = 19, 30
    xl, yl, zl = 26, 9, 11
    zz, yy, xx = np.meshgrid(np.arange(zl), np.arange(yl), np.arange(xl), indexing="ij")
    mask = (((yy - (yl - 1) / 2) ** 2 / 16.0 + (zz - (zl - 1) / 2) ** 2 / 25.0) < 1.0).astype(np.uint8)
    mask[:, :, 10:13] &= (yy[:, :, 10:13] > 2).astype(np.uint8)          # a baffle
    vtk = tmp_path / "mask.vtk"
    with open(vtk, "w") as f:
        f.write("# vtk DataFile Version 2.0\nfluid mask\nASCII\n\nDATASET STRUCTURED_POINTS\n")
        f.write("DIMENSIONS    %d   %d   %d\n\nORIGIN    -3.5   1.25   0.0\nSPACING   2.0   1.0   0.5\n\n" % (xl, yl, zl))
        f.write("POINT_DATA   %d\nSCALARS inputfluidMask unsigned_char\nLOOKUP_TABLE default\n\n" % mask.size)
        flat = mask.reshape(-1)
        for i in range(0, flat.size, 40):
            f.write(" ".join(str(v) for v in flat[i:i + 40]) + " \n")
    xml = tmp_path / "pipe.xml"
    xml.write_text('<?xml version="1.0" ?>\n<scenario name="Pipe flow">\n  <domain vtk-file="%s">\n'
                   '    <boundary extent="z0" condition="noslip" />\n    <boundary extent="zmax" condition="noslip" />\n'
                   '    <boundary extent="x0" condition="inflow" vx="0.03" vy="0" vz="0" />\n'
                   '    <boundary extent="xmax" condition="outflow" />\n    <boundary extent="y0" condition="noslip" />\n'
                   '    <boundary extent="ymax" condition="noslip" />\n  </domain>\n</scenario>\n' % vtk)
    cfg = tmp_path / "run.cfg"
    cfg.write_text("tau = 0.6\ntimesteps = %d\ntimesteps-per-plot = 0\noutput-dir = %s\nscenario-file = %s\n"
                   % (steps, tmp_path / "vtk", xml))
    env = dict(os.environ, LBM_B200_ARITHMETIC="exact")
    dump = tmp_path / "dump.bin"
    r = subprocess.run([exe, str(Q), str(cfg), str(steps), str(dump)], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    dims, f0, f1, kinds, rho, u = read_dump(dump, steps)
    assert dims == (xl, yl, zl, Q)
    want = O.oracle().run(Q, xl, yl, zl, 0.6, O.channel_boxes(xl, yl, zl), steps, f_init=f0, fluid_mask=mask)
    assert np.array_equal(kinds, want["kind"])
    assert np.array_equal(f1, want["f"])
    assert np.array_equal(rho, want["rho"]) and np.array_equal(u, want["u"])
    # origin / spacing of the mask file reach the .vts point coordinates: spacing*i + origin - 1 (io/vtk.hpp:33)
    blob = open(tmp_path / "vtk" / ("Pipe flow.%d.vts" % steps), "rb").read()
    start = blob.index(b"_", blob.index(b"<AppendedData")) + 1
    nv = struct.unpack("<Q", blob[start:start + 8])[0]
    nd = struct.unpack("<Q", blob[start + 8 + nv:start + 16 + nv])[0]
    pts_off = start + 16 + nv + nd
    npts = struct.unpack("<Q", blob[pts_off:pts_off + 8])[0]
    pts = np.frombuffer(blob, dtype=np.float64, count=npts // 8, offset=pts_off + 8).reshape(zl, yl, xl, 3)
    assert np.array_equal(pts[0, 0, 0], [2.0 * 1 - 3.5 - 1, 1.0 * 1 + 1.25 - 1, 0.5 * 1 + 0.0 - 1])
    assert np.array_equal(pts[-1, -1, -1], [2.0 * xl - 3.5 - 1, 1.0 * yl + 1.25 - 1, 0.5 * zl + 0.0 - 1])
