"""GPU tests of the round-2 work: device-side geometry (boxes / masks / null tags painted by kernels),
handler maps per lattice ("literal" edits, io/vtk.hpp:145-146), read-back across slab cuts after edits,
split read-out, grouped steps, CUDA graphs, checkpoint of both lattices, hand-shake fail-fast."""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle as O
import cases
from test_parity_gpu import assert_bitwise, checker, run_cpu, TAU

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------
# geometry on the device
def test_boxes_painted_on_device_equal_oracle_kind_map():
    from lbm_b200 import capi
    for seed in range(6):
        case = cases.random_scenario(900 + seed, 19)
        want = O.oracle().run(19, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 0,
                              fluid_mask=case.get("fluid_mask"), null_opt=False, want=("kind",))["kind"]
        with capi.Domain(19, case["xl"], case["yl"], case["zl"], TAU) as d:
            if case.get("fluid_mask") is not None:
                d.set_fluid_mask(case["fluid_mask"])
            d.set_boxes(case["boxes"])
            assert np.array_equal(d.kind(), want)
            k, b = d.geometry_planes(1, 2)
            plane = (case["xl"] + 2) * (case["yl"] + 2)
            assert np.array_equal(k, want[plane:3 * plane])
            # handler ids: 0 = the fluid operator, then one per mask / box in the order of the calls
            assert b.max() <= len(case["boxes"]) + 1


def test_handler_table_and_box_ids():
    from lbm_b200 import capi
    n = 6
    with capi.Domain(19, n, n, n, TAU, exact=True) as d:
        d.set_handlers([(capi.FLUID, (0, 0, 0), 1.0), (capi.NOSLIP, (0, 0, 0), 1.0), (capi.MOVINGWALL, (0.05, 0, 0), 1.0)])
        boxes = O.cavity_boxes(n, n, n)
        d.paint_boxes([b[3] for b in boxes], [2 if b[0] == capi.MOVINGWALL else 1 for b in boxes])
        with pytest.raises(capi.LbmError):
            d.paint_boxes([(0, 0, 0, 0, 0, 0)], [7])                       # id outside the table
        with pytest.raises(capi.LbmError):
            d.set_handlers([(capi.FLUID, (0, 0, 0), 1.0)])                 # a table must extend the old one
        with pytest.raises(capi.LbmError):
            d.set_handlers([(capi.NOSLIP, (0, 0, 0), 1.0)] * 3)            # ... without changing kinds
        d.step(20)
        got = d.download()
    assert_bitwise("cavity through paint_boxes", got, run_cpu(19, cases.cavity(n), 20)["f"])


def test_dense_maps_are_checked_on_the_device():
    from lbm_b200 import capi
    n = 5
    kind, bcid, table = capi.paint_boxes(n, n, n, 1, O.cavity_boxes(n, n, n))
    with capi.Domain(19, n, n, n, TAU, exact=True) as d:
        d.set_boxes(O.cavity_boxes(n, n, n))
        before = d.kind()
        bad = kind.copy(); bad[3] = 42
        with pytest.raises(capi.LbmError, match="unknown kind"):
            d.set_geometry(bad, bcid, table)
        bad = bcid.copy(); bad[0] = 99
        with pytest.raises(capi.LbmError, match="outside the table"):
            d.set_geometry(kind, bad, table)
        bad = bcid.copy(); bad[0] = 1                                       # a no-slip cell pointing at the lid handler
        with pytest.raises(capi.LbmError, match="kind differs"):
            d.set_geometry(kind, bad, table)
        bad = kind.copy(); bad[(n + 2) * (n + 2) * 2 + (n + 2) * 2 + 2] = capi.PERIODIC
        with pytest.raises(capi.LbmError, match="ghost-shell"):
            d.set_geometry(bad, bcid, table)
        assert np.array_equal(d.kind(), before)                            # rejected maps change nothing
        d.step(5)                                                          # ... and the old geometry still runs
        d.set_geometry(kind, bcid, table)
        assert np.array_equal(d.kind(), kind)


def test_plane_edits_and_null_tags_report_like_the_reference():
    """set_nonfluid_cells_nullcollide tags the collide field only: Domain::cell() reports NullCollision on even,
    the former handler on odd step counts after the call (domain.hpp:108-109 through cell())"""
    from lbm_b200 import capi
    case = cases.channel(14, 8, 8, block=(4, 9, 2, 6, 2, 6))           # a block thick enough to have buried cells
    with capi.Domain(19, case["xl"], case["yl"], case["zl"], TAU, exact=True) as d:
        d.set_boxes(case["boxes"])
        untagged = d.kind()
        n = d.tag_null_cells()
        assert n == 4 * 3 * 3                                            # the block's core
        tagged = d.kind()
        assert (tagged == capi.NULL).sum() == n and np.array_equal(tagged[tagged != capi.NULL], untagged[tagged != capi.NULL])
        assert d.tag_null_cells() == 0                                   # idempotent
        want = O.oracle().run(19, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 0, want=("kind",))["kind"]
        assert np.array_equal(tagged, want)
        d.step(1)
        assert np.array_equal(d.kind(), untagged)
        d.step(1)
        assert np.array_equal(d.kind(), tagged)
        d.step(7)
        ref = checker().run(19, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 9)
        assert_bitwise("populations with tags", d.download(), ref["f"])
        assert np.array_equal(d.kind(), ref["kind"])
        # edit one plane through dense maps: turn a fluid cell into a no-slip cell (handler id 1 = first box)
        k, b = d.geometry_planes(3, 1)
        idx = 4 * (case["xl"] + 2) + 2
        assert k[idx] == capi.FLUID
        k[idx], b[idx] = capi.NOSLIP, 1
        d.set_geometry_planes(k, b, 3, 1)
        k2, b2 = d.geometry_planes(3, 1)
        assert k2[idx] == capi.NOSLIP and b2[idx] == 1


# ---------------------------------------------------------------------------------------------------
# literal mask semantics (io/vtk.hpp:145-146): handler on the collide field only
@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("null_opt", [False, True])
def test_literal_mask_is_bit_identical_to_the_reference(Q, null_opt):
    from lbm_b200 import capi
    case = cases.masked_pipe()
    steps = 21
    rng = np.random.default_rng(3 + Q)
    _, w = O.oracle().model(Q)
    n_all = (case["xl"] + 2) * (case["yl"] + 2) * (case["zl"] + 2)
    f0 = np.tile(w, (n_all, 1)) * (1 + 0.05 * rng.standard_normal((n_all, Q)))
    ref = checker().run(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], steps, fluid_mask=case["fluid_mask"],
                        mask_literal=True, null_opt=null_opt, f_init=f0)
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True) as d:
        d.set_fluid_mask(case["fluid_mask"], literal=True)
        d.set_boxes(case["boxes"])
        d.upload(f0)
        if null_opt:
            d.tag_null_cells(literal=True)
        d.step(steps)
        f = d.download()
        rho, u = d.macroscopic()
        kind = d.kind()
    assert_bitwise("literal mask populations", f, ref["f"])
    assert_bitwise("literal mask density", rho, ref["rho"])
    assert_bitwise("literal mask velocity", u, ref["u"])
    assert np.array_equal(kind, ref["kind"])
    # and it really is another algorithm than the both-lattices default
    both = checker().run(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], steps, fluid_mask=case["fluid_mask"],
                         mask_literal=False, null_opt=null_opt, f_init=f0)
    assert not np.array_equal(both["rho"], ref["rho"])


def test_literal_mask_random_scenarios():
    from lbm_b200 import capi
    done = 0
    for seed in range(40):
        case = cases.random_scenario(7000 + seed, 19)
        if case.get("fluid_mask") is None:
            continue
        steps = 5 + seed % 4
        ref = checker().run(19, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], steps, fluid_mask=case["fluid_mask"],
                            mask_literal=True, f_init=case["f_init"])
        with capi.Domain(19, case["xl"], case["yl"], case["zl"], TAU, exact=True) as d:
            d.set_fluid_mask(case["fluid_mask"], literal=True)
            d.set_boxes(case["boxes"])
            d.upload(case["f_init"])
            d.tag_null_cells(literal=True)
            d.step(steps)
            assert_bitwise("seed %d populations" % seed, d.download(), ref["f"])
            rho, u = d.macroscopic()
            assert_bitwise("seed %d density" % seed, rho, ref["rho"])
            assert np.array_equal(d.kind(), ref["kind"])
        done += 1
    assert done >= 10


def read_legacy_vtk_mask(path):
    """minimal reader of the fixture (STRUCTURED_POINTS, ASCII, one unsigned_char scalar)"""
    with open(path) as fh:
        text = fh.read()
    head, data = text.split("LOOKUP_TABLE default")
    dims = [int(v) for v in head.split("DIMENSIONS")[1].split()[:3]]
    mask = np.array(data.split(), dtype=np.uint8)
    assert mask.size == dims[0] * dims[1] * dims[2]
    return dims, mask.reshape(dims[2], dims[1], dims[0])


@pytest.mark.parametrize("literal", [True, False])
def test_reference_pipe_scenario_bitwise(literal):
    """build/scenarios/pipe.xml + pipe.vtk (250x54x54, 408838 fluid cells) as the reference runs it
    (src/main.cpp:37-52): mask -> boundaries -> set_nonfluid_cells_nullcollide -> steps"""
    from lbm_b200 import capi
    (xl, yl, zl), mask = read_legacy_vtk_mask(os.path.join(os.path.dirname(__file__), "golden", "pipe.vtk"))
    assert (xl, yl, zl) == (250, 54, 54) and int(mask.sum()) == 408838
    Q, steps = 19, 12
    boxes = O.channel_boxes(xl, yl, zl)
    ref = checker().run(Q, xl, yl, zl, TAU, boxes, steps, fluid_mask=mask, mask_literal=literal)
    with capi.Domain(Q, xl, yl, zl, TAU, exact=True) as d:
        d.set_fluid_mask(mask, literal=literal)
        d.set_boxes(boxes)
        d.tag_null_cells(literal=literal)
        d.step(steps)
        assert_bitwise("pipe populations", d.download(), ref["f"])
        rho, u = d.macroscopic()
    assert_bitwise("pipe density", rho, ref["rho"])
    assert_bitwise("pipe velocity", u, ref["u"])


# ---------------------------------------------------------------------------------------------------
# multi-slab: edits after read-back (kernels.cuh materialize), masks across cuts, grouped steps
def test_mask_across_slab_cuts():
    from lbm_b200.slabs import LocalSlabStack
    from test_multislab_gpu import single
    Q = 19
    rng = np.random.default_rng(11)
    case = cases.channel(16, 7, 12)
    case["fluid_mask"] = (rng.random((12, 7, 16)) > 0.25).astype(np.uint8)
    f1, (rho1, u1) = single(Q, case, 30)
    st = LocalSlabStack(Q, 16, 7, 12, TAU, case["boxes"], 4, exact=True, fluid_mask=case["fluid_mask"])
    try:
        st.step(30)
        fn = st.download()
        rhon, un = st.macroscopic()
    finally:
        st.close()
    assert_bitwise("masked slabs populations", fn, f1)
    assert_bitwise("masked slabs density", rhon, rho1)


@pytest.mark.parametrize("edit", ["upload_planes", "set_boxes", "checkpoint"])
def test_edits_after_readback_on_slabs_equal_single_domain(edit, tmp_path):
    """After a read-back the ghost planes hold the neighbours' pushed edge planes; wall cells inside them must
    be materialised for OUR links before the next sweep pulls stored values (`first`)."""
    from lbm_b200 import capi
    from lbm_b200.slabs import LocalSlabStack
    Q, (xl, yl, zl) = 19, (10, 9, 12)
    boxes = O.cavity_boxes(xl, yl, zl)
    extra = [(O.NOSLIP, (0.0, 0.0, 0.0), 1.0, (4, 6, 3, 5, 5, 8))]       # straddles the cut of a 2-slab split (6|7)

    def scenario(runner_step, download, upload_plane, set_boxes, save, load):
        runner_step(14)
        mid = download()
        if edit == "upload_planes":
            plane = (xl + 2) * (yl + 2)
            f = mid.reshape(zl + 2, plane, Q).copy()
            f[6, 5 * (xl + 2) + 5] *= 1.01                                 # one fluid cell next to the cut
            upload_plane(f[6], 6)
        elif edit == "set_boxes":
            set_boxes(extra)
        else:
            save()
            runner_step(3)                                                 # wander off, then come back
            load()
        runner_step(9)
        return download()

    with capi.Domain(Q, xl, yl, zl, TAU, exact=True) as d:
        d.set_boxes(boxes)
        ck = str(tmp_path / "single.ck")
        want = scenario(d.step, d.download, lambda p, z: d.upload_planes(p, z, 1), d.set_boxes,
                        lambda: d.save_checkpoint(ck), lambda: d.load_checkpoint(ck))
    st = LocalSlabStack(Q, xl, yl, zl, TAU, boxes, 2, exact=True)
    try:
        def up(p, z):
            for (zf, nz), s in zip(st.ranges, st.slabs):
                if zf - 1 <= z <= zf + nz:
                    s.upload_planes(p, z - (zf - 1), 1)

        def sb(b):
            for s in st.slabs:
                s.set_boxes(b)
        paths = [str(tmp_path / ("slab%d.ck" % i)) for i in range(2)]
        got = scenario(st.step, st.download, up, sb,
                       lambda: [s.save_checkpoint(p) for s, p in zip(st.slabs, paths)],
                       lambda: [s.load_checkpoint(p) for s, p in zip(st.slabs, paths)])
    finally:
        st.close()
    kind = O.oracle().run(Q, xl, yl, zl, TAU, boxes + (extra if edit == "set_boxes" else []), 0, want=("kind",))["kind"]
    assert_bitwise("fluid cells after " + edit, got[kind == O.FLUID], want[kind == O.FLUID])
    assert_bitwise("all cells after " + edit, got, want)


def test_step_group_and_graphs_equal_plain_steps():
    from lbm_b200 import capi
    from lbm_b200.slabs import LocalSlabStack
    Q, case = 27, cases.channel(20, 8, 12, block=(6, 9, 2, 5, 3, 8))
    steps = 53                                                             # 3 graph runs of 16 + 5 single steps
    want = run_cpu(Q, case, steps)
    for graphs in (0, 1):
        with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True) as d:
            d.set_graphs(graphs)
            d.set_boxes(case["boxes"])
            l0 = d.launch_count()
            d.step(steps)
            assert d.steps_done() == steps
            assert d.launch_count() - l0 >= steps
            assert_bitwise("graphs=%d" % graphs, d.download(), want["f"])
        st = LocalSlabStack(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 3, exact=True)
        try:
            for s in st.slabs:
                s.set_graphs(graphs)
            st.step(steps)                                                 # lbm_b200_step_group
            assert_bitwise("slabs, graphs=%d" % graphs, st.download(), want["f"])
        finally:
            st.close()


def test_hand_shake_timeout_fails_fast(monkeypatch):
    """a neighbour that never steps: the waiting slab gives up after the (shortened) timeout, and every later
    step is refused at once instead of sweeping on stale ghost planes"""
    import subprocess
    import sys
    code = r'''
import sys, time
sys.path.insert(0, %r)
from lbm_b200 import capi
from lbm_b200.slabs import LocalSlabStack
import _oracle as O
st = LocalSlabStack(19, 8, 8, 8, 0.6, O.cavity_boxes(8, 8, 8), 2)
a = st.slabs[0]
a.step(1)            # epoch 0: nothing to wait for
a.step(1)            # waits for the neighbour's first sweep, which never comes
try:
    a.sync()
    print("NO ERROR")
except capi.LbmError as e:
    print("SYNC", e.code)
t0 = time.time()
try:
    a.step(1)
    print("STEP ACCEPTED")
except capi.LbmError as e:
    print("STEP", e.code, "%%.3f" %% (time.time() - t0))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LBM_B200_HALO_TIMEOUT_MS="300", PYTHONPATH=os.path.join(root, "tests"))
    r = subprocess.run([sys.executable, "-c", code % root], capture_output=True, text=True, env=env, timeout=300)
    assert "SYNC -5" in r.stdout, r.stdout + r.stderr
    assert "STEP -5" in r.stdout and "STEP ACCEPTED" not in r.stdout, r.stdout + r.stderr


# ---------------------------------------------------------------------------------------------------
# read-out and checkpoints
def test_split_readout_overlaps_with_steps():
    from lbm_b200 import capi
    Q, n = 19, 20
    boxes = O.cavity_boxes(n, n, n)
    with capi.Domain(Q, n, n, n, TAU, exact=True) as d:
        d.set_boxes(boxes)
        d.step(15)
        rho_want, u_want = d.macroscopic()
        rho = capi.HostBuffer(n ** 3)
        u = capi.HostBuffer(3 * n ** 3)
        d.macroscopic_begin(rho.address, u.address)
        d.step(10)                                   # the snapshot is of step 15, whatever runs afterwards
        d.macroscopic_end()
        assert_bitwise("snapshot density", rho.array.reshape(n, n, n), rho_want)
        assert_bitwise("snapshot velocity", u.array.reshape(n, n, n, 3), u_want)
        d.macroscopic_begin(rho.address, None)       # a second read-out waits for the first implicitly
        d.macroscopic_begin(rho.address, u.address)
        d.macroscopic_end()
        want = run_cpu(Q, cases.cavity(n), 25)
        assert_bitwise("later density", rho.array.reshape(n, n, n), want["rho"])
        assert_bitwise("later velocity", u.array.reshape(n, n, n, 3), want["u"])
        rho.free(); u.free()
    assert capi.lib.lbm_b200_host_free(C.c_void_p(12345)) == -1          # not one of ours


def test_checkpoint_keeps_both_lattices_and_rejects_mismatches(tmp_path):
    """an uncovered ghost shell is collided in place in BOTH lattices (domain.hpp:147-155), so the stream field
    carries state; mode and geometry are recorded"""
    from lbm_b200 import capi
    Q = 19
    case = cases.weird(Q)
    ck = str(tmp_path / "w.ck")
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True) as a:
        a.set_boxes(case["boxes"])
        a.upload(case["f_init"])
        a.step(7)
        a.save_checkpoint(ck)
        a.step(8)
        end = a.download()
    assert_bitwise("uninterrupted vs oracle", end, run_cpu(Q, case, 15)["f"])
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True) as b:
        b.set_boxes(case["boxes"])
        b.load_checkpoint(ck)
        b.step(8)
        assert_bitwise("restart with an uncovered shell", b.download(), end)
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=False) as c:
        c.set_boxes(case["boxes"])
        with pytest.raises(capi.LbmError, match="arithmetic"):
            c.load_checkpoint(ck)
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True) as e:
        e.set_boxes(case["boxes"][:-1])
        with pytest.raises(capi.LbmError, match="geometry"):
            e.load_checkpoint(ck)


def test_reciprocal_division_is_ieee_division():
    """kernels.cuh div_rcp (the bit-identical mode's quotients by C_S^2, 2 C_S^4, 2 C_S^2, tau and rho from a correctly
    rounded reciprocal + two FMA corrections) == __ddiv_rn on 5 x 2*10^7 operands per call, several tau / seeds"""
    from lbm_b200 import capi
    for seed, tau in ((1, 0.6), (2, 0.51), (3, 1.0), (4, 1.9999), (5, 0.75)):
        assert capi.selftest_division(20_000_000, seed, tau) == 0


# ---------------------------------------------------------------------------------------------------
# guessed pulls of x-face cells (kernels.cuh SWEEP_XFACE): right guesses, wrong guesses, no guess
def _xface_case(Q, seed=5):
    xl, yl, zl = 12, 14, 12     # 120 cells per x face away from the edges; the patches spoil about a quarter of them
    rng = np.random.default_rng(seed)
    boxes = O.cavity_boxes(xl, yl, zl) + O.face_boxes(xl, yl, zl, [
        ((0, 0, 3, 4, 2, 3), O.MOVINGWALL, (0.0, 0.02, -0.01)),      # a patch of another handler on the x = 0 face
        ((xl + 1, xl + 1, 2, 4, 4, 7), O.INFLOW, (-0.02, 0.0, 0.0)),  # ... and on the x = xl+1 face
        ((1, 2, 7, 8, 3, 4), O.NOSLIP),                               # an obstacle that touches the x = 1 cells
    ])
    f0 = rng.random(((xl + 2) * (yl + 2) * (zl + 2), Q)) * 0.1 + 0.05
    return dict(xl=xl, yl=yl, zl=zl, boxes=boxes, f_init=f0)


@pytest.mark.parametrize("Q", [15, 19, 27])
def test_xface_guesses_bitwise_with_mixed_faces(Q, monkeypatch):
    """most x-face cells see a NoSlipBoundary (the guess), some see a moving wall / an inflow / an obstacle:
    both the confirmed and the refuted guesses must give the reference's populations, with and without the mode"""
    from test_parity_gpu import run_gpu
    case = _xface_case(Q)
    cpu = run_cpu(Q, case, 7)
    for xface in ("1", "0"):
        monkeypatch.setenv("LBM_B200_XFACE", xface)
        gpu = run_gpu(Q, case, 7, exact=True)
        assert_bitwise("x-face guesses (LBM_B200_XFACE=%s)" % xface, gpu["f"], cpu["f"])
        assert_bitwise("density", gpu["rho"], cpu["rho"])


@pytest.mark.parametrize("Q", [15, 19, 27])
def test_xface_guesses_periodic_with_patches(Q, monkeypatch):
    """x periodic (the guess) except for a wall patch on each x face, y / z walls: guessed == unguessed, bit for bit,
    in both arithmetic modes"""
    from lbm_b200 import capi
    xl, yl, zl = 16, 12, 11
    rng = np.random.default_rng(11)
    boxes = O.face_boxes(xl, yl, zl, [("z0", O.NOSLIP), ("zmax", O.MOVINGWALL, (0.03, 0.0, 0.0)), ("y0", O.NOSLIP),
                                      ("ymax", O.NOSLIP), ("x0", O.PERIODIC), ("xmax", O.PERIODIC),
                                      ((0, 0, 3, 5, 3, 5), O.NOSLIP), ((xl + 1, xl + 1, 2, 3, 2, 6), O.NOSLIP)])
    f0 = rng.random(((xl + 2) * (yl + 2) * (zl + 2), Q)) * 0.1 + 0.05
    for exact in (True, False):
        got = {}
        for xface in ("1", "0"):
            monkeypatch.setenv("LBM_B200_XFACE", xface)
            with capi.Domain(Q, xl, yl, zl, TAU, exact=exact) as d:
                d.set_boxes(boxes)
                d.upload(f0)
                d.step(9)
                got[xface] = d.download()
        inner = cases.interior_index(xl, yl, zl)
        assert_bitwise("periodic x faces with patches, exact=%s" % exact, got["1"][inner], got["0"][inner])
