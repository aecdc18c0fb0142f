"""Multi-slab parity: an N-slab run must equal the single-domain run BITWISE (the split changes no
arithmetic).  Logical slabs on one GPU exercise the same peer-store + hand-shake code that
runs across GPUs; the torchrun tests need >= 2 devices and skip otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle as O
import cases

pytestmark = pytest.mark.gpu
TAU = 0.6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def single(Q, case, steps, exact=True):
    from lbm_b200 import capi
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=exact) as d:
        if case.get("fluid_mask") is not None:
            d.set_fluid_mask(case["fluid_mask"])
        if case["boxes"]:
            d.set_boxes(case["boxes"])
        if case.get("f_init") is not None:
            d.upload(case["f_init"])
        d.step(steps)
        return d.download(), d.macroscopic()


def stacked(Q, case, steps, n_slabs, exact=True, axis=2):
    from lbm_b200.slabs import LocalSlabStack
    st = LocalSlabStack(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], n_slabs, exact=exact,
                        fluid_mask=case.get("fluid_mask"), axis=axis)
    try:
        if case.get("f_init") is not None:
            st.upload(case["f_init"])
        st.step(steps)
        return st.download(), st.macroscopic()
    finally:
        st.close()


def fluid_rows(case, Q):
    kind = O.oracle().run(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 0,
                          fluid_mask=case.get("fluid_mask"), want=("kind",))["kind"]
    return kind == O.FLUID


@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_cavity_slabs_equal_single_domain(Q, n_slabs):
    case = dict(xl=12, yl=10, zl=13, boxes=O.cavity_boxes(12, 10, 13))
    f1, (rho1, u1) = single(Q, case, 40)
    fn, (rhon, un) = stacked(Q, case, 40, n_slabs)
    fl = fluid_rows(case, Q)
    assert np.array_equal(fn[fl], f1[fl])
    assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)


@pytest.mark.parametrize("Q", [19, 27])
def test_channel_with_obstacle_across_the_cut(Q):
    # the no-slip block spans z = 3..8, the cuts of a 4-slab split of zl = 12 fall at 3|4, 6|7, 9|10
    case = cases.channel(20, 8, 12, block=(6, 9, 2, 5, 3, 8))
    f1, (rho1, u1) = single(Q, case, 50)
    fn, (rhon, un) = stacked(Q, case, 50, 4)
    # every cell, obstacle cells next to the cuts included (full edge planes are pushed across the
    # cuts before the read-back pass)
    assert np.array_equal(fn, f1)
    assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)


def test_fast_mode_slabs_equal_single_domain_bitwise():
    case = dict(xl=16, yl=16, zl=16, boxes=O.cavity_boxes(16, 16, 16))
    f1, _ = single(19, case, 30, exact=False)
    fn, _ = stacked(19, case, 30, 2, exact=False)
    fl = fluid_rows(case, 19)
    assert np.array_equal(fn[fl], f1[fl])


def test_random_state_upload_slabs():
    Q = 19
    case = cases.channel(14, 7, 9, block=(4, 6, 2, 4, 2, 6))
    rng = np.random.default_rng(7)
    case["f_init"] = rng.random(((14 + 2) * (7 + 2) * (9 + 2), Q)) * 0.1 + 0.05
    f1, _ = single(Q, case, 12)
    fn, _ = stacked(Q, case, 12, 3)
    fl = fluid_rows(case, Q)
    assert np.array_equal(fn[fl], f1[fl])


def _torchrun(n, *args):
    port = 29533 + (sum(map(ord, "".join(args))) % 40)        # distinct rendezvous ports for back-to-back launches
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + list(args)
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_two_processes_two_gpus(transport):
    from lbm_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun(2, "--transport", transport)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_CHECK OK" in r.stdout


@pytest.mark.parametrize("n_slabs", [2, 3])
def test_periodic_box_over_a_ring_of_slabs(n_slabs):
    """periodic z closed by connecting the last slab to the first: equals the single-domain periodic run"""
    from lbm_b200 import capi
    from lbm_b200.slabs import LocalSlabStack
    Q, n = 19, 12
    case = cases.periodic_random(Q, n=n)
    boxes = cases.periodic_shell_boxes(n, n, n)
    with capi.Domain(Q, n, n, n, TAU, exact=True) as d:
        d.set_boxes(boxes)
        d.upload(case["f_init"])
        d.step(25)
        want = d.download()
    st = LocalSlabStack(Q, n, n, n, TAU, boxes, n_slabs, exact=True, periodic_z=True)
    try:
        st.upload(case["f_init"])
        st.step(25)
        got = st.download()
    finally:
        st.close()
    inner = cases.interior_index(n, n, n)
    assert np.array_equal(got[inner], want[inner])
    # without the ring connection the library refuses instead of computing something else
    lone = capi.Domain(Q, n, n, n, TAU, z_first=1, zl_local=6)
    try:
        lone.set_boxes(boxes)
        with pytest.raises(capi.LbmError):
            lone.step(1)
    finally:
        lone.close()


# ---------------------------------------------------------------------------------------------------
# y-slabs (x-z planes): the same engine with y as the slowest index -- for domains whose z extent is
# shorter than the number of GPUs, or flat ones ("general decomposition")
@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_cavity_y_slabs_equal_single_domain(Q, n_slabs):
    case = dict(xl=12, yl=13, zl=9, boxes=O.cavity_boxes(12, 13, 9))
    f1, (rho1, u1) = single(Q, case, 40)
    fn, (rhon, un) = stacked(Q, case, 40, n_slabs, axis=1)
    assert np.array_equal(fn, f1)
    assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)


@pytest.mark.parametrize("Q", [19, 27])
def test_channel_with_obstacle_across_the_y_cut(Q):
    # the no-slip block spans y = 2..9, the cuts of a 4-slab split of yl = 12 fall at 3|4, 6|7, 9|10
    case = cases.channel(20, 12, 7, block=(6, 9, 2, 9, 2, 5))
    f1, (rho1, u1) = single(Q, case, 50)
    fn, (rhon, un) = stacked(Q, case, 50, 4, axis=1)
    assert np.array_equal(fn, f1)
    assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)


@pytest.mark.parametrize("Q", [15, 19, 27])
def test_flat_shearflow_scenario_in_y_slabs(Q):
    """build/scenarios/shearflow.xml is 8 x 8 x 20 (pressure / outflow in z, free-slip in x, no-slip in y): four
    y-slabs of two rows each -- slabs too thin for the split edge/interior launches"""
    case = cases.shearflow()
    f1, (rho1, u1) = single(Q, case, 60)
    fn, (rhon, un) = stacked(Q, case, 60, 4, axis=1)
    assert np.array_equal(fn, f1)
    assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)
    want = O.oracle().run(Q, 8, 8, 20, TAU, case["boxes"], 60)
    assert np.array_equal(fn, want["f"])


def test_y_slabs_fast_mode_mask_and_random_state():
    Q = 19
    rng = np.random.default_rng(3)
    case = cases.channel(16, 12, 7)
    case["fluid_mask"] = (rng.random((7, 12, 16)) > 0.25).astype(np.uint8)
    case["f_init"] = rng.random(((16 + 2) * (12 + 2) * (7 + 2), Q)) * 0.1 + 0.05
    for exact in (True, False):
        # (12 steps: the random far-from-equilibrium state blows up after ~15, and NaN != NaN)
        f1, (rho1, u1) = single(Q, case, 12, exact=exact)
        fn, (rhon, un) = stacked(Q, case, 12, 3, exact=exact, axis=1)
        assert np.isfinite(f1).all()
        assert np.array_equal(fn, f1)
        assert np.array_equal(rhon, rho1) and np.array_equal(un, u1)


@pytest.mark.parametrize("n_slabs", [2, 3])
def test_periodic_box_over_a_ring_of_y_slabs(n_slabs):
    from lbm_b200 import capi
    from lbm_b200.slabs import LocalSlabStack
    Q, n = 19, 12
    case = cases.periodic_random(Q, n=n)
    boxes = cases.periodic_shell_boxes(n, n, n)
    with capi.Domain(Q, n, n, n, TAU, exact=True) as d:
        d.set_boxes(boxes)
        d.upload(case["f_init"])
        d.step(25)
        want = d.download()
    st = LocalSlabStack(Q, n, n, n, TAU, boxes, n_slabs, exact=True, periodic_z=True, axis=1)
    try:
        st.upload(case["f_init"])
        st.step(25)
        got = st.download()
    finally:
        st.close()
    inner = cases.interior_index(n, n, n)
    assert np.array_equal(got[inner], want[inner])


def test_single_y_slab_handle_equals_plain_domain():
    """a handle created along y that owns every row stores y slowest -- same results, every host array unchanged"""
    from lbm_b200 import capi
    Q = 27
    case = cases.weird(Q)
    f1, (rho1, u1) = single(Q, case, 20)
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=True, axis=1, z_first=1, zl_local=case["yl"]) as d:
        d.set_boxes(case["boxes"])
        d.upload(case["f_init"])
        d.tag_null_cells()               # (the checker tags on its own: set_nonfluid_cells_nullcollide)
        d.step(20)
        f2 = d.download()
        rho2, u2 = d.macroscopic()
        kind = d.kind()
    assert np.array_equal(f2, f1) and np.array_equal(rho2, rho1) and np.array_equal(u2, u1)
    want = O.oracle().run(Q, case["xl"], case["yl"], case["zl"], TAU, case["boxes"], 0, want=("kind",))["kind"]
    assert np.array_equal(kind, want)
