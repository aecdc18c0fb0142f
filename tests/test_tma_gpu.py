"""The TMA-fed sweep (sweep_tma_kernel: persistent blocks, shared-memory ring filled by cp.async.bulk.tensor)
must be indistinguishable from the one-thread-per-cell sweep: bit-identical to the CPU checker in EXACT
arithmetic, bit-identical to sweep_kernel in FAST arithmetic (same finish_cell code), for both box shapes
(128 x 2, 32 x 8), ragged tile edges, walls, obstacles, masks and slab cuts."""
import numpy as np
import pytest

import _oracle as O
import cases
from test_parity_gpu import assert_bitwise, run_cpu, TAU

pytestmark = pytest.mark.gpu


def run(Q, case, steps, exact, tma):
    from lbm_b200 import capi
    with capi.Domain(Q, case["xl"], case["yl"], case["zl"], TAU, exact=exact) as d:
        d.set_sweep_engine(tma=tma, checked=0)
        if case.get("fluid_mask") is not None:
            d.set_fluid_mask(case["fluid_mask"])
        if case["boxes"]:
            d.set_boxes(case["boxes"])
        if case.get("f_init") is not None:
            d.upload(case["f_init"])
        d.step(steps)
        f = d.download()
        rho, u = d.macroscopic()
        n_tma = d.tma_launch_count()
    return dict(f=f, rho=rho, u=u), n_tma


SHAPES = {
    "box32": lambda: cases.channel(40, 12, 10, block=(10, 14, 4, 8, 0, 5)),       # 32 x 8 boxes, ragged in x and y
    "box32wide": lambda: cases.channel(70, 9, 5, block=(30, 40, 2, 6, 1, 3)),     # 32 x 8 boxes, three per row, ragged
    "box128": lambda: cases.channel(150, 5, 3, block=(60, 90, 1, 3, 1, 2)),       # 128 x 2 boxes, second box mostly outside
    "cavity128": lambda: cases.cavity(128),                                           # full tiles, many tiles per block
}


@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("name", ["box32", "box32wide", "box128"])
def test_tma_sweep_exact_is_bit_identical_to_the_checker(Q, name):
    case = SHAPES[name]()
    steps = 30
    cpu = run_cpu(Q, case, steps)
    gpu, n_tma = run(Q, case, steps, exact=True, tma=1)
    assert n_tma == steps - 1          # the first step pulls stored boundary values through sweep_kernel
    assert_bitwise(name + " populations", gpu["f"], cpu["f"])
    assert_bitwise(name + " density", gpu["rho"], cpu["rho"])
    assert_bitwise(name + " velocity", gpu["u"], cpu["u"])


@pytest.mark.parametrize("Q", [15, 19, 27])
@pytest.mark.parametrize("name", sorted(SHAPES))
def test_tma_sweep_fast_equals_direct_sweep_bitwise(Q, name):
    case = SHAPES[name]()
    steps = 12 if name == "cavity128" else 30
    a, n_a = run(Q, case, steps, exact=False, tma=1)
    b, n_b = run(Q, case, steps, exact=False, tma=0)
    assert n_a == steps - 1 and n_b == 0
    assert_bitwise(name + " populations", a["f"], b["f"])
    assert_bitwise(name + " density", a["rho"], b["rho"])


def test_tma_sweep_with_mask_and_random_state():
    for Q in (15, 19, 27):
        case = cases.masked_pipe(48, 14, 12)
        rng = np.random.default_rng(5)
        case["f_init"] = rng.random(((48 + 2) * (14 + 2) * (12 + 2), Q)) * 0.1 + 0.05
        cpu = run_cpu(Q, case, 20)
        gpu, n_tma = run(Q, case, 20, exact=True, tma=1)
        assert n_tma == 19
        assert_bitwise("masked pipe populations", gpu["f"], cpu["f"])


def test_tma_sweep_on_slabs_equals_single_domain():
    from lbm_b200 import capi
    from lbm_b200.slabs import LocalSlabStack
    Q, n, steps = 19, 40, 25
    boxes = O.cavity_boxes(n, n, n)
    with capi.Domain(Q, n, n, n, TAU, exact=True) as d:
        d.set_sweep_engine(tma=0, checked=0)
        d.set_boxes(boxes)
        d.step(steps)
        want = d.download()
    st = LocalSlabStack(Q, n, n, n, TAU, boxes, 3, exact=True)
    try:
        for s in st.slabs:
            s.set_sweep_engine(tma=1, checked=0)
        st.step(steps)
        got = st.download()
        assert sum(s.tma_launch_count() for s in st.slabs) > 0
    finally:
        st.close()
    assert_bitwise("slabs with the TMA sweep", got, want)


def test_tma_sweep_is_opt_in():
    from lbm_b200 import capi
    with capi.Domain(19, 256, 128, 128, TAU) as d:
        d.set_boxes(O.cavity_boxes(256, 128, 128))
        d.step(4)
        assert d.tma_launch_count() == 0
        d.set_sweep_engine(tma=1)
        d.step(4)
        assert d.tma_launch_count() == 4
