"""CPU: the Python scenario / config readers (lbm_b200/scenario.py) agree with the hand-built box lists the
other tests use, reproduce the reference's error messages, and drive the oracle from the shipped XML files."""
import os

import numpy as np
import pytest

import _oracle as O
import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shipped_scenarios_parse_to_the_expected_boxes(built):
    import scenario_reader as scenario
    sc = scenario.load_scenario(os.path.join(ROOT, "scenarios", "cavity64.xml"))
    assert (sc["name"], sc["xl"], sc["yl"], sc["zl"]) == ("Cavity64", 64, 64, 64)
    assert sc["boxes"] == O.cavity_boxes(64, 64, 64)
    sc = scenario.load_scenario(os.path.join(ROOT, "scenarios", "channel_d3q27.xml"))
    assert sc["boxes"] == O.channel_boxes(1024, 256, 256)
    sc = scenario.load_scenario(os.path.join(ROOT, "scenarios", "shear_small.xml"))
    assert sc["boxes"] == cases.shearflow()["boxes"]
    sc = scenario.load_scenario(os.path.join(ROOT, "scenarios", "step_small.xml"))
    assert sc["boxes"][0] == (O.INFLOW, (0.0, 0.0, 0.04), 1.0, (11, 21, 0, 13, 0, 0))
    assert sc["boxes"][1] == (O.NOSLIP, (0.0, 0.0, 0.0), 1.0, (0, 10, 0, 13, 0, 12))
    # the oracle runs straight from the parsed scenario
    out = O.oracle().run(19, sc["xl"], sc["yl"], sc["zl"], 0.6, sc["boxes"], 3, want=("kind", "rho"))
    assert out["kind"].max() == O.INFLOW and np.isfinite(out["rho"]).all()


def test_reference_scenarios_parse_when_present(built):
    import scenario_reader as scenario
    base = "/root/reference/build/scenarios"
    if not os.path.isdir(base):
        pytest.skip("/root/reference absent")
    sc = scenario.load_scenario(os.path.join(base, "cavity.xml"))
    assert (sc["xl"], sc["yl"], sc["zl"]) == (40, 40, 40) and sc["boxes"] == O.cavity_boxes(40, 40, 40)
    sc = scenario.load_scenario(os.path.join(base, "shearflow.xml"))
    assert sc["boxes"] == cases.shearflow()["boxes"]
    sc = scenario.load_scenario(os.path.join(base, "step.xml"))
    assert sc["boxes"] == cases.step_flow()["boxes"]
    cwd = os.getcwd()
    os.chdir("/root/reference/build")         # pipe.xml names its mask relative to the build directory
    try:
        sc = scenario.load_scenario(os.path.join(base, "pipe.xml"))
    finally:
        os.chdir(cwd)
    assert (sc["xl"], sc["yl"], sc["zl"]) == (250, 54, 54)
    assert int(sc["fluid_mask"].sum()) == 408838 and sc["fluid_mask"].size == 729000
    assert sc["boxes"] == O.channel_boxes(250, 54, 54)
    cfg = scenario.load_config("/root/reference/build/config.cfg")
    assert cfg["tau"] == 0.6 and cfg["timesteps"] == 1000 and cfg["timesteps-per-plot"] == 1
    assert cfg["scenario-file"] == "scenarios/cavity.xml" and cfg["collision-model"] == "bgk"


def test_error_messages_follow_the_reference(built, tmp_path):
    import scenario_reader as scenario
    def bad(text, needle):
        p = tmp_path / "s.xml"
        p.write_text(text)
        with pytest.raises(scenario.ScenarioError) as e:
            scenario.load_scenario(str(p))
        assert needle in str(e.value)
    bad("<nonsense", "could not be read properly!")
    bad('<other name="x"/>', "Scenario node missing!")
    bad('<scenario><domain xl="2" yl="2" zl="2"/></scenario>', "Scenario name is missing!")
    bad('<scenario name="s"></scenario>', "Domain node is missing!")
    bad('<scenario name="s"><domain xl="2" yl="2"/></scenario>', "Neither vtk-file nor xl/yl/zl attribute provided")
    bad('<scenario name="s"><domain xl="2" yl="2" zl="2"><boundary condition="noslip"/></domain></scenario>', 'Missing attribute "extent"')
    bad('<scenario name="s"><domain xl="2" yl="2" zl="2"><boundary extent="z0" condition="slippery"/></domain></scenario>', "slippery boundary condition not supported!")
    bad('<scenario name="s"><domain xl="2" yl="2" zl="2"><boundary extent="0 1 2" condition="noslip"/></domain></scenario>', "is not complete! Must be six values!")
    bad('<scenario name="s"><domain xl="2" yl="2" zl="2"><boundary extent="z0" condition="movingwall" vx="1"/></domain></scenario>', 'Missing attribute "vy"')
    bad('<scenario name="s"><domain xl="2" yl="2" zl="2"><boundary extent="z0" condition="pressure"/></domain></scenario>', 'Missing attribute "rho-in"')
    cfg = tmp_path / "c.cfg"
    cfg.write_text("tau = 0.6\ntimesteps = 5\n")
    with pytest.raises(scenario.ScenarioError):
        scenario.load_config(str(cfg))
    cfg.write_text("tau = 0.4\ntimesteps = 5\ntimesteps-per-plot = 0\nscenario-file = x.xml\n")
    with pytest.raises(scenario.ScenarioError):
        scenario.load_config(str(cfg))
    for name in ("cavity64.cfg", "cavity512.cfg"):
        c = scenario.load_config(os.path.join(ROOT, "configs", name))
        assert c["tau"] == 0.6 and os.path.exists(os.path.join(ROOT, c["scenario-file"]))


def test_reference_pipe_scenario_oracle_equals_compiled_reference(built):
    """the reference's own pipe fixture (250 x 54 x 54 mask, inflow/outflow): restatement == reference's code,
    with the mask tagged on both lattices and with the reference's literal collide-field-only tagging"""
    import scenario_reader as scenario
    ref = O.ref()
    if ref is None or not os.path.isdir("/root/reference/build/scenarios"):
        pytest.skip("needs /root/reference and oracle/_ref")
    cwd = os.getcwd()
    os.chdir("/root/reference/build")
    try:
        sc = scenario.load_scenario("scenarios/pipe.xml")
    finally:
        os.chdir(cwd)
    for literal in (False, True):
        kw = dict(fluid_mask=sc["fluid_mask"], mask_literal=literal, want=("f", "kind"))
        a = ref.run(19, sc["xl"], sc["yl"], sc["zl"], 0.6, sc["boxes"], 4, **kw)
        b = O.oracle().run(19, sc["xl"], sc["yl"], sc["zl"], 0.6, sc["boxes"], 4, **kw)
        assert np.array_equal(a["kind"], b["kind"])
        assert np.array_equal(a["f"].view(np.uint64), b["f"].view(np.uint64))
