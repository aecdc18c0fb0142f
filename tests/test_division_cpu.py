"""CPU check of the division-free quotients of the bit-identical mode (kernels.cuh div_rcp): the same
sequence in C (fma) against IEEE division for the reference's constant divisors and random ones -- tools/selftest/div_check.c.
The device-side counterpart is tests/test_round2_gpu.py::test_reciprocal_division_is_ieee_division."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_fma_corrections_give_the_ieee_quotient(tmp_path):
    try:
        if " fma " not in open("/proc/cpuinfo").read().replace("\n", " ") + " ":
            pytest.skip("host CPU without FMA instructions")
    except OSError:
        pytest.skip("cannot read /proc/cpuinfo")
    exe = str(tmp_path / "div_check")
    src = os.path.join(ROOT, "tools", "selftest", "div_check.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if "bad5=" in l]
    assert len(lines) >= 12, out                      # 11 constant divisors + the random-divisor block
    assert all(re.search(r"bad5=0\b", l) for l in lines), out
