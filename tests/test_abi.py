"""CPU: the C-ABI library loads, exports every symbol include/lbm_b200.h declares, its host-only entry
points agree with the oracle, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from lbm_b200 import capi
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(capi.lib, s), "missing export: " + s
    assert sorted(capi.SYMBOLS) == syms
    assert capi.lib.lbm_b200_abi_version() == 2


@pytest.mark.parametrize("Q", [15, 19, 27])
def test_model_tables_match_oracle(built, Q):
    from lbm_b200 import capi
    c, w = capi.model(Q)
    oc, ow = O.oracle().model(Q)
    assert np.array_equal(c, oc) and np.array_equal(w.view(np.uint64), ow.view(np.uint64))
    for q in range(Q):
        assert capi.velocity_index(Q, int(c[q, 0]), int(c[q, 1]), int(c[q, 2])) == q
        assert capi.lib.lbm_b200_model_inv(Q, q) == Q - 1 - q
    if Q != 27:
        assert capi.velocity_index(Q, 1, 1, 0 if Q == 15 else 1) == -1


def test_bad_arguments_and_missing_gpu_fail_loudly(built):
    from lbm_b200 import capi
    h = C.c_void_p()
    assert capi.lib.lbm_b200_create(C.byref(h), 17, 8, 8, 8, 0.6, -1) == -1           # EINVAL
    assert b"15, 19 or 27" in capi.lib.lbm_b200_last_error()
    assert capi.lib.lbm_b200_create(C.byref(h), 19, 0, 8, 8, 0.6, -1) == -1
    assert capi.lib.lbm_b200_create(C.byref(h), 19, 8, 8, 8, -1.0, -1) == -1
    assert capi.lib.lbm_b200_create_slab(C.byref(h), 19, 8, 8, 8, 5, 6, 0.6, -1) == -1  # slab outside the domain
    assert capi.lib.lbm_b200_step(None, 1) == -1
    assert capi.lib.lbm_b200_model(11, None, None) == -1
    if capi.device_count() == 0:
        with pytest.raises(capi.LbmError) as e:
            capi.Domain(19, 8, 8, 8, 0.6)
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "lbm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                for needle in ("_oracle", "liblbm_oracle", "libref_lbm", "oracle/", "oracle.h", "oracle_run", "ref_run"):
                    assert needle not in text, "%s mentions %s" % (os.path.join(base, f), needle)
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "_oracle" not in open(os.path.join(base, f)).read()
