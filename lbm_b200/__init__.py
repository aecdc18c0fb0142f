"""lbm_b200 -- B200-native fused collide-stream sweep, drop-in for the hot path of
hackerbruecke/lbm (D3Q15 / D3Q19 / D3Q27, fp64 BGK).

  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/lbm_b200.h)
  capi.py      ctypes binding of that ABI (used by tests/ and bench.py)
  slabs.py     z-slab decomposition across GPUs, one process per GPU

The reference-facing C++ host surface lives in include/lbm/.
"""
from . import capi  # noqa: F401  (fails loudly when the CUDA library is missing)
from .capi import Domain, LbmError  # noqa: F401
