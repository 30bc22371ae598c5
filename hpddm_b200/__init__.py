"""hpddm_b200 -- B200-native (sm_100a) implementation of HPDDM's one-/two-level
RAS preconditioner-apply hot path.

Layout
  csrc/      CUDA kernels + the C ABI (include/hpddm_b200.h) -> lib/libhpddm_b200.so
  host/      header-only C++ mirror of the reference surface (HPDDM::Schwarz, SUBDOMAIN plugin)
  capi.py    ctypes binding of the C ABI
  schwarz.py Python mirror of HPDDM::Schwarz over the C ABI (the reference ships the same
             kind of binding: interface/hpddm.py)
"""
from . import capi  # noqa: F401
from .schwarz import Decomposition, KrylovOperator, Schwarz  # noqa: F401
