"""ugcore_b200 — B200-native assembled-matrix linear-solve path for UG4/ugcore.

GMG V-cycle (AssembledMultiGridCycle) preconditioning CG / BiCGStab on CRS matrices as
hand-written sm_100a CUDA behind a C ABI (include/ug4b200.h), with a host-side C++ mirror
of ugcore's operator API (csrc/host/) and a descriptor-level driver
(include/ug4b200_solver.h) that this package binds with ctypes.

There is no CPU fallback: loading fails loudly if the native libraries are missing and
every compute entry point needs a CUDA device.
"""
from __future__ import annotations

from . import capi  # noqa: F401  (loads the shared libraries or raises)
from .solver import Solver, DeviceBuffer, device_available  # noqa: F401
from . import problems  # noqa: F401

__all__ = ["capi", "Solver", "DeviceBuffer", "device_available", "problems"]
