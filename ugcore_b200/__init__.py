"""ugcore_b200 — B200-native assembled-matrix linear-solve path for UG4/ugcore.

GMG V-cycle (AssembledMultiGridCycle) preconditioning CG / BiCGStab on CRS matrices as
hand-written sm_100a CUDA behind a C ABI (include/ug4b200.h), with a host-side C++ mirror
of ugcore's operator API (csrc/host/) and a descriptor-level driver
(include/ug4b200_solver.h) that this package binds with ctypes.

There is no CPU fallback: loading fails loudly if the native libraries are missing and
every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import importlib

# `problems` (the host-only generator binding) is importable on its own: the CPU reference arm of bench.py and the
# oracle tests use it without pulling the CUDA libraries into their process.  Everything else resolves lazily
# (PEP 562) and loads libug4b200*.so on first access — or raises: there is no fallback.
_LAZY = {"capi": (".capi", None), "solver": (".solver", None), "dist": (".dist", None), "io": (".io", None),
         "problems": (".problems", None),
         "Solver": (".solver", "Solver"), "DeviceBuffer": (".solver", "DeviceBuffer"),
         "device_available": (".solver", "device_available")}

__all__ = ["capi", "Solver", "DeviceBuffer", "device_available", "problems"]


def __getattr__(name):
    if name in _LAZY:
        mod, attr = _LAZY[name]
        m = importlib.import_module(mod, __name__)
        v = m if attr is None else getattr(m, attr)
        globals()[name] = v
        return v
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
