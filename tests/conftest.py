"""pytest configuration: `gpu` marker + shared fixtures.

CPU suite (-m "not gpu"): oracle vs golden vectors / compiled reference, generator KATs,
C-ABI symbol export, host logic.  GPU suite (-m gpu): parity of the CUDA path against the
oracle through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.Oracle("port")


@pytest.fixture(scope="session")
def orc_ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle.Oracle("ref")


@pytest.fixture(scope="session")
def gpu_ctx():
    """Kernel-level context on cuda:0 (fails, never skips, if the device is unusable)."""
    import ctypes as C
    from ugcore_b200 import capi
    ctx = C.c_void_p()
    capi.check(capi.dev.ug4b200_ctx_create(0, None, C.byref(ctx)))
    yield ctx
    capi.dev.ug4b200_ctx_destroy(ctx)
