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


@pytest.fixture(scope="session", params=["value_indexed", "plain", "unbatched"])
def gpu_ctx(request):
    """Kernel-level context on cuda:0 (fails, never skips, if the device is unusable).  Every
    kernel test runs twice: with the value-indexed entry stream (u16 dictionary index + u16 column
    offset, chosen automatically for matrices with repeated values) and with the plain
    8-byte-value / 4-byte-column stream forced (UG4B200_NO_COMPRESS=1).  Small operands are recorded
    and executed by the one-cluster batch kernel (csrc/batch.cu) in both; "unbatched" turns that off
    (UG4B200_BATCH=0) so the stand-alone kernels keep their coverage on the same small cases."""
    import ctypes as C
    from ugcore_b200 import capi
    ctx = C.c_void_p()
    if request.param == "plain":
        os.environ["UG4B200_NO_COMPRESS"] = "1"
    if request.param == "unbatched":
        os.environ["UG4B200_BATCH"] = "0"
    try:
        capi.check(capi.dev.ug4b200_ctx_create(0, None, C.byref(ctx)))
    finally:
        os.environ.pop("UG4B200_NO_COMPRESS", None)
        os.environ.pop("UG4B200_BATCH", None)
    yield ctx
    capi.dev.ug4b200_ctx_destroy(ctx)
