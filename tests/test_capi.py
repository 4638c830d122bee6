"""CPU suite: the C-ABI libraries load, export every symbol the headers declare, and fail
loudly (no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ug4b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from ugcore_b200 import capi
    dev_syms = _declared("ug4b200.h")
    host_syms = [s for s in _declared("ug4b200_solver.h") if s not in dev_syms]
    assert len(dev_syms) > 60 and len(host_syms) >= 18
    for s in dev_syms:
        assert hasattr(capi.dev, s), f"libug4b200.so does not export {s}"
        assert s in capi.DEV_API, f"capi.py has no prototype for {s}"
    for s in host_syms:
        assert hasattr(capi.host, s), f"libug4b200_host.so does not export {s}"
        assert s in capi.HOST_API, f"capi.py has no prototype for {s}"
    # and nothing is bound that the headers do not declare
    assert set(capi.DEV_API) <= set(dev_syms)
    assert set(capi.HOST_API) <= set(host_syms)


def test_struct_layouts_match_headers():
    """ctypes mirrors must have the C layout (compiled probe)."""
    from ugcore_b200 import capi
    probe = r'''
    #include "ug4b200_solver.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ug4b200_conv_state), offsetof(ug4b200_conv_state, done),
      sizeof(ug4b200_fin), sizeof(ug4b200_coef), sizeof(ug4b200_matrix_info), sizeof(ug4b200_solver_desc),
      offsetof(ug4b200_solver_desc, flags)); return 0; }'''
    exe = "/tmp/ug4b200_layout_probe"
    subprocess.run(["gcc", "-x", "c", "-", "-I" + os.path.join(ROOT, "include"), "-o", exe], input=probe.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    got = [int(v) for v in out]
    exp = [C.sizeof(capi.ConvState), capi.ConvState.done.offset, C.sizeof(capi.Fin), C.sizeof(capi.Coef),
           C.sizeof(capi.MatrixInfo), C.sizeof(capi.SolverDesc), capi.SolverDesc.flags.offset]
    assert got == exp


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return not os.path.exists("/dev/nvidia0")


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a host without a GPU")
def test_no_cpu_fallback_without_device():
    from ugcore_b200 import capi
    import ugcore_b200 as ug
    ctx = C.c_void_p()
    rc = capi.dev.ug4b200_ctx_create(0, None, C.byref(ctx))
    assert rc != 0
    assert b"no CPU fallback" in capi.dev.ug4b200_last_error(None)
    assert not ug.device_available()
    from ugcore_b200 import problems as pr
    with pytest.raises(capi.UG4B200Error):
        ug.Solver.from_problem({"type": "cg"}, pr.Problem(dim=2, num_refs=2))


def test_host_coloring_helpers():
    """ug4b200_color_greedy / ug4b200_color_check are host functions: usable without a GPU."""
    import numpy as np
    from ugcore_b200 import capi, problems as pr
    A = pr.Problem(dim=3, num_refs=3).matrix()
    color = np.zeros(A.nrows, np.int32)
    nc = C.c_int()
    capi.dev.ug4b200_color_greedy(A.nrows, A.rowptr.ctypes.data_as(C.c_void_p), A.cols.ctypes.data_as(C.c_void_p),
                                  color.ctypes.data_as(C.c_void_p), C.byref(nc))
    assert nc.value == 8  # parity colouring of the 27-point stencil
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
    off = rows != A.cols
    assert np.all(color[rows[off]] != color[A.cols[off]])
    # the natural order is NOT a valid single colour
    cp = np.array([0, A.nrows], np.int64)
    assert capi.dev.ug4b200_color_check(A.nrows, A.rowptr.ctypes.data_as(C.c_void_p), A.cols.ctypes.data_as(C.c_void_p),
                                        1, cp.ctypes.data_as(C.c_void_p)) != 0


def test_product_does_not_touch_the_oracle():
    """Nothing under ugcore_b200/ may import, link or reference oracle/ (③)."""
    pkg = os.path.join(ROOT, "ugcore_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(base, f)


def test_descriptor_defaults_follow_solver_util_lua():
    """The descriptor layer plays util.solver.CreateSolver: absent entries take util.solver.defaults
    (scripts/util/solver_util.lua:423-575, GMRES(5) at :669-670) — for the product and for the oracle alike."""
    import oracle
    from ugcore_b200 import solver as S
    for make, ILU_, GS_, JAC_ in ((lambda d: S.make_desc(d), 6, 2, 1), (oracle.make_desc, 6, 2, 1)):
        d = make({"type": "cg"})
        assert d.precond == ILU_ and d.ilu_beta == 0.0 and d.max_steps == 100 and d.min_defect == 1e-12 and d.rel_reduction == 1e-6
        assert make({"type": "bicgstab", "precond": None}).precond == 0 and make({"type": "linear", "precond": "none"}).precond == 0
        assert make({"type": "gmres"}).restart == 5
        g = make({"type": "linear", "precond": {"type": "gmg", "topLevel": 3}})
        assert (g.nu1, g.nu2, g.smoother, g.cycle, g.base_lev, g.base_solver) == (3, 3, GS_, 1, 0, 3)
        j = make({"type": "cg", "precond": "jac"})
        assert j.precond == JAC_ and j.damp == 0.66
        assert make({"type": "cg", "precond": {"type": "jac", "damping": 0.5}}).damp == 0.5
        assert make({"type": "cg", "precond": {"type": "ilu", "beta": 0.25}}).ilu_beta == 0.25


def test_descriptor_errors_name_the_problem():
    from ugcore_b200 import solver as S
    for bad, word in (({"type": "sor"}, "linear solver"), ({"type": "cg", "precond": "amg"}, "preconditioner"),
                      ({"type": "cg", "precond": {"type": "gmg", "topLevel": 2, "smoother": "vanka"}}, "smoother"),
                      ({"type": "cg", "precond": {"type": "ilu", "ordering": "nested-dissection"}}, "ILU ordering")):
        with pytest.raises(ValueError, match=word):
            S.make_desc(bad)
    assert S.make_desc({"type": "bicgstab"}).restart == 0 and S.make_desc({"type": "bicgstab", "restart": 6}).restart == 6


def test_profile_zones_carry_ugcore_names():
    """NVTX ranges around the launches use ugcore's profiler zone names (mg_solver_impl.hpp:1698 GMG_PreSmooth, :1800
    GMG_Restrict_Transfer, :1863 GMG_Prolongate_Transfer, :1913 GMG_PostSmooth, :1990 GMG_BaseSolver_Apply; cg.h:105
    CG_apply_return_defect; sparsematrix_impl.h:298 SparseMatrix_axpy): the names are in the host library."""
    blob = open(os.path.join(ROOT, "ugcore_b200", "lib", "libug4b200_host.so"), "rb").read()
    for name in (b"GMG_PreSmooth", b"GMG_Restrict_Transfer", b"GMG_Prolongate_Transfer", b"GMG_PostSmooth", b"GMG_BaseSolver_Apply",
                 b"GMG_Apply_lmgc", b"CG_apply_return_defect", b"LS_ApplyReturnDefect", b"SparseMatrix_axpy"):
        assert name in blob, name
