"""The drop-in boundary compiled against ugcore's REAL headers.

tests/boundary/boundary.cpp is built with -DUG4B200_WITH_UGCORE -I/root/reference/ugbase: csrc/host/*.h then sit on
ugcore's own common.h / smart_pointer.h / small_algebra.h / algebra_type.h and on the real ILinearOperator,
MatrixOperator, ILinearIterator, IPreconditioner, IVectorDebugWriter (the interface headers that need no boost);
static_asserts in that file pin the class relations, this module builds and runs it.

  CPU (here):  the translation unit compiles and links; `boundary_test types` — host only — casts an
               ug::ILinearOperator SmartPtr back to ug::MatrixOperator<GPUSparseMatrix<double>, GPUVector<double>>
               with ugcore's cast_dynamic, clones a Jacobi<GPUAlgebra> through ug::ILinearIterator, raises ugcore's UGError.
  GPU:         `boundary_test solve` — the same assembled matrix in CPUAlgebra (ugcore's SparseMatrix / Vector) and
               GPUAlgebra, y = A x and y -= A x through ug::ILinearOperator* bit-identical; GMG-CG through
               ug::ILinearIterator / ug::ILinearOperator pointers; its history against the oracle's.
/root/reference does not exist on the GPU box: the binary is built here (also by __graft_entry__.build()) and travels.
"""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BDIR = os.path.join(ROOT, "tests", "boundary")
EXE = os.path.join(BDIR, "_build", "boundary_test")
HAVE_REF = os.path.isdir("/root/reference/ugbase")


def _run(*args):
    r = subprocess.run([EXE, *args], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    return r.returncode, json.loads(lines[-1])


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference absent (GPU box): the prebuilt binary is used")
def test_boundary_translation_unit_compiles_against_ugcore_headers():
    r = subprocess.run(["make", "-C", BDIR, "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(EXE)
    # the stand-ins really are out of the picture in that configuration
    src = open(os.path.join(ROOT, "ugcore_b200", "csrc", "host", "operators.h")).read()
    assert '#include "lib_algebra/operator/interface/preconditioner.h"' in src and "#ifndef UG4B200_WITH_UGCORE" in src


@pytest.mark.skipif(not os.path.exists(EXE) and not HAVE_REF, reason="boundary_test not built")
def test_boundary_types_through_ugcore_base_classes():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", BDIR, "all"], check=True, capture_output=True)
    rc, res = _run("types")
    assert rc == 0 and res["ok"] and res["with_ugcore"], res
    assert res["algebra_type"] == [1, 1] and res["block_algebra_type"] == [1, 3]     # AlgebraType::GPU == 1 (algebra_type.h:55)


def test_integration_patches_apply_to_the_reference_tree(tmp_path):
    """integration/*.patch are real unified diffs against ugcore: each one applies cleanly (dry run) to the file it names."""
    if not HAVE_REF:
        pytest.skip("/root/reference absent")
    pdir = os.path.join(ROOT, "integration")
    patches = sorted(f for f in os.listdir(pdir) if f.endswith(".patch"))
    assert len(patches) >= 4
    for f in patches:
        r = subprocess.run(["patch", "--dry-run", "-p1", "-d", "/root/reference", "-o", os.devnull, "-i", os.path.join(pdir, f)],
                           capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stdout, r.stderr)


@pytest.mark.gpu
def test_gpu_solve_through_ugcore_interfaces_matches_oracle():
    """GMG V(2,2) Jacobi + CG wired by hand from the GPU classes, called only through ug::ILinearOperator /
    ug::ILinearIterator pointers of the REAL headers; SpMV vs ugcore's CPUAlgebra bit-exact; history vs the oracle."""
    import oracle
    from helpers import gmg_desc, oracle_levels, rel_hist_err
    from ugcore_b200 import problems as pr
    assert os.path.exists(EXE), "tests/boundary/_build/boundary_test must be built where /root/reference exists"
    refs = 4
    rc, res = _run("solve", str(refs))
    assert rc == 0 and res["ok"], res
    assert res["spmv_bit_exact"] and res["matmul_minus_bit_exact"] and res["converged"]
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    prob = pr.Problem(dim=3, num_refs=refs)
    lv = oracle_levels(orc, prob, 0, refs)
    xo, oko, ho = oracle.OSolver(orc, gmg_desc(refs), lv[refs][0], lv).apply(np.array(prob.rhs()))
    h = np.array(res["history"])
    assert oko and len(h) == len(ho) and res["steps"] == len(ho) - 1
    assert rel_hist_err(h, ho) < 1e-10
    assert abs(res["solution_norm"] - np.linalg.norm(xo)) <= 1e-9 * np.linalg.norm(xo)
    assert 0.0 < res["cycle_reduction"] < 0.2        # one V(2,2) cycle through ILinearIterator::apply_update_defect


def test_patched_cpu_algebra_types_header_compiles_with_the_gpu_algebra(tmp_path):
    """integration/0001 applied to a scratch copy of cpu_algebra_types.h inside a symlink overlay of ugbase, with
    csrc/host installed as lib_algebra/gpu_algebra/ug4b200: `#define UG_GPU` + the patched header
    give ug::GPUAlgebra next to ug::CPUAlgebra, the way bridge/util_algebra_dependent.h:63-67 expects it."""
    if not HAVE_REF:
        pytest.skip("/root/reference absent")
    ref = "/root/reference/ugbase"
    ov = tmp_path / "ugbase"
    ov.mkdir()
    for e in os.listdir(ref):
        if e != "lib_algebra":
            os.symlink(os.path.join(ref, e), ov / e)
    la = ov / "lib_algebra"
    la.mkdir()
    for e in os.listdir(os.path.join(ref, "lib_algebra")):
        if e not in ("cpu_algebra_types.h", "gpu_algebra"):
            os.symlink(os.path.join(ref, "lib_algebra", e), la / e)
    (la / "gpu_algebra").mkdir()
    os.symlink(os.path.join(ROOT, "ugcore_b200", "csrc", "host"), la / "gpu_algebra" / "ug4b200")
    r = subprocess.run(["patch", "-p1", "-d", "/root/reference", "-o", str(la / "cpu_algebra_types.h"), "-i",
                        os.path.join(ROOT, "integration", "0001-cpu_algebra_types_h.patch")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "lib_algebra/cpu_algebra_types.h"\n'
                  "static_assert(ug::GPUAlgebra::blockSize == 1 && ug::CPUAlgebra::blockSize == 1, \"\");\n"
                  "static_assert(ug::GPUBlockAlgebra<3>::blockSize == ug::CPUBlockAlgebra<3>::blockSize, \"\");\n"
                  "ug::GPUAlgebra::matrix_type* pm; ug::GPUAlgebra::vector_type* pv; ug::CPUAlgebra::matrix_type* pc;\n"
                  "int main() { return ug::GPUAlgebra::get_type().type() == ug::AlgebraType::GPU ? 0 : 1; }\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-DUG_GPU", "-DUG4B200_WITH_UGCORE", f"-I{ov}", f"-I{ROOT}/include",
                        str(tu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
