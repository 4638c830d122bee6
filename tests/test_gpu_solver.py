"""GPU parity, solver level: GMG-preconditioned Krylov solves through the descriptor C ABI
against the CPU oracle (compiled reference kernels when oracle/_ref is present).

Tolerances are north_star's: fp64 residual histories within 1e-10 relative per iteration,
iteration count +-1, final solution within 1e-9 relative L2.
"""
import json
import os

import numpy as np
import pytest

from helpers import gmg_desc, make_rhs, oracle_levels, rel_hist_err, sens_tol

pytestmark = pytest.mark.gpu

HIST_TOL = 1e-10
SOL_TOL = 1e-9
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "residual_histories.json")


def _best_oracle():
    import oracle
    return oracle.Oracle("ref" if oracle.have_ref() else "port")


def _compare(prob, desc, flags=0):
    import oracle
    import ugcore_b200 as ug
    orc = _best_oracle()
    pc = desc.get("precond")
    b = prob.rhs()
    if isinstance(pc, dict) and pc.get("type") == "gmg":
        lv = oracle_levels(orc, prob, pc["baseLevel"], pc["topLevel"])
        osol = oracle.OSolver(orc, desc, lv[pc["topLevel"]][0], lv)
    else:
        osol = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
    xo, oko, ho = osol.apply(b)
    s = ug.Solver.from_problem(desc, prob, flags=flags)
    xg, okg, hg = s.apply(b)
    assert okg == oko
    assert abs(len(hg) - len(ho)) <= 1, (len(hg), len(ho))
    # north_star's 1e-10 — unless the reference's own history moves by more than a tenth of that under a reordered sum
    # (measured on this very solve: helpers.sens_tol; 1e-10 for every GMG-CG case)
    tol = sens_tol(orc, osol, b, base=HIST_TOL)
    assert rel_hist_err(hg, ho) < tol, (rel_hist_err(hg, ho), tol, hg, ho)
    assert np.linalg.norm(xg - xo) <= max(SOL_TOL, tol) * np.linalg.norm(xo)
    return s, hg, ho


def test_poisson3d_gmg_cg_matches_oracle():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=4)
    s, hg, ho = _compare(prob, gmg_desc(4))
    assert len(hg) == len(ho)


@pytest.mark.parametrize("base", [(2, 2, 2), (3, 3, 3), (2, 1, 1)])
def test_poisson3d_several_base_cells_matches_oracle(base):
    """Base grids of more than one cell: the coarsest level then has interior DoFs, so the base solve (dense LU on the
    device) contributes to every cycle — on the one-cell unit cube it only ever sees Dirichlet rows.  2x2x2 cells with
    numRefs = 7 is the 257^3 grid of BASELINE configs[2] (bench.py --scaling strong), 3x3x3 the elasticity bench grid."""
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3, base=base)
    s, hg, ho = _compare(prob, gmg_desc(3))
    assert len(hg) == len(ho)
    prob = pr.Problem(dim=3, num_refs=2, base=base, problem=pr.ELASTICITY)
    _compare(prob, gmg_desc(2, reduction=1e-8, its=200))


def test_poisson2d_cfg1_standin_gmg_cg():
    """S1: 2-D unit square, quads, GMG(Jacobi V(2,2)) + CG (BASELINE.json configs[0] stand-in)."""
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=2, num_refs=6)
    _compare(prob, gmg_desc(6))


@pytest.mark.parametrize("flags", [1, 2, 4, 8, 1 | 4])
def test_execution_variants_agree(flags):
    """host scalars / no graph / unfused Jacobi / reference's extra top-level defect update:
    all are the same algorithm and must give the same history."""
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3)
    _compare(prob, gmg_desc(3), flags=flags)


@pytest.mark.parametrize("graph_flags", [0, 2])
def test_batched_small_operations_are_bit_identical(graph_flags):
    """The one-cluster batch kernel (csrc/batch.cu) executes the recorded coarse-level operations in
    call order with the stand-alone kernels' arithmetic: histories and solutions must be equal bit
    for bit with recording switched off, with and without CUDA-graph replay, and the batch kernel
    must actually have run."""
    import ctypes as C
    import ugcore_b200 as ug
    from ugcore_b200 import capi, problems as pr
    from ugcore_b200.solver import host_ctx
    prob = pr.Problem(dim=3, num_refs=4)
    ctx = host_ctx()
    out = {}
    for on in (1, 0):
        capi.check(capi.dev.ug4b200_batch_enable(ctx, on, -1), ctx)
        n0, cl = C.c_int64(), C.c_int()
        capi.dev.ug4b200_batch_stats(ctx, C.byref(n0), C.byref(cl))
        s = ug.Solver.from_problem(gmg_desc(4), prob, flags=graph_flags)
        x, ok, h = s.apply(prob.rhs())
        n1 = C.c_int64()
        capi.dev.ug4b200_batch_stats(ctx, C.byref(n1), C.byref(cl))
        assert ok
        out[on] = (np.array(x), np.array(h), n1.value - n0.value, cl.value)
    capi.check(capi.dev.ug4b200_batch_enable(ctx, 1, -1), ctx)
    assert out[1][3] >= 1, "no cluster launch available on this device"
    assert out[1][2] > 0 and out[0][2] == 0, (out[1][2], out[0][2])
    assert np.array_equal(out[1][1], out[0][1])
    assert np.array_equal(out[1][0], out[0][0])


@pytest.mark.parametrize("cycle,nu", [("V", (1, 1)), ("V", (3, 3)), ("W", (2, 2)), ("F", (2, 1)), ("V", (2, 0))])
def test_cycle_types_and_smoothing_counts(cycle, nu):
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3)
    desc = gmg_desc(3, cycle=cycle, nu=nu, solver="linear" if nu[1] == 0 else "cg", reduction=1e-8)
    _compare(prob, desc)


def test_hierarchical_dof_order():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=4, order=pr.ORDER_HIER)
    _compare(prob, gmg_desc(4))


def test_base_level_above_zero_and_coarse_cg():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=4)
    _compare(prob, gmg_desc(4, base=2))
    d = gmg_desc(4, base=2, base_solver={"type": "cg", "convCheck": {"iterations": 500, "absolute": 1e-30, "reduction": 1e-14}})
    _compare(prob, d)


def test_linear_solver_with_gmg():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3)
    _compare(prob, gmg_desc(3, solver="linear", reduction=1e-8))


def test_cg_jacobi_and_plain_cg():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=2, num_refs=4)
    cc = {"iterations": 400, "absolute": 1e-12, "reduction": 1e-8}
    _compare(prob, {"type": "cg", "precond": {"type": "jac", "damp": 0.66}, "convCheck": cc})
    _compare(prob, {"type": "cg", "precond": None, "convCheck": cc})


def test_max_steps_reached_reports_failure():
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3)
    s, hg, ho = _compare(prob, gmg_desc(3, its=3, reduction=1e-30))
    assert s.steps == 3 and len(hg) == 4


def test_zero_rhs_converges_immediately():
    import ugcore_b200 as ug
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=2)
    s = ug.Solver.from_problem(gmg_desc(2), prob)
    x, ok, h = s.apply(np.zeros(prob.num_dofs))
    assert ok and len(h) == 1 and h[0] == 0.0 and not x.any()


def test_convdiff_bicgstab_gmg_multicolor_gs():
    """S4: upwind convection-diffusion, BiCGStab + GMG with multicolour GS smoothing.  The
    oracle runs the reference's lexicographic gs_step_LL over the same colour-sorted matrices."""
    import ctypes as C
    import oracle
    import ugcore_b200 as ug
    from ugcore_b200 import problems as pr
    from test_gpu_kernels import _color_sorted
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-1)
    desc = gmg_desc(3, solver="bicgstab", smoother={"type": "gs", "relax": 1.0}, reduction=1e-8)
    s = ug.Solver.from_problem(desc, prob)
    xg, okg, hg = s.apply(prob.rhs())
    assert okg
    # oracle on colour-permuted levels (perm from the same greedy colouring the library uses)
    orc = _best_oracle()
    perms, lv = {}, {}
    for l in range(0, 4):
        PA, perm, _ = _color_sorted(prob.matrix(l))
        perms[l] = perm
        lv[l] = [orc.matrix(PA), None, None]
    from ugcore_b200.problems import Crs
    import scipy.sparse as sp
    for l in range(1, 4):
        for name, crs in (("P", prob.prolongation(l)), ("R", prob.restriction(l))):
            M = sp.csr_matrix((crs.vals + 0.0, crs.cols, crs.rowptr), shape=(crs.nrows, crs.ncols))
            # keep explicit zeros: add a marker, permute, remove it
            M.data += 10.0
            pr_, pc_ = (perms[l], perms[l - 1]) if name == "P" else (perms[l - 1], perms[l])
            coo = M.tocoo()
            Mp = sp.csr_matrix((coo.data, (pr_[coo.row], pc_[coo.col])), shape=M.shape)
            Mp.sort_indices()
            c = Crs(M.shape[0], M.shape[1], 1, Mp.indptr.astype(np.int64), Mp.indices.astype(np.int32), Mp.data - 10.0)
            lv[l][1 if name == "P" else 2] = orc.matrix(c)
    lv = {l: tuple(v) for l, v in lv.items()}
    osol = oracle.OSolver(orc, desc, lv[3][0], lv)
    bperm = np.empty_like(prob.rhs()); bperm[perms[3]] = prob.rhs()
    xo, oko, ho = osol.apply(bperm)
    assert oko
    assert abs(len(hg) - len(ho)) <= 1
    # north_star's 1e-10 unless the reference's own history moves more under a reordered sum (measured, helpers.sens_tol)
    assert rel_hist_err(hg, ho) < sens_tol(orc, osol, bperm)
    xo_orig = xo[perms[3]]
    assert np.linalg.norm(xg - xo_orig) <= 1e-7 * np.linalg.norm(xo_orig)


def test_elasticity_block3_gmg_cg():
    """S5: Q1 linear elasticity, 3x3 block-CRS, block-Jacobi GMG + CG."""
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.ELASTICITY)
    _compare(prob, gmg_desc(3, reduction=1e-8, its=200))


def test_golden_history_fixture():
    """GPU vs the committed golden histories (generated with the compiled reference kernels)."""
    import ugcore_b200 as ug
    from ugcore_b200 import problems as pr
    with open(GOLDEN) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        prob = pr.Problem(**case["problem"])
        s = ug.Solver.from_problem(case["desc"], prob)
        x, ok, h = s.apply(make_rhs(prob, case.get("rhs_seed")))
        ref = np.array(case["history"])
        assert ok == case["converged"], case["name"]
        assert abs(len(h) - len(ref)) <= 1, case["name"]
        assert rel_hist_err(h, ref) < HIST_TOL, case["name"]
        assert abs(np.linalg.norm(x) - case["solution_norm"]) <= 1e-9 * case["solution_norm"], case["name"]


def test_full_size_properties_129cubed():
    """BASELINE configs[1] at full size (2.1M DoF): properties that do not need the oracle —
    the returned defect history matches a freshly computed ||b - A x|| (the reference's own
    debug check, preconditioned_linear_operator_inverse.h:165-177), monotone energy-norm-like
    decrease of the defect, constant GMG iteration count vs the 65^3 grid."""
    import ugcore_b200 as ug
    from ugcore_b200 import problems as pr
    its = {}
    for refs in (6, 7):
        prob = pr.Problem(dim=3, num_refs=refs)
        s = ug.Solver.from_problem(gmg_desc(refs), prob)
        b = prob.rhs()
        x, ok, h = s.apply(b)
        assert ok
        its[refs] = len(h) - 1
        A = prob.matrix().to_scipy()
        fresh = np.linalg.norm(b - A @ x)
        assert abs(fresh - h[-1]) <= 1e-6 * h[0]
        assert h[-1] / h[0] < 1e-10
        assert np.all(np.diff(h) < 0)
        # discretisation error is O(h^2)
        assert np.abs(x - prob.exact()).max() < 0.6 * (0.5 ** refs) ** 2 * 10
    assert abs(its[6] - its[7]) <= 1


def test_full_size_129cubed_history_matches_the_reference():
    """BASELINE configs[1] at full size against the ORACLE (compiled ugcore kernels, one serial solve of 2.1 M DoF, ~10 s):
    the kernels that bench.py times — persistent bulk-copy SpMV on the value-indexed stream, fused smoothing, batched
    coarse levels, device-resident CG in a CUDA graph — are the ones compared here, at north_star's tolerances."""
    prob = __import__("ugcore_b200").problems.Problem(dim=3, num_refs=7)
    s, hg, ho = _compare(prob, gmg_desc(7))
    assert len(hg) == len(ho) == 9


@pytest.mark.parametrize("order", ["hier", "hier_cmk"])
def test_65cubed_hierarchical_order_matches_the_reference(order):
    """ugcore's DoF order after global refinement (coarse vertices first, then edge / face / volume midpoints): the
    16-bit column window of the value-indexed stream does not fit, x-gathers lose their locality, the plain 12 B stream
    runs — and with Cuthill-McKee at upload a banded numbering is restored.  Same bar as above."""
    import oracle
    import ugcore_b200 as ug
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=6, order=pr.ORDER_HIER)
    desc = gmg_desc(6)
    orc = _best_oracle()
    lv = oracle_levels(orc, prob, 0, 6)
    osol = oracle.OSolver(orc, desc, lv[6][0], lv)
    b = np.array(prob.rhs())
    xo, oko, ho = osol.apply(b)
    s = ug.Solver.from_problem(desc, prob, order="cmk" if order == "hier_cmk" else None)
    xg, okg, hg = s.apply(b)
    assert okg and oko and len(hg) == len(ho)
    assert rel_hist_err(hg, ho) < HIST_TOL
    assert np.linalg.norm(xg - xo) <= SOL_TOL * np.linalg.norm(xo)


def test_value_indexed_and_plain_streams_give_identical_histories():
    """The value-indexed entry stream is lossless: the whole solve (history, iterates) must be
    bit-identical to a run with the plain stream (UG4B200_NO_COMPRESS=1, separate process because
    the host layer's context reads the switch once)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests');"
            "import numpy as np; import ugcore_b200 as ug; from ugcore_b200 import problems as pr; from helpers import gmg_desc;"
            "p = pr.Problem(dim=3, num_refs=5); s = ug.Solver.from_problem(gmg_desc(5), p); x, ok, h = s.apply(p.rhs());"
            "print('RES ' + json.dumps({'ok': bool(ok), 'h': [float(v).hex() for v in h], 'x': float(np.sum(x)).hex()}))") % (root, root)
    out = []
    for nc in ("0", "1"):
        env = dict(os.environ, UG4B200_NO_COMPRESS=nc)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES ")]
        assert line, r.stdout[-2000:] + r.stderr[-2000:]
        out.append(json.loads(line[0][4:]))
    assert out[0]["ok"] and out[1]["ok"]
    assert out[0]["h"] == out[1]["h"] and out[0]["x"] == out[1]["x"]


@pytest.mark.parametrize("precond", ["jac", "gs", "ilu", "gmg"])
@pytest.mark.parametrize("solver", ["cg", "bicgstab", "linear"])
def test_reinit_with_a_changed_matrix_recaptures_the_graph(solver, precond):
    """solver:init(J, u) with a re-assembled J on the SAME solver object (every Newton / time step in ugcore): first
    with other values, then with another pattern at the same n (an explicit-zero connection removed from some rows).
    The device-resident loops replay a CUDA graph that bakes in the matrix mirror and the preconditioner's buffers;
    the pool hands the re-created work vectors the same addresses, so only the generation of the device data
    tells a stale graph from a valid one.  Every solve must equal the oracle's solve of the matrix it was given."""
    import oracle
    import ugcore_b200 as ug
    from ugcore_b200 import capi, problems as pr
    orc = _best_oracle()
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF if solver != "cg" else pr.POISSON, eps=0.5)
    A0 = prob.matrix()
    cc = {"iterations": 200, "absolute": 1e-14, "reduction": 1e-8}
    if precond == "gmg":
        desc = gmg_desc(3, solver=solver, reduction=1e-8)
    else:
        pc = {"jac": {"type": "jac", "damp": 0.66}, "gs": {"type": "sgs" if solver == "cg" else "gs"},
              "ilu": {"type": "ilu", "ordering": "multicolor"}}[precond]
        desc = {"type": solver, "precond": pc, "convCheck": cc}
    flags = 0      # device-resident loops + CUDA graphs are the default for CG, BiCGStab and LinearSolver
    # second matrix: same pattern, other values (2 A + 25 % more diagonal); third: another pattern at the same n (the
    # explicit zeros that the Dirichlet rows keep, sparsematrix_util.h:850-861, removed: nnz and the slices change)
    rows = np.repeat(np.arange(A0.nrows), np.diff(A0.rowptr))
    diag = rows == A0.cols
    v1 = 2.0 * np.asarray(A0.vals) + (0.0 if precond == "gmg" else np.where(diag, 0.25 * np.asarray(A0.vals), 0.0))
    A1 = pr.Crs(A0.nrows, A0.ncols, 1, A0.rowptr.copy(), A0.cols.copy(), v1)
    dirich = np.asarray(prob.dirichlet(), bool)
    keep = ~(dirich[rows] & ~diag)
    assert keep.sum() < A0.nnz
    rp2 = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=A0.nrows))]).astype(np.int64)
    A2 = pr.Crs(A0.nrows, A0.ncols, 1, rp2, A0.cols[keep].copy(), np.asarray(A0.vals)[keep].copy())
    b = make_rhs(prob, 7)

    def scaled(M, f):
        return pr.Crs(M.nrows, M.ncols, M.block, M.rowptr, M.cols, f * np.asarray(M.vals))

    def levels_of(A):
        """GMG hierarchy that goes with the top-level matrix A: the coarse operators are re-assembled together with it
        (A1 = 2 A0 + ... on the top level comes with 2 x the coarse level matrices, so the cycle stays a preconditioner)"""
        if precond != "gmg":
            return None
        f = 2.0 if A is A1 else 1.0
        return {l: (A if l == 3 else scaled(prob.matrix(l), f), prob.prolongation(l) if l else None, prob.restriction(l) if l else None)
                for l in range(0, 4)}

    def oracle_solve(A):
        if precond == "gmg":
            lvA = levels_of(A)
            lv = {l: (orc.matrix(t[0]), orc.matrix(t[1]) if t[1] is not None else None, orc.matrix(t[2]) if t[2] is not None else None)
                  for l, t in lvA.items()}
            return oracle.OSolver(orc, desc, lv[3][0], lv).apply(b)
        if precond in ("gs", "ilu"):   # the device sweeps in the greedy multicolour order of the given pattern
            from helpers import greedy_color_perm, permute_crs
            perm, _ = greedy_color_perm(A)
            PA = permute_crs(A, perm, perm)
            pb = np.empty_like(b); pb[perm] = b
            odesc = dict(desc)
            if precond == "ilu":
                odesc["precond"] = {"type": "ilu"}
            xo, ok, h = oracle.OSolver(orc, odesc, orc.matrix(PA)).apply(pb)
            return xo[perm], ok, h
        return oracle.OSolver(orc, desc, orc.matrix(A)).apply(b)

    s = ug.Solver(desc, A0, levels_of(A0), flags=flags)
    for A in (A0, A1, A2, A0):
        if A is not A0 or s._inited:
            s.set_matrix(A, levels_of(A))
        xg, okg, hg = s.apply(b)
        xo, oko, ho = oracle_solve(A)
        assert okg == oko
        assert abs(len(hg) - len(ho)) <= 1, (len(hg), len(ho))
        assert rel_hist_err(hg, ho) < 1e-8, (hg, ho)    # a stale graph gives O(1) errors; exact parity is other tests' job
        assert np.linalg.norm(xg - xo) <= 1e-7 * np.linalg.norm(xo)


def test_bicgstab_breakdown_ends_the_solve_like_the_reference():
    """(v, r0) == 0 (bicgstab.h:276-281: "return false"): a matrix that maps the first search direction to 0.  Host loop,
    device-resident loop and oracle all report failure, none of them leaves inf / NaN in x (the device loop flags the
    breakdown in the finaliser that would divide, before x is updated; the guard turns the rest of the graph into no-ops)."""
    import oracle
    import ugcore_b200 as ug
    from ugcore_b200 import capi
    from ugcore_b200.problems import Crs
    n = 96
    A = Crs(n, n, 1, np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), np.zeros(n))
    b = np.linspace(1.0, 2.0, n)
    desc = {"type": "bicgstab", "precond": None, "convCheck": {"iterations": 20, "absolute": 1e-12, "reduction": 1e-8}}
    orc = _best_oracle()
    xo, oko, ho = oracle.OSolver(orc, desc, orc.matrix(A)).apply(b)
    assert not oko
    for flags in (capi.FLAG_HOST_SCALARS, 0, capi.FLAG_NO_GRAPH):
        x, ok, h = ug.Solver(desc, A, flags=flags).apply(b)
        assert not ok, flags
        assert np.isfinite(x).all() and np.array_equal(x, xo), flags
        assert len(h) == len(ho) and h[0] == pytest.approx(ho[0], rel=1e-14), flags
