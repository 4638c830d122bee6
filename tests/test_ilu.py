"""ILU(0) / ILU(beta), GMRES and Cuthill-McKee ordering.

CPU (-m "not gpu"):
  * the oracle's port of FactorizeILUSorted / FactorizeILUBeta / invert_L / invert_U / ComputeCuthillMcKeeOrder
    against the REAL reference functions compiled into oracle/_ref (operator/preconditioner/ilu.h:110-322,
    ordering_strategies/algorithms/native_cuthill_mckee.cpp) — bit for bit;
  * the product's host-side factorisation and ordering (csrc/host/ilu_factor.h, through the C ABI) against
    the same compiled reference — bit for bit;
  * the level sets the device sweeps are scheduled by: valid (no dependency inside a level) and minimal;
  * the oracle's restated GMRES and ILU-preconditioned solvers converge and agree with scipy's view of the
    solution.
GPU: ILU / GMRES solves against the oracle (first GPU run pending).
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, greedy_color_perm, permute_crs, rel_hist_err, sens_tol
from ugcore_b200 import problems as pr


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


PROBLEMS = {
    "poisson3d": lambda: pr.Problem(dim=3, num_refs=3),
    "convdiff3d": lambda: pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-1),
    "poisson2d": lambda: pr.Problem(dim=2, num_refs=4),
    "hier3d": lambda: pr.Problem(dim=3, num_refs=3, order=pr.ORDER_HIER),
}


def _host_ilu(A, beta=0.0, sort_eps=1e-50):
    from ugcore_b200.capi import check_host, host
    va = np.array(A.vals, dtype=np.float64)
    check_host(host.ug4b200_host_ilu_factorize(A.nrows, _p(np.ascontiguousarray(A.rowptr)), _p(np.ascontiguousarray(A.cols)),
                                               _p(va), beta, sort_eps))
    return va


def _host_cmk(A, reverse, consec):
    from ugcore_b200.capi import check_host, host
    ni = np.zeros(A.nrows, np.int64)
    check_host(host.ug4b200_host_cuthill_mckee(A.nrows, _p(np.ascontiguousarray(A.rowptr)), _p(np.ascontiguousarray(A.cols)),
                                               int(reverse), int(consec), _p(ni)))
    return ni


def _host_levels(A, lower):
    from ugcore_b200.capi import check_host, host
    lev = np.zeros(A.nrows, np.int32)
    nl = C.c_int()
    check_host(host.ug4b200_host_level_sets(A.nrows, _p(np.ascontiguousarray(A.rowptr)), _p(np.ascontiguousarray(A.cols)),
                                            int(lower), _p(lev), C.byref(nl)))
    return lev, nl.value


@pytest.mark.parametrize("kind", sorted(PROBLEMS))
@pytest.mark.parametrize("beta", [0.0, 0.5])
def test_ilu_factorisation_is_bit_identical_to_the_reference(kind, beta, orc, orc_ref):
    prob = PROBLEMS[kind]()
    A = prob.matrix()
    rp, ci, va_ref = orc_ref.matrix(A).ilu(beta).export()
    assert np.array_equal(rp, A.rowptr) and np.array_equal(ci, A.cols)          # ILU(0): the pattern is A's
    assert np.array_equal(orc.matrix(A).ilu(beta).export()[2], va_ref)          # port of the oracle
    assert np.array_equal(_host_ilu(A, beta), va_ref)                           # the product's host factorisation
    # and one application of (LU)^-1: port == reference
    d = np.random.default_rng(3).standard_normal(A.nrows)
    assert np.array_equal(orc.matrix(A).ilu(beta).ilu_apply(d), orc_ref.matrix(A).ilu(beta).ilu_apply(d))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("beta", [0.0, 0.4])
def test_block_ilu_factorisation_is_bit_identical_to_the_reference(dim, beta, orc_ref):
    """2x2 / 3x3 blocks (elasticity): DenseMatrix /=, *, -= and the Cramer inverses restated in
    csrc/host/ilu_factor.h against the real small_algebra templates under the reference's FactorizeILU*."""
    from ugcore_b200.capi import check_host, host
    prob = pr.Problem(dim=dim, num_refs=2 if dim == 3 else 3, problem=pr.ELASTICITY)
    A = prob.matrix()
    assert A.block == dim
    va = np.array(A.vals, dtype=np.float64)
    check_host(host.ug4b200_host_ilu_factorize_block(A.block, A.nrows, _p(np.ascontiguousarray(A.rowptr)),
                                                     _p(np.ascontiguousarray(A.cols)), _p(va), beta, 1e-50))
    assert np.array_equal(va, orc_ref.matrix(A).ilu(beta).export()[2])


def test_ilu_is_exact_on_its_pattern(orc_ref):
    """ILU(0): (L U)_ij = A_ij wherever A stores an entry (Saad, Iterative Methods, Prop. 10.2)."""
    import scipy.sparse as sp
    A = PROBLEMS["convdiff3d"]().matrix()
    rp, ci, va = orc_ref.matrix(A).ilu(0.0).export()
    F = sp.csr_matrix((va, ci, rp), shape=(A.nrows, A.nrows))
    L = sp.tril(F, -1) + sp.identity(A.nrows)
    U = sp.triu(F, 0)
    LU = (L @ U).tocsr()
    S = A.to_scipy()
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
    assert np.allclose(np.asarray(LU[rows, A.cols]).ravel(), np.asarray(S[rows, A.cols]).ravel(), rtol=1e-12, atol=1e-14)


def test_ilu_near_zero_pivot_is_reported():
    from ugcore_b200.capi import host
    rp = np.array([0, 2, 4], np.int64); ci = np.array([0, 1, 0, 1], np.int32); va = np.array([0.0, 1.0, 1.0, 1.0])
    assert host.ug4b200_host_ilu_factorize(2, _p(rp), _p(ci), _p(va), 0.0, 1e-50) != 0
    assert b"near-zero" in host.ug4b200_host_last_error()


@pytest.mark.parametrize("kind", sorted(PROBLEMS))
def test_cuthill_mckee_equals_the_reference(kind, orc, orc_ref):
    A = PROBLEMS[kind]().matrix()
    for reverse in (True, False):
        for consec in (True, False):
            ref = orc_ref.matrix(A).cuthill_mckee(reverse, consec)
            assert np.array_equal(np.sort(ref), np.arange(A.nrows))
            assert np.array_equal(orc.matrix(A).cuthill_mckee(reverse, consec), ref)
            assert np.array_equal(_host_cmk(A, reverse, consec), ref)
    # it does what it is for: the bandwidth of the hierarchically numbered matrix shrinks
    if kind == "hier3d":
        ni = _host_cmk(A, True, False)
        rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
        assert np.abs(ni[rows] - ni[A.cols]).max() < np.abs(rows - A.cols).max()


def test_cuthill_mckee_with_unconnected_rows(orc, orc_ref):
    """Rows without connections go to the end (bPreserveConsec = false) or keep their place (true)."""
    from ugcore_b200.problems import Crs
    A0 = pr.Problem(dim=2, num_refs=2).matrix()
    keep = np.ones(A0.nrows, bool); keep[[3, 4, 11]] = False
    lens = np.where(keep, np.diff(A0.rowptr), 0)
    rows = np.repeat(np.arange(A0.nrows), np.diff(A0.rowptr))
    sel = keep[rows] & keep[A0.cols]
    lens = np.bincount(rows[sel], minlength=A0.nrows)
    A = Crs(A0.nrows, A0.ncols, 1, np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), A0.cols[sel].copy(), A0.vals[sel].copy())
    for reverse in (True, False):
        for consec in (True, False):
            ref = orc_ref.matrix(A).cuthill_mckee(reverse, consec)
            assert np.array_equal(orc.matrix(A).cuthill_mckee(reverse, consec), ref)
            assert np.array_equal(_host_cmk(A, reverse, consec), ref)


@pytest.mark.parametrize("kind", ["poisson3d", "hier3d"])
def test_level_sets_schedule_the_triangular_solves(kind):
    A = PROBLEMS[kind]().matrix()
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
    for lower in (True, False):
        lev, nl = _host_levels(A, lower)
        dep = (A.cols < rows) if lower else (A.cols > rows)
        assert np.all(lev[A.cols[dep]] < lev[rows[dep]])              # a row only depends on earlier levels
        need = np.zeros(A.nrows, np.int64)
        np.maximum.at(need, rows[dep], lev[A.cols[dep]] + 1)
        assert np.array_equal(need, lev) and nl == lev.max() + 1      # and sits on the earliest possible one
    # a colour-sorted matrix has at most as many levels as colours
    perm, cptr = greedy_color_perm(A)
    lev, nl = _host_levels(permute_crs(A, perm, perm), True)
    assert nl <= cptr.size - 1


DESCS = {
    "gmres_ilu": {"type": "gmres", "restart": 10, "precond": {"type": "ilu"}, "convCheck": {"iterations": 50, "absolute": 1e-12, "reduction": 1e-8}},
    "gmres": {"type": "gmres", "restart": 20, "precond": None, "convCheck": {"iterations": 200, "absolute": 1e-12, "reduction": 1e-6}},
    "bicgstab_ilub": {"type": "bicgstab", "precond": {"type": "ilu", "beta": 0.3}, "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}},
    "cg_ilu": {"type": "cg", "precond": {"type": "ilu"}, "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}},
    "linear_ilu": {"type": "linear", "precond": {"type": "ilu"}, "convCheck": {"iterations": 300, "absolute": 1e-12, "reduction": 1e-6}},
}


@pytest.mark.parametrize("name", sorted(DESCS))
def test_oracle_solvers_with_ilu_and_gmres(name, orc, orc_ref):
    """The restated GMRES (gmres.h:104-278) and ILU preconditioner on both backends: identical histories, and the
    returned solution really has the defect the history claims."""
    prob = PROBLEMS["poisson3d" if name == "cg_ilu" else "convdiff3d"]()
    A = prob.matrix()
    b = np.array(prob.rhs())
    xr, okr, hr = oracle.OSolver(orc_ref, DESCS[name], orc_ref.matrix(A)).apply(b)
    xp, okp, hp = oracle.OSolver(orc, DESCS[name], orc.matrix(A)).apply(b)
    assert okr and okp and np.array_equal(hr, hp) and np.array_equal(xr, xp)
    res = np.linalg.norm(b - A.to_scipy() @ xr)
    assert abs(res - hr[-1]) <= 1e-6 * hr[0]
    if name == "gmres":          # unpreconditioned: one convergence-check update per inner step, the last one of a cycle
        assert (len(hr) - 1) % 20 == 0


def _tiny():
    from ugcore_b200.problems import Crs
    return Crs(1, 1, 1, np.array([0, 1], np.int64), np.array([0], np.int32), np.array([2.0]))


GMRES_TINY = {"type": "gmres", "restart": 3, "precond": {"type": "ilu"}, "convCheck": {"iterations": 50, "absolute": 1e-12, "reduction": 1e-10}}


def test_oracle_gmres_without_lucky_breakdown_handling(orc, orc_ref):
    """gmres.h always runs all `restart` inner steps; once the Krylov space is exhausted h[j+1][j] = 0 and the
    normalisation divides by it (gmres.h:255).  The restated loop keeps that: NaN defect, not converged."""
    for o in (orc, orc_ref):
        x, ok, h = oracle.OSolver(o, GMRES_TINY, o.matrix(_tiny())).apply(np.array([4.0]))
        assert not ok and np.isnan(x).all() and h[0] == 4.0 and np.isnan(h[1])


GOLDEN_ILU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ilu_gmres_histories.json")


def _golden_solver(orc, case):
    from helpers import oracle_levels
    prob = pr.Problem(**case["problem"])
    desc = case["desc"]
    pc = desc.get("precond")
    if isinstance(pc, dict) and pc.get("type") == "gmg":
        lv = oracle_levels(orc, prob, pc["baseLevel"], pc["topLevel"])
        return prob, oracle.OSolver(orc, desc, lv[pc["topLevel"]][0], lv)
    return prob, oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_golden_ilu_gmres_histories(kind, request):
    """tests/golden/ilu_gmres_histories.json was generated with the reference's own ILU kernels (oracle/_ref);
    both backends reproduce it bit for bit (the port has no block ILU: that case is the ref backend's alone)."""
    import json
    from helpers import make_rhs
    orc = request.getfixturevalue("orc" if kind == "port" else "orc_ref")
    with open(GOLDEN_ILU) as f:
        gold = json.load(f)
    assert gold["oracle_backend"] == "ref" and len(gold["cases"]) == 7
    for case in gold["cases"]:
        if kind == "port" and case["problem"].get("problem") == 2:
            continue
        prob, s = _golden_solver(orc, case)
        x, ok, h = s.apply(make_rhs(prob, case.get("rhs_seed")))
        assert ok == case["converged"], case["name"]
        assert np.array_equal(h, np.array(case["history"])), case["name"]
        assert np.linalg.norm(x) == case["solution_norm"], case["name"]


# ---- GPU ----------------------------------------------------------------------------------------------------


def _best():
    return oracle.Oracle("ref" if oracle.have_ref() else "port")


# (the two apply-level tests only compose GPU-verified kernels — the Gauss-Seidel sweeps, gather / scatter — with the
#  host factorisation checked above: not gated)
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["poisson3d", "convdiff3d"])
@pytest.mark.parametrize("beta", [0.0, 0.4])
def test_gpu_ilu_multicolor_apply_is_bit_identical(kind, beta):
    """One application of the ILU preconditioner in the multicolour ordering == the reference's
    FactorizeILU* + invert_L + invert_U on the colour-permuted matrix, bit for bit."""
    import ugcore_b200 as ug
    prob = PROBLEMS[kind]()
    A = prob.matrix()
    desc = {"type": "linear", "precond": {"type": "ilu", "beta": beta, "ordering": "multicolor"},
            "convCheck": {"iterations": 1, "absolute": 1e-30, "reduction": 1e-30}}
    s = ug.Solver(desc, A)
    d = np.random.default_rng(5).standard_normal(A.nrows)
    c = s.precond_apply(d)
    perm, _ = greedy_color_perm(A)
    F = _best().matrix(permute_crs(A, perm, perm)).ilu(beta)
    dp = np.empty_like(d); dp[perm] = d
    assert np.array_equal(c, F.ilu_apply(dp)[perm])


@pytest.mark.gpu
def test_gpu_block_ilu_multicolor_apply_and_solve():
    """3x3-block ILU(0) on elasticity: one application bit-identical to the reference in the multicolour ordering,
    and ILU-preconditioned CG (natural ordering, level-scheduled) vs the oracle."""
    import ugcore_b200 as ug
    if not oracle.have_ref():
        pytest.skip("block ILU oracle needs oracle/_ref (the port restates the scalar case only)")
    orc = oracle.Oracle("ref")
    prob = pr.Problem(dim=3, num_refs=2, problem=pr.ELASTICITY)
    A, b = prob.matrix(), prob.block
    desc = {"type": "cg", "precond": {"type": "ilu", "ordering": "multicolor"},
            "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}}
    d = np.random.default_rng(8).standard_normal(A.nrows * b)
    c = ug.Solver(desc, A).precond_apply(d)
    perm, _ = greedy_color_perm(A)
    pp = np.repeat(perm * b, b) + np.tile(np.arange(b), perm.size)
    dp = np.empty_like(d); dp[pp] = d
    assert np.array_equal(c, orc.matrix(permute_crs(A, perm, perm)).ilu(0.0).ilu_apply(dp)[pp])
    desc["precond"] = {"type": "ilu"}
    x, ok, h = ug.Solver(desc, A).apply(prob.rhs())
    osol = oracle.OSolver(orc, desc, orc.matrix(A))
    xo, oko, ho = osol.apply(np.array(prob.rhs()))
    assert ok and oko and abs(len(h) - len(ho)) <= 1
    assert rel_hist_err(h, ho) < sens_tol(orc, osol, np.array(prob.rhs()))
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)


@pytest.mark.gpu
@pytest.mark.parametrize("ordering", ["natural", "cmk"])
def test_gpu_ilu_level_scheduled_apply(ordering):
    """Natural / Cuthill-McKee ordering: level-scheduled sweeps; same numbers as the reference up to the
    order in which a row sums its terms."""
    import ugcore_b200 as ug
    prob = PROBLEMS["convdiff3d"]()
    A = prob.matrix()
    desc = {"type": "linear", "precond": {"type": "ilu", "ordering": ordering},
            "convCheck": {"iterations": 1, "absolute": 1e-30, "reduction": 1e-30}}
    c = ug.Solver(desc, A).precond_apply(np.array(prob.rhs()))
    orc = _best()
    if ordering == "cmk":
        ni = orc.matrix(A).cuthill_mckee(False, True)        # NativeCuthillMcKeeOrdering: reverse = false, consecutive
        F = orc.matrix(permute_crs(A, ni, ni)).ilu(0.0)
        dp = np.empty(A.nrows); dp[ni] = prob.rhs()
        ref = F.ilu_apply(dp)[ni]
    else:
        ref = orc.matrix(A).ilu(0.0).ilu_apply(np.array(prob.rhs()))
    assert np.linalg.norm(c - ref) <= 1e-13 * np.linalg.norm(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DESCS))
def test_gpu_solvers_with_ilu_and_gmres_match_oracle(name):
    """CG / BiCGStab / LinearSolver / GMRES with ILU (natural ordering, level-scheduled) vs the oracle."""
    import ugcore_b200 as ug
    prob = PROBLEMS["poisson3d" if name == "cg_ilu" else "convdiff3d"]()
    A = prob.matrix()
    b = np.array(prob.rhs())
    x, ok, h = ug.Solver(DESCS[name], A).apply(b)
    orc = _best()
    osol = oracle.OSolver(orc, DESCS[name], orc.matrix(A))
    xo, oko, ho = osol.apply(b)
    assert ok and oko and abs(len(h) - len(ho)) <= 1
    tol = sens_tol(orc, osol, b)      # 1e-10, or 10 x the reference's own movement under a reordered sum (BiCGStab, GMRES)
    # natural-ordering ILU runs level-scheduled on the device: a row of the triangular solves accumulates its terms in
    # another order than the reference does, which the reduction-order yardstick does not see; GMRES + ILU
    # ends its single cycle at 1.05e-12 x the start defect, a hair above the round-off floor -> floor 1e-11 here
    assert rel_hist_err(h, ho, floor=1e-11) < tol, (rel_hist_err(h, ho, floor=1e-11), tol)
    assert np.linalg.norm(x - xo) <= 1e-7 * np.linalg.norm(xo)


@pytest.mark.gpu
def test_gpu_gmg_with_ilu_smoother_matches_oracle():
    """GMG V(2,2) with ILU(0) smoothing in the multicolour ordering + CG: the oracle runs on the colour-permuted
    hierarchy (like the Gauss-Seidel test)."""
    import ugcore_b200 as ug
    prob = pr.Problem(dim=3, num_refs=3)
    desc = gmg_desc(3, smoother={"type": "ilu", "ordering": "multicolor"})
    x, ok, h = ug.Solver.from_problem(desc, prob).apply(prob.rhs())
    orc = _best()
    perms = {l: (greedy_color_perm(prob.matrix(l))[0] if l else np.arange(prob.matrix(0).nrows)) for l in range(0, 4)}
    lv = {}
    for l in range(0, 4):
        lv[l] = (orc.matrix(permute_crs(prob.matrix(l), perms[l], perms[l])),
                 orc.matrix(permute_crs(prob.prolongation(l), perms[l], perms[l - 1])) if l else None,
                 orc.matrix(permute_crs(prob.restriction(l), perms[l - 1], perms[l])) if l else None)
    bp = np.empty(prob.num_dofs); bp[perms[3]] = prob.rhs()
    osol = oracle.OSolver(orc, desc, lv[3][0], lv)
    xo, oko, ho = osol.apply(bp)
    assert ok and oko and abs(len(h) - len(ho)) <= 1
    # the cycle reduces the defect by 1e-5 per step: the last step sits at 3e-12 x the start defect, where the reference's
    # own history moves by ~1e-9 under a reordered sum (sens_tol measures it)
    assert rel_hist_err(h, ho) < sens_tol(orc, osol, bp)
    assert np.linalg.norm(x - xo[perms[3]]) <= 1e-9 * np.linalg.norm(xo)


@pytest.mark.gpu
def test_gpu_edge_cases_ilu_gmres():
    """1x1 and diagonal systems, a matrix without diagonal (error like ugcore's), zero right-hand side, and the
    reference's GMRES behaviour once the Krylov space is exhausted (see the oracle test above)."""
    import scipy.sparse as sp
    import ugcore_b200 as ug
    from ugcore_b200.problems import Crs
    cc = {"iterations": 50, "absolute": 1e-12, "reduction": 1e-10}

    def crs(M):
        M = sp.csr_matrix(M); M.sort_indices()
        return Crs(M.shape[0], M.shape[1], 1, M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float64))

    for t in ("cg", "bicgstab", "linear"):
        x, ok, h = ug.Solver({"type": t, "precond": {"type": "ilu"}, "convCheck": cc}, _tiny()).apply(np.array([4.0]))
        assert ok and x[0] == 2.0
    x, ok, h = ug.Solver(GMRES_TINY, _tiny()).apply(np.array([4.0]))
    assert not ok and np.isnan(x).all() and h[0] == 4.0 and np.isnan(h[1])
    x, ok, h = ug.Solver({"type": "cg", "precond": {"type": "ilu", "ordering": "cmk"}, "convCheck": cc},
                         crs(np.diag([1., 2., 4., 8.]))).apply(np.ones(4))
    assert ok and np.array_equal(x, [1.0, 0.5, 0.25, 0.125])
    with pytest.raises(Exception, match="no diagonal entry"):
        ug.Solver({"type": "cg", "precond": {"type": "ilu"}, "convCheck": cc}, crs(np.array([[0, 1.], [1., 0]]))).apply(np.ones(2))
    prob = pr.Problem(dim=2, num_refs=3)
    x, ok, h = ug.Solver({"type": "gmres", "restart": 5, "precond": {"type": "ilu"}, "convCheck": cc}, prob.matrix()).apply(np.zeros(prob.num_dofs))
    assert ok and len(h) == 1 and h[0] == 0.0 and not x.any()


@pytest.mark.gpu
def test_gpu_golden_ilu_gmres_fixture():
    """GPU vs the committed histories of the reference's ILU / GMRES (no /root/reference needed on the GPU box)."""
    import json
    import ugcore_b200 as ug
    from helpers import make_rhs
    with open(GOLDEN_ILU) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        prob = pr.Problem(**case["problem"])
        pc = case["desc"].get("precond")
        if isinstance(pc, dict) and pc.get("type") == "gmg":
            s = ug.Solver.from_problem(case["desc"], prob)
        else:
            s = ug.Solver(case["desc"], prob.matrix())
        x, ok, h = s.apply(make_rhs(prob, case.get("rhs_seed")))
        ref = np.array(case["history"])
        assert ok == case["converged"], case["name"]
        assert abs(len(h) - len(ref)) <= 1, case["name"]
        # Krylov methods amplify the round-off of another summation order (device reductions, level-scheduled rows)
        tol = 1e-10 if case["desc"]["type"] in ("linear",) or isinstance(pc, dict) and pc.get("type") == "gmg" else 1e-7
        assert rel_hist_err(h, ref) < tol, (case["name"], rel_hist_err(h, ref))
        assert abs(np.linalg.norm(x) - case["solution_norm"]) <= 1e-7 * case["solution_norm"], case["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [0, 2])
def test_gpu_device_resident_bicgstab_equals_host_loop(graph):
    """UG4B200_FLAG_DEVICE_BICGSTAB: rho, alpha, omega, beta and the convergence state stay on the device and an
    iteration is one CUDA graph.  Same kernels, same scalar expressions as the reference-shaped host loop
    (bicgstab.h:161-380) -> identical histories and iterates, also when the check after the half step ends the solve
    and when the step limit is hit."""
    import ugcore_b200 as ug
    from ugcore_b200 import capi
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-1)
    cases = {"gmg_gs": gmg_desc(3, solver="bicgstab", smoother={"type": "gs", "relax": 1.0}, reduction=1e-8),
             "ilu": {"type": "bicgstab", "precond": {"type": "ilu"}, "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}},
             "none": {"type": "bicgstab", "precond": None, "convCheck": {"iterations": 200, "absolute": 1e-12, "reduction": 1e-8}},
             "maxsteps": {"type": "bicgstab", "precond": None, "convCheck": {"iterations": 7, "absolute": 1e-12, "reduction": 1e-30}}}
    for name, desc in cases.items():
        mk = (lambda fl: ug.Solver.from_problem(desc, prob, flags=fl)) if name == "gmg_gs" else (lambda fl: ug.Solver(desc, prob.matrix(), flags=fl))
        x0, ok0, h0 = mk(capi.FLAG_HOST_SCALARS).apply(prob.rhs())
        x1, ok1, h1 = mk(graph).apply(prob.rhs())
        assert ok0 == ok1 and ok0 == (name != "maxsteps"), name
        assert np.array_equal(h0, h1) and np.array_equal(x0, x1), name


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [0, 2])
def test_gpu_device_resident_linear_solver_equals_host_loop(graph):
    """UG4B200_FLAG_DEVICE_LINEAR: LinearSolver (linear_solver.h:114-196) with the convergence state on the device and one
    CUDA graph per iteration — same histories and iterates as the reference-shaped loop, also when the step limit is hit."""
    import ugcore_b200 as ug
    from ugcore_b200 import capi
    prob = pr.Problem(dim=3, num_refs=3)
    cases = {"gmg": gmg_desc(3, solver="linear", reduction=1e-8), "gmg_W": gmg_desc(3, solver="linear", cycle="W", reduction=1e-8),
             "jac": {"type": "linear", "precond": {"type": "jac", "damping": 0.66}, "convCheck": {"iterations": 50, "absolute": 1e-12, "reduction": 1e-30}},
             "ilu": {"type": "linear", "precond": {"type": "ilu"}, "convCheck": {"iterations": 300, "absolute": 1e-12, "reduction": 1e-6}},
             "none": {"type": "linear", "precond": None, "convCheck": {"iterations": 5, "absolute": 1e-12, "reduction": 1e-6}}}
    for name, desc in cases.items():
        mk = (lambda fl: ug.Solver.from_problem(desc, prob, flags=fl)) if name.startswith("gmg") else (lambda fl: ug.Solver(desc, prob.matrix(), flags=fl))
        x0, ok0, h0 = mk(capi.FLAG_HOST_SCALARS).apply(prob.rhs())
        x1, ok1, h1 = mk(graph).apply(prob.rhs())
        assert ok0 == ok1 and ok0 == (name not in ("jac", "none")), name
        assert np.array_equal(h0, h1) and np.array_equal(x0, x1), name


def _restart_desc(rs):
    return {"type": "bicgstab", "restart": rs, "precond": {"type": "jac", "damping": 0.8},
            "convCheck": {"iterations": 200, "absolute": 1e-12, "reduction": 1e-8}}


def test_oracle_bicgstab_periodic_restart(orc, orc_ref):
    """BiCGStab::set_restart (bicgstab.h:161-216): r0, p, v and the scalars are reset every n steps — a different,
    converging iteration; both backends agree bit for bit."""
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2)
    b = np.array(prob.rhs())
    hs = {}
    for rs in (0, 3, 4):
        x, ok, h = oracle.OSolver(orc_ref, _restart_desc(rs), orc_ref.matrix(prob.matrix())).apply(b)
        xp, okp, hp = oracle.OSolver(orc, _restart_desc(rs), orc.matrix(prob.matrix())).apply(b)
        assert ok and okp and np.array_equal(h, hp) and np.array_equal(x, xp)
        hs[rs] = h
    assert len(hs[4]) != len(hs[0]) or not np.array_equal(hs[4], hs[0])


@pytest.mark.gpu
@pytest.mark.parametrize("rs", [0, 3, 4])
def test_gpu_bicgstab_periodic_restart_host_and_device_loop(rs):
    """The restart schedule of the device-resident loop (decided by the host: an iteration is two steps) against the host
    loop (identical) and the oracle."""
    import ugcore_b200 as ug
    from ugcore_b200 import capi
    prob = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2)
    desc = _restart_desc(rs)
    x0, ok0, h0 = ug.Solver(desc, prob.matrix(), flags=capi.FLAG_HOST_SCALARS).apply(prob.rhs())
    x1, ok1, h1 = ug.Solver(desc, prob.matrix()).apply(prob.rhs())
    assert ok0 and ok1 and np.array_equal(h0, h1) and np.array_equal(x0, x1)
    orc = _best()
    osol = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
    xo, oko, ho = osol.apply(np.array(prob.rhs()))
    assert oko and abs(len(h0) - len(ho)) <= 1 and rel_hist_err(h0, ho) < sens_tol(orc, osol, np.array(prob.rhs()))
