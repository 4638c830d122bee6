"""CPU suite: pins the oracle.

1. golden vector of the reference's own test-suite: tests/ref/sm_transpose.out (nnz of
   set_as_transpose_of / set_as_transpose_of2 on the matrices built by tests/sm_transpose.cc);
2. FV1 known-answer stencils;
3. port backend == compiled reference templates (oracle/_ref), bit for bit, kernel by kernel
   and for whole solves (only where oracle/_ref was built, i.e. /root/reference was present);
4. committed golden residual histories (tests/golden/, generated with oracle/_ref).
"""
import json
import os

import numpy as np
import pytest

from helpers import gmg_desc, make_rhs, oracle_levels
from ugcore_b200 import problems as pr
from ugcore_b200.problems import Crs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "residual_histories.json")


def _from_dict(n, m, ent):
    """CRS with sorted rows from {(r, c): v} (what SparseMatrix::operator()(r,c) = v builds)."""
    rows = [[] for _ in range(n)]
    for (r, c), v in ent.items():
        rows[r].append((c, v))
    rp, ci, va = [0], [], []
    for r in rows:
        r.sort()
        ci += [c for c, _ in r]
        va += [v for _, v in r]
        rp.append(len(ci))
    return Crs(n, m, 1, np.array(rp, np.int64), np.array(ci, np.int32), np.array(va, float))


def _sm_transpose_test0(N, M):
    """tests/sm_transpose.cc:83-125"""
    e = {}
    for i in range(1, min(N, M)):
        e[(i, i)] = 2.0
        e[(i, i - 1)] = 1.0
        e[(i - 1, i)] = 1.0
    if N > 6 and M > 6:
        e[(5, 5)] = 0.0
        e[(5, 0)] = 0.0
        e[(1, 6)] = 1.0
    return _from_dict(N, M, e)


def _sm_transpose_test3(N):
    """tests/sm_transpose.cc:127-160"""
    return _from_dict(N, N, {(i, j): float(i + j) for i in range(N) for j in range(N)})


# tests/ref/sm_transpose.out: (nnz set_as_transpose_of, nnz set_as_transpose_of2)
SM_TRANSPOSE_GOLDEN = [
    (lambda: _sm_transpose_test0(2, 2), (2, 2), 3, 3),
    (lambda: _sm_transpose_test0(10000, 10000), (10000, 10000), 29999, 29997),
    (lambda: _sm_transpose_test0(7, 10), (10, 7), 20, 18),
    (lambda: _sm_transpose_test0(10, 7), (7, 10), 20, 18),
    (lambda: _sm_transpose_test3(10), (10, 10), 100, 99),
    (lambda: _sm_transpose_test3(300), (300, 300), 90000, 89999),
]


@pytest.mark.parametrize("case", range(len(SM_TRANSPOSE_GOLDEN)))
@pytest.mark.parametrize("kind", ["port", "ref"])
def test_sm_transpose_golden(case, kind, request):
    orc = request.getfixturevalue("orc" if kind == "port" else "orc_ref")
    make, shape, nnz_of, nnz_of2 = SM_TRANSPOSE_GOLDEN[case]
    A = orc.matrix(make())
    B, C2 = A.transpose(keep_zeros=True), A.transpose(keep_zeros=False)
    assert (B.nrows, B.ncols) == shape and (C2.nrows, C2.ncols) == shape
    assert B.nnz == nnz_of and C2.nnz == nnz_of2
    # is_equal(B, C) of the reference test: equal up to explicit zeros
    rb, cb, vb = B.export()
    rc, cc, vc = C2.export()
    import scipy.sparse as sp
    Mb = sp.csr_matrix((vb, cb, rb), shape=shape); Mb.eliminate_zeros()
    Mc = sp.csr_matrix((vc, cc, rc), shape=shape); Mc.eliminate_zeros()
    assert (Mb != Mc).nnz == 0
    # and it is the transpose
    rp, ci, va = A.export()
    Ma = sp.csr_matrix((va, ci, rp), shape=(A.nrows, A.ncols))
    assert abs(Mb - Ma.T).max() == 0


def test_fv1_known_answer_stencils():
    """h*{27/8, -3/16, -5/32, -3/64} (3-D hex) and {3, -1/2, -1/4} (2-D quad), exact in fp64."""
    p = pr.Problem(dim=3, num_refs=3)
    A = p.matrix().to_scipy()
    h, n = 1 / 8, 9
    mid = 4 + n * 4 + n * n * 4
    row = A.getrow(mid).toarray().ravel()
    vals = {}
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                vals.setdefault(abs(di) + abs(dj) + abs(dk), set()).add(row[mid + di + n * dj + n * n * dk])
    assert vals == {0: {h * 27 / 8}, 1: {-h * 3 / 16}, 2: {-h * 5 / 32}, 3: {-h * 3 / 64}}
    assert row.sum() == 0.0
    assert abs(A - A.T)[np.ix_(p.dirichlet() == 0, p.dirichlet() == 0)].max() == 0
    p2 = pr.Problem(dim=2, num_refs=3)
    A2 = p2.matrix().to_scipy()
    r = A2.getrow(4 + 9 * 4).toarray().ravel()
    assert sorted(set(r[r != 0])) == [-0.5, -0.25, 3.0]
    # Dirichlet rows: identity with the pattern retained (explicit zeros stored)
    c = p.matrix()
    d0 = np.flatnonzero(p.dirichlet())[0]
    seg = slice(c.rowptr[d0], c.rowptr[d0 + 1])
    assert c.rowptr[d0 + 1] - c.rowptr[d0] == 8 and np.sum(c.vals[seg]) == 1.0 and c.vals[seg][c.cols[seg] == d0] == 1.0


def test_sizes_match_survey():
    p = pr.Problem(dim=3, num_refs=5)
    assert [p.matrix(l).nrows for l in range(6)] == [8, 27, 125, 729, 4913, 35937]
    assert p.matrix(5).nnz == (3 * 33 - 2) ** 3


KERNEL_PROBLEMS = [
    dict(dim=3, num_refs=3), dict(dim=3, num_refs=3, order=1), dict(dim=2, num_refs=4),
    dict(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2), dict(dim=3, num_refs=2, problem=pr.ELASTICITY),
    dict(dim=2, num_refs=3, problem=pr.ELASTICITY),
]


@pytest.mark.parametrize("pargs", KERNEL_PROBLEMS, ids=lambda d: "-".join(f"{k}{v}" for k, v in d.items()))
def test_port_equals_compiled_reference_kernels(orc, orc_ref, pargs):
    rng = np.random.default_rng(0)
    p = pr.Problem(**pargs)
    A = p.matrix()
    b = A.block
    n = A.nrows * b
    x, y, v = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    Ap, Ar = orc.matrix(A), orc_ref.matrix(A)
    assert np.array_equal(Ap.apply(x), Ar.apply(x))
    assert np.array_equal(Ap.matmul_minus(y, x), Ar.matmul_minus(y, x))
    assert np.array_equal(Ap.axpy(0.7, v, -1.3, x), Ar.axpy(0.7, v, -1.3, x))
    assert np.array_equal(Ap.axpy(0.5, None, 2.0, x, dest=y), Ar.axpy(0.5, None, 2.0, x, dest=y))
    assert np.array_equal(Ap.axpy(0.0, None, 0.25, x), Ar.axpy(0.0, None, 0.25, x))
    for damp, blk in ((0.66, True), (0.8, False)):
        assert np.array_equal(Ap.jacobi(x, damp, blk), Ar.jacobi(x, damp, blk))
    for kind in ("ll", "ur", "sgs"):
        assert np.array_equal(Ap.gs(x, kind, 0.9), Ar.gs(x, kind, 0.9))
    assert orc.dot(x, y, b) == orc_ref.dot(x, y, b)
    assert orc.norm(x, b) == orc_ref.norm(x, b)
    assert np.array_equal(orc.scale_add2(0.3, x, -1.7, y), orc_ref.scale_add2(0.3, x, -1.7, y))
    assert np.array_equal(orc.scale_add3(1.0, x, 0.25, y, -0.5, v), orc_ref.scale_add3(1.0, x, 0.25, y, -0.5, v))
    top = p.num_refs
    P, R = p.prolongation(top), p.restriction(top)
    xc, xf = rng.standard_normal(P.ncols * b), rng.standard_normal(P.nrows * b)
    assert np.array_equal(orc.matrix(P).axpy(0.0, None, 1.0, xc, vblock=b), orc_ref.matrix(P).axpy(0.0, None, 1.0, xc, vblock=b))
    c0 = rng.standard_normal(R.nrows * b)
    assert np.array_equal(orc.matrix(R).apply_ignore_zero_rows(c0, 1.0, xf, vblock=b),
                          orc_ref.matrix(R).apply_ignore_zero_rows(c0, 1.0, xf, vblock=b))
    base = pr.Problem(**{**pargs, "num_refs": 1}).matrix()
    bb = rng.standard_normal(base.nrows * b)
    assert np.allclose(orc.matrix(base).lu_solve(bb), orc_ref.matrix(base).lu_solve(bb), rtol=1e-12, atol=1e-14)


def test_restriction_is_transpose_of_prolongation_with_dirichlet_adjust(orc):
    p = pr.Problem(dim=3, num_refs=3)
    P, R = p.prolongation(3), p.restriction(3)
    PT = orc.matrix(P).transpose(keep_zeros=True)
    rp, ci, va = PT.export()
    assert np.array_equal(rp, R.rowptr) and np.array_equal(ci, R.cols)  # same pattern, zeros kept
    cd = p.dirichlet(2) != 0
    rows = np.repeat(np.arange(R.nrows), np.diff(R.rowptr))
    assert np.array_equal(va[~cd[rows]], R.vals[~cd[rows]])            # interior coarse rows: exactly P^T
    for c in np.flatnonzero(cd)[:50]:                                   # Dirichlet coarse rows: injection
        seg = R.vals[R.rowptr[c]:R.rowptr[c + 1]]
        assert seg.sum() == 1.0 and set(seg) <= {0.0, 1.0}


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_golden_histories(kind, request):
    """Both oracle backends reproduce the committed fixtures (bit for bit: they were generated
    by oracle/_ref and the port is bit-identical to it)."""
    import oracle
    orc = request.getfixturevalue("orc" if kind == "port" else "orc_ref")
    with open(GOLDEN) as f:
        gold = json.load(f)
    assert gold["oracle_backend"] == "ref"
    for case in gold["cases"]:
        prob = pr.Problem(**case["problem"])
        desc = case["desc"]
        pc = desc.get("precond")
        if isinstance(pc, dict) and pc.get("type") == "gmg":
            lv = oracle_levels(orc, prob, pc["baseLevel"], pc["topLevel"])
            s = oracle.OSolver(orc, desc, lv[pc["topLevel"]][0], lv)
        else:
            s = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
        x, ok, h = s.apply(make_rhs(prob, case.get("rhs_seed")))
        assert ok == case["converged"]
        assert np.array_equal(h, np.array(case["history"])), case["name"]
        assert np.linalg.norm(x) == case["solution_norm"], case["name"]


def test_oracle_invariants(orc):
    """Cross-checks that do not need the reference: defect returned by
    apply_return_defect equals a fresh b - A x; the V(2,2)-Jacobi GMG with exact base solve is a
    symmetric operator on the interior; GMG iteration counts are h-independent."""
    import oracle
    its = []
    for refs in (3, 4, 5):
        p = pr.Problem(dim=3, num_refs=refs)
        lv = oracle_levels(orc, p)
        s = oracle.OSolver(orc, gmg_desc(refs), lv[refs][0], lv)
        b = make_rhs(p, 5)
        x, ok, h = s.apply(b)
        assert ok
        its.append(len(h) - 1)
        A = p.matrix().to_scipy()
        assert abs(np.linalg.norm(b - A @ x) - h[-1]) < 1e-9 * h[0]
    assert max(its) - min(its) <= 1
    p = pr.Problem(dim=3, num_refs=3)
    lv = oracle_levels(orc, p)
    s = oracle.OSolver(orc, gmg_desc(3), lv[3][0], lv)
    u, v = make_rhs(p, 1), make_rhs(p, 2)
    Mu, Mv = s.precond_apply(u), s.precond_apply(v)
    assert abs(v @ Mu - u @ Mv) < 1e-12 * abs(v @ Mu)
