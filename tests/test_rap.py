"""Galerkin coarse operators (gmg:set_rap(true)).

CPU: the product's host-side AddMultiplyOf (csrc/host/sparse_util.h) against the REFERENCE's own
AddMultiplyOf compiled into oracle/_ref (algebra_common/sparsematrix_util.h:152-230) and the port —
bit for bit.  GPU: GMG-CG whose level operators are built by RAP at init against the oracle solving
with the oracle's RAP hierarchy.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, rel_hist_err
from ugcore_b200 import problems as pr


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _host_rap(R, A, P):
    from ugcore_b200.capi import check_host, host
    h = C.c_void_p()
    arrs = [np.ascontiguousarray(x.rowptr, np.int64) for x in (R, A, P)]
    cols = [np.ascontiguousarray(x.cols, np.int32) for x in (R, A, P)]
    vals = [np.ascontiguousarray(x.vals, np.float64) for x in (R, A, P)]
    check_host(host.ug4b200_host_rap(R.nrows, A.nrows, _p(arrs[0]), _p(cols[0]), _p(vals[0]), _p(arrs[1]), _p(cols[1]), _p(vals[1]),
                                     _p(arrs[2]), _p(cols[2]), _p(vals[2]), C.byref(h)))
    try:
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check_host(host.ug4b200_io_matrix_info(h, C.byref(nr), C.byref(nc), C.byref(nnz), None, None))
        rp, ci, va = np.zeros(nr.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
        check_host(host.ug4b200_io_matrix_export(h, _p(rp), _p(ci), _p(va), None))
    finally:
        host.ug4b200_io_matrix_free(h)
    return rp, ci, va


@pytest.mark.parametrize("kind", ["poisson2d", "poisson3d", "convdiff3d", "hier3d"])
def test_host_rap_is_bit_identical_to_reference_addmultiplyof(kind, orc, request):
    prob = {"poisson2d": lambda: pr.Problem(dim=2, num_refs=4),
            "poisson3d": lambda: pr.Problem(dim=3, num_refs=3),
            "convdiff3d": lambda: pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF),
            "hier3d": lambda: pr.Problem(dim=3, num_refs=3, order=pr.ORDER_HIER)}[kind]()
    top = prob.num_refs
    A, P, R = prob.matrix(top), prob.prolongation(top), prob.restriction(top)
    rp, ci, va = _host_rap(R, A, P)
    backends = [orc]
    if oracle.have_ref():
        backends.append(oracle.Oracle("ref"))
    for o in backends:
        orp, oci, ova = o.matrix(A).rap(o.matrix(R), o.matrix(P)).export()
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci), o.kind
        assert np.array_equal(va, ova), o.kind
    # sanity: it is the triple product (scipy sums in another order: tolerance)
    S = (R.to_scipy() @ A.to_scipy() @ P.to_scipy()).toarray()
    import scipy.sparse as sp
    assert np.allclose(sp.csr_matrix((va, ci, rp), shape=S.shape).toarray(), S, rtol=1e-13, atol=1e-15)


def test_port_rap_equals_reference_for_blocks(orc):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = oracle.Oracle("ref")
    prob = pr.Problem(dim=3, num_refs=2, problem=pr.ELASTICITY)
    A, P, R = prob.matrix(2), prob.prolongation(2), prob.restriction(2)
    a = orc.matrix(A).rap(orc.matrix(R), orc.matrix(P)).export()
    b = ref.matrix(A).rap(ref.matrix(R), ref.matrix(P)).export()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def _rap_levels(orc, prob, top, base=0):
    """level -> (A, P, R) oracle matrices with A_{l-1} = R_l A_l P_l below the top level"""
    lv = {}
    A = orc.matrix(prob.matrix(top))
    for l in range(top, base - 1, -1):
        P = orc.matrix(prob.prolongation(l)) if l > base else None
        R = orc.matrix(prob.restriction(l)) if l > base else None
        lv[l] = (A, P, R)
        if l > base:
            A = A.rap(R, P)
    return lv


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["poisson3d", "elasticity3d"])
def test_gmg_with_rap_matches_oracle(kind):
    import ugcore_b200 as ug
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    if kind == "poisson3d":
        prob, top = pr.Problem(dim=3, num_refs=4), 4
    else:
        prob, top = pr.Problem(dim=3, num_refs=3, problem=pr.ELASTICITY), 3
    desc = gmg_desc(top)
    desc["precond"] = dict(desc["precond"], rap=True)
    s = ug.Solver.from_problem(desc, prob)
    x, ok, h = s.apply(prob.rhs())
    lv = _rap_levels(orc, prob, top)
    xo, oko, ho = oracle.OSolver(orc, desc, lv[top][0], lv).apply(np.array(prob.rhs()))
    assert ok and oko
    assert abs(len(h) - len(ho)) <= 1
    assert rel_hist_err(h, ho) < 1e-10          # north_star: residual history within 1e-10 relative per iteration
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-9
    # and the Galerkin hierarchy really differs from the re-discretised one
    x2, ok2, h2 = ug.Solver.from_problem(gmg_desc(top), prob).apply(prob.rhs())
    assert ok2 and not np.array_equal(np.array(h), np.array(h2))
