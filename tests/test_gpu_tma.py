"""GPU parity of the bulk-copy (TMA) staged SpMV kernel, forced on for small matrices:
same bit-exact bar as the register-staged kernel (tests/test_gpu_kernels.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import Dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["x_staged", "value_indexed", "plain"])
def tma_ctx(request):
    """x_staged: UG4B200_XSTAGE=1 — value-indexed words + x operand staged in shared memory (spmv1_xs_kernel; by default
    only built where the 16-bit column window of the value-indexed stream does not fit, e.g. 257^3); value_indexed:
    spmv1_vi_kernel, x gathered from global memory; plain: UG4B200_NO_COMPRESS=1 (spmv1_tma_kernel, 12 B per entry)."""
    from ugcore_b200 import capi
    os.environ["UG4B200_TMA_MIN_SLICES"] = "0"
    if request.param == "plain":
        os.environ["UG4B200_NO_COMPRESS"] = "1"
    if request.param == "x_staged":
        os.environ["UG4B200_XSTAGE"] = "1"
    ctx = C.c_void_p()
    try:
        capi.check(capi.dev.ug4b200_ctx_create(0, None, C.byref(ctx)))
    finally:
        del os.environ["UG4B200_TMA_MIN_SLICES"]
        os.environ.pop("UG4B200_NO_COMPRESS", None)
        os.environ.pop("UG4B200_XSTAGE", None)
    ctx.mode = request.param
    yield ctx
    capi.dev.ug4b200_ctx_destroy(ctx)


def _info(D, dA):
    from ugcore_b200 import capi
    info = capi.MatrixInfo()
    D.dev.ug4b200_matrix_get_info(dA, C.byref(info))
    return info


@pytest.fixture()
def D(tma_ctx):
    d = Dev(tma_ctx)
    yield d
    d.free_all()


def _cases():
    from ugcore_b200 import problems as pr
    return [pr.Problem(dim=3, num_refs=4), pr.Problem(dim=3, num_refs=3, order=pr.ORDER_HIER),
            pr.Problem(dim=2, num_refs=6), pr.Problem(dim=3, num_refs=4, problem=pr.CONVDIFF, eps=1e-2),
            pr.Problem(dim=3, num_refs=2, base=(3, 1, 2))]


@pytest.mark.parametrize("i", range(5))
def test_tma_spmv_family_bit_exact(D, orc, i):
    from ugcore_b200 import capi
    prob = _cases()[i]
    rng = np.random.default_rng(i)
    A = prob.matrix()
    n = A.nrows
    oA, dA = orc.matrix(A), D.matrix(A)
    info = _info(D, dA)
    # the x-staged stream exists exactly for the banded (lexicographic) numberings, and only in its mode
    if D.ctx.mode != "x_staged":
        assert not info.x_staged
    elif i != 1:      # (the hierarchical numbering scatters a slice's columns: usually not stageable, not asserted)
        assert info.x_staged and info.value_indexed, (i, info.x_staged, info.value_indexed, info.num_distinct_values)
    x, y0, v = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    dx, dv = D.up(x), D.up(v)
    dy = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, dy, dx, 1))
    assert np.array_equal(D.down(dy, n), oA.apply(x))
    dy = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_matmul_minus(D.ctx, dA, dy, dx, 1))
    assert np.array_equal(D.down(dy, n), oA.matmul_minus(y0, x))
    dd = D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 0.7, dv, -1.3, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.axpy(0.7, v, -1.3, x))
    dd = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 0.0, None, 0.25, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.axpy(0.0, None, 0.25, x))
    # fused smoother
    dinv = D.alloc(n * 8)
    D.chk(D.dev.ug4b200_jacobi_prepare(D.ctx, dA, 0.66, 1, dinv))
    sc = rng.standard_normal(n)
    dsd, dsc, dout = D.up(y0), D.up(sc), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_jacobi_smooth_fused(D.ctx, dA, dinv, dsd, dx, dout, dsc, capi.SMOOTH_ADD_IN | capi.SMOOTH_JACOBI))
    sd_ref = oA.matmul_minus(y0, x)
    assert np.array_equal(D.down(dsd, n), sd_ref)
    assert np.array_equal(D.down(dout, n), oA.jacobi(sd_ref, 0.66))
    assert np.array_equal(D.down(dsc, n), sc + x)
    # fused dot
    S = D.alloc(64)
    fin = capi.Fin(capi.FIN_STORE, S, None, None, None)
    dq = D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_matrix_apply_dot_ds(D.ctx, dA, dq, dx, fin))
    q = oA.apply(x)
    assert np.array_equal(D.down(dq, n), q)
    assert abs(D.down(S, 1)[0] - orc.dot(q, x)) <= 1e-13 * np.sum(np.abs(q * x))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)


def test_tma_ragged_and_empty_rows(D, orc):
    """Random ragged pattern incl. empty rows, empty slices and rows longer than a chunk."""
    from ugcore_b200.problems import Crs
    rng = np.random.default_rng(42)
    n, m = 1000, 777
    lens = rng.integers(0, 40, n)
    lens[100:170] = 0           # two completely empty slices
    lens[5] = 0
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ci = np.concatenate([np.sort(rng.choice(m, l, replace=False)) for l in lens]).astype(np.int32)
    va = rng.standard_normal(ci.size)
    A = Crs(n, m, 1, rp, ci, va)
    x, d0 = rng.standard_normal(m), rng.standard_normal(n)
    oA, dA = orc.matrix(A), D.matrix(A)
    dx, dd = D.up(x), D.up(d0)
    D.chk(D.dev.ug4b200_matrix_apply_ignore_zero_rows(D.ctx, dA, dd, 1.0, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.apply_ignore_zero_rows(d0, 1.0, x))
    D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, dd, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.apply(x))
    dd = D.up(d0)
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 1.0, dd, 0.37, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.axpy(1.0, None, 0.37, x, dest=d0))


def test_tma_transfers_and_banded_ragged_matrix(D, orc):
    """The x-staged path on what is not a level operator: P (rows of 1 / 2 / 4 / 8 entries, rectangular, odd number of
    columns), R = P^T with Dirichlet rows, and a ragged banded matrix with few distinct values, empty rows, an empty
    slice and an odd number of columns (the last run of the last slices ends on the 16-byte unit behind the vector)."""
    from ugcore_b200 import problems as pr
    from ugcore_b200.problems import Crs
    prob = pr.Problem(dim=3, num_refs=4)
    rng = np.random.default_rng(9)
    for M in (prob.prolongation(4), prob.restriction(4), prob.prolongation(3)):
        oM, dM = orc.matrix(M), D.matrix(M)
        x, d0 = rng.standard_normal(M.ncols), rng.standard_normal(M.nrows)
        dx, dd = D.up(x), D.up(d0)
        D.chk(D.dev.ug4b200_matrix_apply_ignore_zero_rows(D.ctx, dM, dd, 1.0, dx, 1))
        assert np.array_equal(D.down(dd, M.nrows), oM.apply_ignore_zero_rows(d0, 1.0, x))
        D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dM, dd, 0.0, None, 1.0, dx, 1))
        assert np.array_equal(D.down(dd, M.nrows), oM.apply(x))
        D.dev.ug4b200_matrix_destroy(D.ctx, dM)
    n, m, band = 2001, 2003, 40
    lens = rng.integers(0, 20, n)
    lens[320:352] = 0
    lens[7] = 0
    vals_pool = np.array([1.0, -0.5, 0.25, 3.0, -0.0, 0.0, 1e-300, -7.5])
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ci = np.concatenate([np.sort(rng.choice(np.arange(max(0, r - band), min(m, r + band)), l, replace=False))
                         for r, l in enumerate(lens)]).astype(np.int32)
    ci[-1] = m - 1 if lens[-1] else ci[-1]
    va = vals_pool[rng.integers(0, vals_pool.size, ci.size)]
    A = Crs(n, m, 1, rp, ci, va)
    oA, dA = orc.matrix(A), D.matrix(A)
    if D.ctx.mode == "x_staged":
        assert _info(D, dA).x_staged == 1
    x, d0 = rng.standard_normal(m), rng.standard_normal(n)
    dx, dd = D.up(x), D.up(d0)
    D.chk(D.dev.ug4b200_matrix_apply_ignore_zero_rows(D.ctx, dA, dd, -1.0, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.apply_ignore_zero_rows(d0, -1.0, x))
    dd = D.up(d0)
    D.chk(D.dev.ug4b200_matrix_matmul_minus(D.ctx, dA, dd, dx, 1))
    assert np.array_equal(D.down(dd, n), oA.matmul_minus(d0, x))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)
