"""GPU parity, kernel level: every C-ABI kernel against the CPU oracle on the same inputs.

Bar: SpMV family, Jacobi, Gauss-Seidel sweeps, LU apply and element-wise vector kernels are
BIT-EXACT (same operation order, no FMA contraction); reductions (tree vs sequential sum)
within 1e-13 relative.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import Dev

pytestmark = pytest.mark.gpu


@pytest.fixture()
def D(gpu_ctx):
    d = Dev(gpu_ctx)
    yield d
    d.free_all()


def _problems():
    from ugcore_b200 import problems as pr
    return [
        ("poisson3d_lex", pr.Problem(dim=3, num_refs=3)),
        ("poisson3d_hier", pr.Problem(dim=3, num_refs=3, order=pr.ORDER_HIER)),
        ("poisson2d", pr.Problem(dim=2, num_refs=5)),
        ("convdiff3d", pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2)),
        ("elasticity3d", pr.Problem(dim=3, num_refs=2, problem=pr.ELASTICITY)),
        ("elasticity2d", pr.Problem(dim=2, num_refs=3, problem=pr.ELASTICITY)),
        ("poisson3d_ragged", pr.Problem(dim=3, num_refs=2, base=(3, 1, 2))),
    ]


@pytest.fixture(scope="module", params=range(7), ids=lambda i: _problems()[i][0])
def prob(request):
    return _problems()[request.param][1]


def test_spmv_family_bit_exact(D, orc, prob):
    rng = np.random.default_rng(1)
    A = prob.matrix()
    b = A.block
    oA = orc.matrix(A)
    dA = D.matrix(A)
    n = A.nrows * b
    x = rng.standard_normal(n)
    y0 = rng.standard_normal(n)
    dx, dy = D.up(x), D.up(y0)
    # apply
    D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, dy, dx, b))
    assert np.array_equal(D.down(dy, n), oA.apply(x))
    # matmul_minus
    dy = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_matmul_minus(D.ctx, dA, dy, dx, b))
    assert np.array_equal(D.down(dy, n), oA.matmul_minus(y0, x))
    # general axpy: alpha, beta arbitrary, separate v
    v = rng.standard_normal(n)
    dv, dd = D.up(v), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 0.7, dv, -1.3, dx, b))
    assert np.array_equal(D.down(dd, n), oA.axpy(0.7, v, -1.3, x))
    # in-place with alpha != 1
    dd = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 0.5, dd, 2.0, dx, b))
    assert np.array_equal(D.down(dd, n), oA.axpy(0.5, None, 2.0, x, dest=y0))
    # alpha == 0, beta general
    dd = D.up(y0)
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dA, dd, 0.0, None, 0.25, dx, b))
    assert np.array_equal(D.down(dd, n), oA.axpy(0.0, None, 0.25, x))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)


def test_transfers_bit_exact(D, orc, prob):
    rng = np.random.default_rng(2)
    top = prob.num_refs
    vb = prob.block
    P, R = prob.prolongation(top), prob.restriction(top)
    oP, oR = orc.matrix(P), orc.matrix(R)
    dP, dR = D.matrix(P), D.matrix(R)
    xc = rng.standard_normal(P.ncols * vb)
    xf = rng.standard_normal(P.nrows * vb)
    # prolongate: axpy(uFine, 0, uFine, dampProl, uCoarse)
    dxc, dxf = D.up(xc), D.up(np.zeros(P.nrows * vb))
    D.chk(D.dev.ug4b200_matrix_axpy(D.ctx, dP, dxf, 0.0, None, 1.0, dxc, vb))
    assert np.array_equal(D.down(dxf, P.nrows * vb), oP.axpy(0.0, None, 1.0, xc, vblock=vb))
    # restrict: apply_ignore_zero_rows
    c0 = rng.standard_normal(R.nrows * vb)
    dc, df = D.up(c0), D.up(xf)
    D.chk(D.dev.ug4b200_matrix_apply_ignore_zero_rows(D.ctx, dR, dc, 1.0, df, vb))
    assert np.array_equal(D.down(dc, R.nrows * vb), oR.apply_ignore_zero_rows(c0, 1.0, xf, vblock=vb))
    for m in (dP, dR):
        D.dev.ug4b200_matrix_destroy(D.ctx, m)


def test_apply_ignore_zero_rows_leaves_empty_rows(D, orc):
    """Rows without connections keep dest (sparsematrix_impl.h:271-288); plain apply zeroes them."""
    from ugcore_b200.problems import Crs
    rp = np.array([0, 2, 2, 3, 3, 3], dtype=np.int64)
    ci = np.array([0, 3, 1], dtype=np.int32)
    va = np.array([2.0, -1.0, 0.0])
    A = Crs(5, 4, 1, rp, ci, va)
    x = np.array([1.0, 2.0, 3.0, 4.0])
    d0 = np.array([9.0, 8.0, 7.0, 6.0, 5.0])
    dA, dx, dd = D.matrix(A), D.up(x), D.up(d0)
    D.chk(D.dev.ug4b200_matrix_apply_ignore_zero_rows(D.ctx, dA, dd, 1.0, dx, 1))
    got = D.down(dd, 5)
    assert np.array_equal(got, orc.matrix(A).apply_ignore_zero_rows(d0, 1.0, x))
    assert np.array_equal(got, [-2.0, 8.0, 0.0, 6.0, 5.0])
    D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, dd, dx, 1))
    assert np.array_equal(D.down(dd, 5), [-2.0, 0.0, 0.0, 0.0, 0.0])


def test_empty_matrix_and_vectors(D):
    from ugcore_b200.problems import Crs
    A = Crs(0, 0, 1, np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0))
    dA = D.matrix(A)
    p = D.alloc(8)
    D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, p, D.alloc(8), 1))
    D.chk(D.dev.ug4b200_vec_set(D.ctx, 0, p, 1.0))
    r = C.c_double(-1)
    D.chk(D.dev.ug4b200_vec_dot(D.ctx, 0, p, p, C.byref(r)))
    assert r.value == 0.0


@pytest.mark.parametrize("n", [1, 2, 31, 1000, 100003])
def test_vector_ops(D, orc, n):
    rng = np.random.default_rng(n)
    a, b, c = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    da, db, dc, dd = D.up(a), D.up(b), D.up(c), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_vec_scale_add2(D.ctx, n, dd, 0.3, da, -1.7, db))
    assert np.array_equal(D.down(dd, n), orc.scale_add2(0.3, a, -1.7, b))
    D.chk(D.dev.ug4b200_vec_scale_add3(D.ctx, n, dd, 1.0, da, 0.25, db, -0.5, dc))
    assert np.array_equal(D.down(dd, n), orc.scale_add3(1.0, a, 0.25, b, -0.5, c))
    # aliasing: x = 1.0*x + alpha*p
    D.chk(D.dev.ug4b200_vec_scale_add2(D.ctx, n, da, 1.0, da, 0.37, db))
    assert np.array_equal(D.down(da, n), orc.scale_add2(1.0, a, 0.37, b))
    a = D.down(da, n)
    D.chk(D.dev.ug4b200_vec_add(D.ctx, n, da, db))
    assert np.array_equal(D.down(da, n), a + b)
    D.chk(D.dev.ug4b200_vec_sub(D.ctx, n, da, dc))
    assert np.array_equal(D.down(da, n), (a + b) - c)
    D.chk(D.dev.ug4b200_vec_scale(D.ctx, n, da, 0.1))
    assert np.array_equal(D.down(da, n), ((a + b) - c) * 0.1)
    # unaligned views (odd offset) must not take the 16-byte path
    if n > 3:
        off = C.c_void_p(dd.value + 8)
        D.chk(D.dev.ug4b200_vec_set(D.ctx, n - 1, off, 2.5))
        assert np.array_equal(D.down(dd, n)[1:], np.full(n - 1, 2.5))
    r = C.c_double()
    D.chk(D.dev.ug4b200_vec_dot(D.ctx, n, db, dc, C.byref(r)))
    ref = orc.dot(b, c)
    assert abs(r.value - ref) <= 1e-13 * np.sum(np.abs(b * c))
    D.chk(D.dev.ug4b200_vec_norm(D.ctx, n, db, C.byref(r)))
    assert abs(r.value - orc.norm(b)) <= 1e-13 * orc.norm(b)


def test_gather_scatter(D):
    rng = np.random.default_rng(5)
    n, blk = 1000, 3
    perm = rng.permutation(n).astype(np.int32)
    src = rng.standard_normal(n * blk)
    ds, dd, di = D.up(src), D.up(np.zeros(n * blk)), D.up(perm, np.int32)
    D.chk(D.dev.ug4b200_vec_gather(D.ctx, n, blk, dd, ds, di))
    assert np.array_equal(D.down(dd, n * blk).reshape(n, blk), src.reshape(n, blk)[perm])
    D.chk(D.dev.ug4b200_vec_scatter(D.ctx, n, blk, dd, di, ds))
    exp = np.zeros((n, blk)); exp[perm] = src.reshape(n, blk)
    assert np.array_equal(D.down(dd, n * blk).reshape(n, blk), exp)
    D.chk(D.dev.ug4b200_vec_scatter_add(D.ctx, n, blk, dd, di, ds))
    assert np.array_equal(D.down(dd, n * blk).reshape(n, blk), exp + exp)


def test_jacobi_bit_exact(D, orc, prob):
    rng = np.random.default_rng(3)
    A = prob.matrix()
    b = A.block
    n = A.nrows * b
    d = rng.standard_normal(n)
    oA, dA = orc.matrix(A), D.matrix(A)
    dinv, dd, dc = D.alloc(A.nrows * b * b * 8), D.up(d), D.up(np.zeros(n))
    for damp, blockinv in ((0.66, 1), (1.0, 1), (0.8, 0)):
        D.chk(D.dev.ug4b200_jacobi_prepare(D.ctx, dA, damp, blockinv, dinv))
        D.chk(D.dev.ug4b200_jacobi_step(D.ctx, A.nrows, b, dinv, dc, dd))
        assert np.array_equal(D.down(dc, n), oA.jacobi(d, damp, bool(blockinv)))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)


def test_fused_smoother_equals_unfused_sequence(D, orc, prob):
    """sc += st_in; sd -= A st_in; st_out = Dinv sd  ==  the reference's three calls, bit for bit."""
    from ugcore_b200 import capi
    rng = np.random.default_rng(4)
    A = prob.matrix()
    b = A.block
    n = A.nrows * b
    sd, st, sc = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    oA, dA = orc.matrix(A), D.matrix(A)
    dinv = D.alloc(A.nrows * b * b * 8)
    D.chk(D.dev.ug4b200_jacobi_prepare(D.ctx, dA, 0.66, 1, dinv))
    dsd, dst, dsc, dout = D.up(sd), D.up(st), D.up(sc), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_jacobi_smooth_fused(D.ctx, dA, dinv, dsd, dst, dout, dsc,
                                            capi.SMOOTH_ADD_IN | capi.SMOOTH_JACOBI))
    sd_ref = oA.matmul_minus(sd, st)
    assert np.array_equal(D.down(dsd, n), sd_ref)
    assert np.array_equal(D.down(dout, n), oA.jacobi(sd_ref, 0.66))
    assert np.array_equal(D.down(dsc, n), sc + st)
    # serial variant: JACOBI | ADD_OUT
    dsd, dsc = D.up(sd), D.up(sc)
    D.chk(D.dev.ug4b200_jacobi_smooth_fused(D.ctx, dA, dinv, dsd, dst, dout, dsc,
                                            capi.SMOOTH_JACOBI | capi.SMOOTH_ADD_OUT))
    assert np.array_equal(D.down(dsc, n), sc + oA.jacobi(sd_ref, 0.66))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)


def _color_sorted(A):
    """Greedy colouring through the library's host helper -> colour-sorted permuted CRS."""
    from ugcore_b200 import capi
    from ugcore_b200.problems import Crs
    n = A.nrows
    color = np.zeros(n, np.int32)
    nc = C.c_int()
    capi.dev.ug4b200_color_greedy(n, A.rowptr.ctypes.data_as(C.c_void_p), A.cols.ctypes.data_as(C.c_void_p),
                                  color.ctypes.data_as(C.c_void_p), C.byref(nc))
    order = np.argsort(color, kind="stable")
    perm = np.empty(n, np.int64); perm[order] = np.arange(n)
    cptr = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=nc.value))]).astype(np.int64)
    M = A.to_scipy() if A.block == 1 else None
    bb = A.block * A.block
    rows = np.repeat(np.arange(n), np.diff(A.rowptr))
    pr, pc = perm[rows], perm[A.cols]
    key = np.lexsort((pc, pr))
    rp = np.concatenate([[0], np.cumsum(np.bincount(pr, minlength=n))]).astype(np.int64)
    vals = A.vals.reshape(-1, bb)[key].ravel().copy()
    PA = Crs(n, n, A.block, rp, pc[key].astype(np.int32), vals)
    return PA, perm, cptr


@pytest.mark.parametrize("kind,name", [(0, "ll"), (1, "ur"), (2, "sgs")])
def test_multicolor_gs_bit_exact(D, orc, prob, kind, name):
    """Multicolour GS == the reference's lexicographic sweep over the colour-sorted matrix."""
    from ugcore_b200 import capi
    rng = np.random.default_rng(6)
    PA, perm, cptr = _color_sorted(prob.matrix())
    assert capi.dev.ug4b200_color_check(PA.nrows, PA.rowptr.ctypes.data_as(C.c_void_p),
                                        PA.cols.ctypes.data_as(C.c_void_p), cptr.size - 1,
                                        cptr.ctypes.data_as(C.c_void_p)) == 0
    b = PA.block
    n = PA.nrows * b
    d = rng.standard_normal(n)
    oA, dA = orc.matrix(PA), D.matrix(PA)
    dd, dc = D.up(d), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_gs_step(D.ctx, dA, cptr.size - 1, cptr.ctypes.data_as(C.c_void_p), kind, 0.9, dc, dd))
    assert np.array_equal(D.down(dc, n), oA.gs(d, name, 0.9))
    D.dev.ug4b200_matrix_destroy(D.ctx, dA)


def test_lu_apply_bit_exact(D, orc):
    from ugcore_b200 import problems as pr
    rng = np.random.default_rng(7)
    for prob in (pr.Problem(dim=3, num_refs=1), pr.Problem(dim=2, num_refs=3),
                 pr.Problem(dim=3, num_refs=1, problem=pr.ELASTICITY)):
        A = prob.matrix()
        n = A.nrows * A.block
        M = A.to_scipy().toarray()
        # host LU exactly as the host layer does (LUDecomp, no_lapack/lu_decomp.h:45-75)
        a = M.copy(); piv = np.zeros(n, np.int32)
        for k in range(n):
            big = k
            for j in range(k + 1, n):
                if abs(a[big, k]) < abs(a[j, k]):
                    big = j
            if big != k:
                a[[k, big]] = a[[big, k]]
            piv[k] = big
            for i in range(k + 1, n):
                a[i, k] = a[i, k] / a[k, k]
                for j in range(k + 1, n):
                    a[i, j] = a[i, j] - a[i, k] * a[k, j]
        bvec = rng.standard_normal(n)
        dlu, dpiv, db, dx = D.up(a.ravel()), D.up(piv, np.int32), D.up(bvec), D.up(np.zeros(n))
        D.chk(D.dev.ug4b200_lu_apply(D.ctx, n, dlu, dpiv, dx, db))
        assert np.array_equal(D.down(dx, n), orc.matrix(A).lu_solve(bvec))


def test_lu_apply_large_base_grid_parallel_order(D, orc):
    """Above 128 unknowns the backward substitution runs column-oriented (parallel; SolveLU's own row order is one serial
    chain of n^2/2 operations): same terms per row, other order -> equal to the reference to round-off, not bit for bit.
    Stand-alone kernel and the recorded (batch) form give the same bits."""
    from ugcore_b200 import problems as pr
    import scipy.linalg
    prob = pr.Problem(dim=3, num_refs=1, base=(3, 3, 3), problem=pr.ELASTICITY)     # 7^3 nodes x 3 = 1029 unknowns
    A = prob.matrix()
    n = A.nrows * A.block
    M = A.to_scipy().toarray()
    lu, piv = scipy.linalg.lu_factor(M)           # any valid LU with row interchanges serves as input of the apply kernel
    bvec = np.random.default_rng(3).standard_normal(n)
    dlu, dpiv, db, dx = D.up(lu.ravel()), D.up(piv.astype(np.int32), np.int32), D.up(bvec), D.up(np.zeros(n))
    D.chk(D.dev.ug4b200_lu_apply(D.ctx, n, dlu, dpiv, dx, db))
    x = D.down(dx, n)
    ref = scipy.linalg.lu_solve((lu, piv), bvec)
    assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.linalg.norm(M @ x - bvec) <= 1e-10 * np.linalg.norm(bvec)


def test_coarse_cg_solves(D, orc):
    from ugcore_b200 import problems as pr
    prob = pr.Problem(dim=3, num_refs=2)
    A = prob.matrix()
    n = A.nrows
    rng = np.random.default_rng(8)
    b = rng.standard_normal(n); b[prob.dirichlet() != 0] = 0.0
    dA, db, dx, dw = D.matrix(A), D.up(b), D.up(np.zeros(n)), D.alloc(4 * n * 8)
    D.chk(D.dev.ug4b200_coarse_cg(D.ctx, dA, dx, db, dw, 500, 1e-30, 1e-14))
    x = D.down(dx, n)
    xref = orc.matrix(A).lu_solve(b)
    assert np.linalg.norm(x - xref) <= 1e-11 * np.linalg.norm(xref)


def test_device_convergence_state_matches_stdconvcheck(D):
    """Device-side StdConvCheck (finaliser of the norm reduction) vs convergence_check_impl.h:162-169."""
    from ugcore_b200 import capi
    st = D.alloc(C.sizeof(capi.ConvState))
    hist = D.alloc(8 * 16)
    D.chk(D.dev.ug4b200_conv_init(D.ctx, st, 3, 1e-12, 1e-3, hist, 16))
    v = D.up(np.array([3.0, 4.0]))
    fin = capi.Fin(capi.FIN_CONV_START, None, None, None, st)
    D.chk(D.dev.ug4b200_vec_dot_ds(D.ctx, 2, v, v, fin))
    fin.op = capi.FIN_CONV_UPDATE
    D.chk(D.dev.ug4b200_vec_scale(D.ctx, 2, v, 0.1))
    D.chk(D.dev.ug4b200_vec_dot_ds(D.ctx, 2, v, v, fin))
    s = capi.ConvState.from_buffer_copy(D.down(st, C.sizeof(capi.ConvState), np.uint8).tobytes())
    assert (s.step, s.done) == (1, 0) and abs(s.current_defect - 0.5) < 1e-15 and s.initial_defect == 5.0
    D.chk(D.dev.ug4b200_vec_scale(D.ctx, 2, v, 1e-3))
    D.chk(D.dev.ug4b200_vec_dot_ds(D.ctx, 2, v, v, fin))
    s = capi.ConvState.from_buffer_copy(D.down(st, C.sizeof(capi.ConvState), np.uint8).tobytes())
    assert (s.step, s.done, s.status) == (2, 1, 1)
    h = D.down(hist, 3)
    assert h[0] == 5.0 and abs(h[1] - 0.5) < 1e-15
    # guarded kernels become no-ops once done is set
    D.chk(D.dev.ug4b200_set_guard(D.ctx, C.c_void_p(st.value + capi.ConvState.done.offset)))
    D.chk(D.dev.ug4b200_vec_set(D.ctx, 2, v, 7.0))
    D.chk(D.dev.ug4b200_set_guard(D.ctx, None))
    assert not np.any(D.down(v, 2) == 7.0)


def test_value_indexed_stream_selection_and_parity(D, orc):
    """The value-indexed entry stream is built only when it is lossless and fits 16 bits twice:
    <= 65536 distinct values and, per 32-row slice, columns within 65535 of the smallest one.
    Whatever is chosen, y = A x stays bit-identical to the CPU."""
    from ugcore_b200 import capi, problems as pr
    plain_ctx = "UG4B200_NO_COMPRESS"  # the 'plain' parametrisation of the context never compresses
    rng = np.random.default_rng(5)

    def crs(n, ncols, rows):
        rowptr = np.zeros(n + 1, np.int64)
        cols, vals = [], []
        for r, (c, v) in enumerate(rows):
            order = np.argsort(c)
            cols.append(np.asarray(c, np.int32)[order]); vals.append(np.asarray(v, np.float64)[order])
            rowptr[r + 1] = rowptr[r] + len(c)
        return pr.Crs(n, ncols, 1, rowptr, np.concatenate(cols).astype(np.int32), np.concatenate(vals))

    n = 300
    band = [(np.unique(np.clip(np.array([r - 3, r - 1, r, r + 1, r + 5]), 0, n - 1)), None) for r in range(n)]
    few = crs(n, n, [(c, rng.choice([1.5, -0.25, 0.0, -0.0, 3.0], size=c.size)) for c, _ in band])
    allrandom = crs(n, n, [(c, rng.standard_normal(c.size)) for c, _ in band])
    wide_n = 70000
    wide = crs(64, wide_n, [(np.array([r, wide_n - 1 - r]), np.array([2.0, -1.0])) for r in range(64)])
    many_n = 70000
    many = crs(many_n, many_n, [(np.array([r]), np.array([float(r) + 0.5])) for r in range(many_n)])
    expect = {"few": True, "allrandom": True, "wide": False, "many": False}  # allrandom: 1500 distinct values <= 65536
    for name, A in (("few", few), ("allrandom", allrandom), ("wide", wide), ("many", many)):
        dA = D.matrix(A)
        info = capi.MatrixInfo()
        D.chk(D.dev.ug4b200_matrix_get_info(dA, C.byref(info)))
        if info.value_indexed:
            assert expect[name], name
            assert info.num_distinct_values <= 65536
        x = rng.standard_normal(A.ncols)
        y = D.up(np.zeros(A.nrows))
        D.chk(D.dev.ug4b200_matrix_apply(D.ctx, dA, y, D.up(x), 1))
        assert np.array_equal(D.down(y, A.nrows), orc.matrix(A).apply(x)), name
        D.chk(D.dev.ug4b200_matrix_destroy(D.ctx, dA))
