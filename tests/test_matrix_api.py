"""Assembly-side API of the GPU matrix type (what DomainDiscretization, constraints and transfers call on
matrix_type before the solve): GPUSparseMatrix (csrc/host/gpu_sparsematrix.h) against the reference's own
SparseMatrix<double> (ugbase/lib_algebra/cpu_algebra/sparsematrix.h:116-343, compiled into oracle/_ref) on the same
scripts of operations — inserting access, +=, scale, clear_retain_structure, resize_and_keep_values, defragment,
set(double), set_as_transpose_of, set_as_copy_of, const access, set_matrix_row / add_matrix_row, is_isolated.
Host only; no device involved."""
import ctypes as C

import numpy as np
import pytest


def _product(ops):
    from ugcore_b200.capi import check_host, host
    ops = np.ascontiguousarray(ops, dtype=np.float64).reshape(-1, 4)
    h = C.c_void_p()
    check_host(host.ug4b200_host_matrix_script(ops.shape[0], ops.ctypes.data_as(C.c_void_p), C.byref(h)))
    try:
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check_host(host.ug4b200_io_matrix_info(h, C.byref(nr), C.byref(nc), C.byref(nnz), None, None))
        rp, ci, va = np.zeros(nr.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check_host(host.ug4b200_io_matrix_export(h, p(rp), p(ci), p(va), None))
        iso = np.zeros(max(nr.value, 1), np.uint8)
        check_host(host.ug4b200_host_matrix_isolated(h, p(iso)))
    finally:
        host.ug4b200_io_matrix_free(h)
    return (nr.value, nc.value), rp, ci, va, iso[:nr.value]


def _random_script(rng, n=12, m=9, nops=160):
    ops = [(0, n, m, 0)]
    rows, cols = n, m
    for _ in range(nops):
        k = rng.integers(0, 100)
        r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
        v = float(rng.integers(-8, 9)) / 4.0            # dyadic values: every sum is exact, zeros do occur
        if k < 40:
            ops.append((1, r, c, v))
        elif k < 65:
            ops.append((2, r, c, v))
        elif k < 70:
            ops.append((3, 0, 0, float(rng.choice([0.5, 2.0, -1.0]))))
        elif k < 73:
            ops.append((4, 0, 0, 0))
        elif k < 78:
            rows, cols = int(rng.integers(max(1, rows - 3), rows + 4)), int(rng.integers(max(1, cols - 3), cols + 4))
            ops.append((5, rows, cols, 0))
        elif k < 83:
            ops.append((6, 0, 0, 0))
        elif k < 85 and rows == cols:
            ops.append((7, 0, 0, v))
        elif k < 88:
            ops.append((8, 0, 0, float(rng.choice([1.0, -2.0]))))
            rows, cols = cols, rows
        elif k < 90:
            ops.append((9, 0, 0, float(rng.choice([1.0, 0.5]))))
        elif k < 94:
            ops.append((10, r, c, 0))
        else:
            nn = int(rng.integers(1, min(cols, 5) + 1))
            cc = rng.choice(cols, size=nn, replace=False)
            ops.append((11 if k < 97 else 13, r, nn, 0))
            ops += [(12, 0, int(x), float(rng.integers(-4, 5)) / 2.0) for x in cc]
    return np.array(ops, dtype=np.float64)


@pytest.mark.parametrize("seed", range(12))
def test_assembly_side_api_matches_the_reference_sparsematrix(seed, orc_ref):
    ops = _random_script(np.random.default_rng(seed))
    M, iso_ref = orc_ref.matrix_script(ops)
    rp_ref, ci_ref, va_ref = M.export()
    shape, rp, ci, va, iso = _product(ops)
    assert shape == (M.nrows, M.ncols)
    # (the reference's nnz counter is not lowered when resize_and_keep_values drops columns, sparsematrix_impl.h:134-135,
    #  so its total_num_connections() may exceed the stored entries: the rows themselves are what counts)
    n = int(rp_ref[-1])
    assert np.array_equal(rp, rp_ref) and np.array_equal(ci, ci_ref[:n]), "pattern (explicit zeros included)"
    assert np.array_equal(va, va_ref[:n])
    assert np.array_equal(iso, iso_ref)


def test_const_access_creates_nothing_and_bad_scripts_fail(orc_ref):
    from ugcore_b200.capi import host
    ops = np.array([(0, 3, 3, 0), (1, 0, 0, 2.0), (10, 2, 1, 0), (10, 0, 0, 0)], dtype=np.float64)
    shape, rp, ci, va, iso = _product(ops)
    assert shape == (3, 3) and list(rp) == [0, 1, 1, 1] and list(ci) == [0] and list(va) == [2.0] and list(iso) == [1, 1, 1]
    h = C.c_void_p()
    bad = np.array([(0, 2, 2, 0), (99, 0, 0, 0)], dtype=np.float64)
    assert host.ug4b200_host_matrix_script(2, bad.ctypes.data_as(C.c_void_p), C.byref(h)) != 0
    assert b"unknown operation" in host.ug4b200_host_last_error()


def _transposed_cases():
    from ugcore_b200 import problems as pr
    out = []
    for prob, l in ((pr.Problem(dim=3, num_refs=3), 3), (pr.Problem(dim=3, num_refs=2, problem=pr.ELASTICITY), 2),
                    (pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF), 3)):
        out += [(prob, prob.matrix(l)), (prob, prob.prolongation(l))]
    return out


def test_apply_transposed_scatter_equals_explicit_transpose(orc, orc_ref):
    """SparseMatrix::apply_transposed scatters row by row (sparsematrix_impl.h:341-370); the product multiplies by the
    explicit transpose instead.  Both sum the same terms in the same order: identical bits — shown here with the
    reference's own two routines (and the port's restatement of the scatter loop)."""
    for _, A in _transposed_cases():
        x = np.random.default_rng(0).standard_normal(A.nrows * A.block)
        y = orc_ref.matrix(A).apply_transposed(x)
        assert np.array_equal(y, orc.matrix(A).apply_transposed(x))
        assert np.array_equal(y, orc_ref.matrix(A).transpose().apply(x))


@pytest.mark.gpu
def test_gpu_apply_transposed_is_bit_identical(orc):
    from ugcore_b200.capi import check_host, host
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for _, A in _transposed_cases():
        x = np.random.default_rng(1).standard_normal(A.nrows * A.block)
        y = np.zeros(A.ncols * A.block)
        check_host(host.ug4b200_host_apply_transposed(A.block, A.nrows, A.ncols, p(np.ascontiguousarray(A.rowptr)),
                                                      p(np.ascontiguousarray(A.cols)), p(np.ascontiguousarray(A.vals)), p(y), p(x)))
        assert np.array_equal(y, orc.matrix(A).apply_transposed(x))


def test_oracle_set_random_and_maxnorm(orc, orc_ref):
    """Vector::set_random / maxnorm (vector_impl.h:91-96, 332-338): the port restates urand over the C library's rand()."""
    for block in (1, 3):
        a, b = orc_ref.set_random(40, block, 7, -1.0, 2.0), orc.set_random(40, block, 7, -1.0, 2.0)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] == np.abs(a[0]).max()
        assert a[0].min() >= -1.0 and a[0].max() < 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("block", [1, 3])
def test_gpu_vector_assembly_side_api(block, orc):
    """GPUVector::set_random (same numbers as ugcore's Vector for the same seed), add / get through index lists on the
    host mirror around a device operation, maxnorm."""
    from ugcore_b200.capi import check_host, host
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n, seed = 200, 11
    idx = np.array([3, 17, 3, 199], np.int64)                       # a repeated index accumulates twice
    addv = np.arange(1, idx.size * block + 1, dtype=np.float64)
    vals, got, mx = np.zeros(n * block), np.zeros(idx.size * block), np.zeros(1)
    check_host(host.ug4b200_host_vector_selftest(block, n, seed, -1.0, 1.0, idx.size, p(idx), p(addv), p(vals), p(got), p(mx)))
    ref, _ = orc.set_random(n, block, seed, -1.0, 1.0)
    assert np.array_equal(vals, ref)
    want = ref.reshape(n, block).copy()
    for k, i in enumerate(idx):
        want[i] += addv.reshape(-1, block)[k]
    want *= 2.0
    assert np.array_equal(got.reshape(-1, block), want[idx])
    assert mx[0] == np.abs(want).max()
