"""Import / export of assembled objects: ConnectionViewer .mat / .vec and
MatrixMarket .mtx — host-side code (csrc/host/matrix_io.h) through the C ABI, checked on the CPU

  * byte for byte against files written by the REFERENCE's own ConnectionViewer writer, and read
    back by the reference's own reader (oracle/_ref: connection_viewer_{output,input}.h compiled
    from /root/reference),
  * against the MatrixMarket format rules of matrix_io_mtx.{h,cpp} (banner, 1-based column-major
    entries, symmetry detection, "%.13e") and an independent reader (scipy.io.mmread),
  * on the edge cases the formats have: explicit zeros (Dirichlet rows), rectangular P in the
    from / to form, marker lines behind the connections, 1-/2-/3-d positions, empty rows.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from ugcore_b200 import io as ugio
from ugcore_b200 import problems as pr


def _crs_equal(A, B):
    return (A.nrows == B.nrows and A.ncols == B.ncols and np.array_equal(A.rowptr, B.rowptr)
            and np.array_equal(A.cols, B.cols) and np.array_equal(A.vals, B.vals))


def _positions(prob, lev):
    d2l = np.asarray(prob.dof_to_lex(lev))
    dims = prob.dims(lev)
    h = 1.0 / (dims[0] - 1)
    pos = np.zeros((d2l.size, 3))
    pos[:, 0] = (d2l % dims[0]) * h
    pos[:, 1] = ((d2l // dims[0]) % dims[1]) * h
    pos[:, 2] = (d2l // (dims[0] * dims[1])) * h
    return pos


@pytest.fixture(scope="module")
def ref():
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    lib = C.CDLL(os.path.join(os.path.dirname(oracle.__file__), "_ref", "liboracle_ref.so"))
    lib.oracle_ref_cv_read_matrix.restype = C.c_void_p
    lib.oracle_ref_cv_read_matrix.argtypes = [C.c_char_p]
    for f in (lib.oracle_ref_cv_read_info, lib.oracle_ref_cv_read_export, lib.oracle_ref_cv_read_free):
        f.restype = None
    lib.oracle_ref_cv_read_info.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.oracle_ref_cv_read_export.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.oracle_ref_cv_read_free.argtypes = [C.c_void_p]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _ref_write_matrix(lib, path, A, pos, dim, from_to=False):
    f = lib.oracle_ref_cv_write_matrix_from_to if from_to else lib.oracle_ref_cv_write_matrix
    rp, ci, va = (np.ascontiguousarray(A.rowptr, np.int64), np.ascontiguousarray(A.cols, np.int32),
                  np.ascontiguousarray(A.vals, np.float64))
    pos = np.ascontiguousarray(pos, np.float64)
    assert f(str(path).encode(), C.c_int64(A.nrows), C.c_int64(A.ncols), _p(rp), _p(ci), _p(va), _p(pos), dim) == 0


@pytest.mark.parametrize("dim,refs", [(2, 3), (3, 2)])
def test_connection_viewer_matrix_matches_reference_writer_bytes(tmp_path, ref, dim, refs):
    prob = pr.Problem(dim=dim, num_refs=refs)
    A = prob.matrix(refs)               # Dirichlet rows keep their pattern: explicit zeros are written as " 0"
    pos = _positions(prob, refs)
    ugio.write_matrix(tmp_path / "ours.mat", A, pos, dim=dim)
    _ref_write_matrix(ref, tmp_path / "ref.mat", A, pos, dim)
    assert (tmp_path / "ours.mat").read_bytes() == (tmp_path / "ref.mat").read_bytes()
    assert (A.vals == 0.0).any()


def test_connection_viewer_from_to_matches_reference_writer_bytes(tmp_path, ref):
    prob = pr.Problem(dim=3, num_refs=2)
    P = prob.prolongation(2)
    pos = np.vstack([_positions(prob, 2), _positions(prob, 1)])   # rows ("to") first, then columns ("from")
    ugio.write_matrix(tmp_path / "ours.mat", P, pos, dim=3, from_to=True)
    _ref_write_matrix(ref, tmp_path / "ref.mat", P, pos, 3, from_to=True)
    assert (tmp_path / "ours.mat").read_bytes() == (tmp_path / "ref.mat").read_bytes()
    # and back: columns are numbered behind the rows
    B, rpos, d = ugio.read_matrix(tmp_path / "ref.mat", keep_zeros=True, n_to=P.nrows)
    assert d == 3 and B.nrows == P.nrows and B.ncols == P.ncols
    assert np.array_equal(B.rowptr, P.rowptr) and np.array_equal(B.cols, P.cols)
    assert np.allclose(B.vals, P.vals, rtol=1e-5)   # 6 significant digits in a default .mat
    assert np.allclose(rpos, pos, atol=1e-6)


def test_connection_viewer_vector_matches_reference_writer_bytes(tmp_path, ref):
    prob = pr.Problem(dim=3, num_refs=2)
    pos = _positions(prob, 2)
    rng = np.random.default_rng(3)
    v = rng.standard_normal(pos.shape[0]) * 10.0 ** rng.integers(-20, 20, pos.shape[0])
    v[0], v[1] = 0.0, -0.0
    ugio.write_vector(tmp_path / "ours.vec", v, pos, dim=3)
    assert ref.oracle_ref_cv_write_vector(str(tmp_path / "ref.vec").encode(), C.c_int64(v.size), _p(v),
                                          _p(np.ascontiguousarray(pos)), 3) == 0
    assert (tmp_path / "ours.vec").read_bytes() == (tmp_path / "ref.vec").read_bytes()
    # the reference's reader and ours agree on the reference's file
    out = np.zeros(v.size)
    assert ref.oracle_ref_cv_read_vector(str(tmp_path / "ref.vec").encode(), C.c_int64(v.size), _p(out)) == 0
    w, rpos, d = ugio.read_vector(tmp_path / "ref.vec")
    assert d == 3 and np.array_equal(w, out)
    # 16 significant digits: not always bit-exact, but far below the 1e-10 parity tolerance
    assert np.allclose(w, v, rtol=2e-16 * 10, atol=0.0)


def test_connection_viewer_reader_equals_reference_reader(tmp_path, ref):
    prob = pr.Problem(dim=3, num_refs=2, problem=pr.CONVDIFF)
    A = prob.matrix(2)
    pos = _positions(prob, 2)
    _ref_write_matrix(ref, tmp_path / "ref.mat", A, pos, 3)
    with open(tmp_path / "ref.mat", "a") as f:          # marker block of the ConnectionViewer: must end the connection list
        f.write("c 1 0 0 1 5\nv 3\n")
    h = ref.oracle_ref_cv_read_matrix(str(tmp_path / "ref.mat").encode())
    assert h
    nr, nc, nnz, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    ref.oracle_ref_cv_read_info(h, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(d))
    rp, ci, va = np.zeros(nr.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
    rpos = np.zeros((nr.value, 3))
    ref.oracle_ref_cv_read_export(h, _p(rp), _p(ci), _p(va), _p(rpos))
    ref.oracle_ref_cv_read_free(h)
    B, bpos, bd = ugio.read_matrix(tmp_path / "ref.mat")          # reference behaviour: zeros are not inserted
    assert bd == d.value == 3
    assert B.nrows == nr.value and B.ncols == nc.value
    # The reference reader only stops at a marker when peek() sees it, but peek() sees the newline that
    # ends the previous line: it runs one extraction too far, `from` becomes 0 (failed extraction), `to`
    # and `value` keep the last line's contents, and a spurious connection (0, last column) = last value
    # is inserted (connection_viewer_input.h:90-103).  That entry is not in the file; everything else
    # must be identical.
    last = (tmp_path / "ref.mat").read_text().splitlines()[-3].split()
    spurious = (0, int(last[1]), float(last[2]))
    import scipy.sparse as sp
    Rm = sp.csr_matrix((va, ci, rp), shape=(nr.value, nc.value)).tolil()
    assert Rm[spurious[0], spurious[1]] == spurious[2] and A.to_scipy()[spurious[0], spurious[1]] == 0.0
    Rm[spurious[0], spurious[1]] = 0.0
    Rm = Rm.tocsr(); Rm.eliminate_zeros(); Rm.sort_indices()
    assert np.array_equal(B.rowptr, Rm.indptr) and np.array_equal(B.cols, Rm.indices) and np.array_equal(B.vals, Rm.data)
    assert np.array_equal(bpos, rpos)
    assert B.nnz < A.nnz                                          # the explicit zeros of the Dirichlet rows are gone ...
    K, _, _ = ugio.read_matrix(tmp_path / "ref.mat", keep_zeros=True)
    assert np.array_equal(K.rowptr, A.rowptr) and np.array_equal(K.cols, A.cols)   # ... unless asked for


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_connection_viewer_lossless_round_trip(tmp_path, dim):
    rng = np.random.default_rng(dim)
    n = 37
    dense = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.15)
    dense[5, :] = 0.0                                             # an empty row
    rows, cols = np.nonzero(dense)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64)
    A = pr.Crs(nrows=n, ncols=n, block=1, rowptr=rowptr, cols=cols.astype(np.int32), vals=dense[rows, cols].copy())
    pos = np.zeros((n, 3)); pos[:, :dim] = rng.random((n, dim))
    ugio.write_matrix(tmp_path / "a.mat", A, pos, dim=dim, precision=17)
    B, bpos, d = ugio.read_matrix(tmp_path / "a.mat")
    assert d == dim and _crs_equal(A, B) and np.array_equal(bpos, pos)
    v = rng.standard_normal(n)
    ugio.write_vector(tmp_path / "a.vec", v, pos, dim=dim, precision=17)
    w, _, _ = ugio.read_vector(tmp_path / "a.vec")
    assert np.array_equal(v, w)


def test_matrix_market_general_format_and_round_trip(tmp_path):
    prob = pr.Problem(dim=3, num_refs=2)
    A = prob.matrix(2)                       # Dirichlet rows without eliminated columns: structurally non-symmetric
    ugio.write_matrix(tmp_path / "a.mtx", A)
    lines = (tmp_path / "a.mtx").read_text().splitlines()
    assert lines[0] == "%%MatrixMarket matrix coordinate real general"
    assert lines[1] == "%Generated with ug4."
    nnz_nonzero = int(np.count_nonzero(A.vals))
    assert lines[2] == f"{A.nrows} {A.ncols} {nnz_nonzero}"           # zeros are not written (matrix_io_mtx.h:416)
    assert len(lines) == 3 + nnz_nonzero
    # entries: 1-based, column-major, "m n  v" / "m n -v" in scientific notation with 13 digits
    S = A.to_scipy().tocsc()
    S.eliminate_zeros(); S.sort_indices()
    k = 0
    for c in range(3):
        for p in range(S.indptr[c], S.indptr[c + 1]):
            val = S.data[p]
            assert lines[3 + k] == f"{S.indices[p] + 1} {c + 1}" + (" " if val < 0 else "  ") + f"{val:.13e}"
            k += 1
    import scipy.io
    M = scipy.io.mmread(str(tmp_path / "a.mtx")).tocsr()
    assert abs(M - A.to_scipy()).max() < 1e-12
    # lossless variant and our own reader
    ugio.write_matrix(tmp_path / "b.mtx", A, precision=16)
    B, pos, _ = ugio.read_matrix(tmp_path / "b.mtx")
    assert pos is None
    Z = A.to_scipy().tocsr(); Z.eliminate_zeros(); Z.sort_indices()
    assert np.array_equal(B.rowptr, Z.indptr) and np.array_equal(B.cols, Z.indices) and np.array_equal(B.vals, Z.data)


def test_matrix_market_symmetric_and_skew(tmp_path):
    import scipy.io
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n = 25
    L = sp.random(n, n, density=0.2, random_state=1, format="csr")
    for name, M in (("symmetric", (L + L.T + sp.eye(n)).tocsr()), ("skew-symmetric", (L - L.T).tocsr())):
        M.sort_indices()
        A = pr.Crs(nrows=n, ncols=n, block=1, rowptr=M.indptr.astype(np.int64), cols=M.indices.astype(np.int32), vals=M.data.copy())
        ugio.write_matrix(tmp_path / "s.mtx", A, precision=16)
        lines = (tmp_path / "s.mtx").read_text().splitlines()
        assert lines[0] == f"%%MatrixMarket matrix coordinate real {name}"
        lower = sp.tril(M).tocsr(); lower.eliminate_zeros()
        assert lines[2] == f"{n} {n} {lower.nnz}"                   # only the lower triangle is stored
        assert abs(scipy.io.mmread(str(tmp_path / "s.mtx")).tocsr() - M).max() == 0.0
        B, _, _ = ugio.read_matrix(tmp_path / "s.mtx")               # expanded again on input (matrix_io_mtx.h:248-252)
        Z = M.copy(); Z.eliminate_zeros(); Z.sort_indices()
        assert np.array_equal(B.rowptr, Z.indptr) and np.array_equal(B.cols, Z.indices) and np.array_equal(B.vals, Z.data)
    # files of other writers: comments, blank line before the size line, integer values
    (tmp_path / "t.mtx").write_text("%%MatrixMarket matrix coordinate integer general\n% a comment\n%\n3 4 3\n1 1 2\n3 4 -7\n2 2 5\n")
    T, _, _ = ugio.read_matrix(tmp_path / "t.mtx")
    assert T.nrows == 3 and T.ncols == 4 and T.to_scipy().toarray().tolist() == [[2, 0, 0, 0], [0, 5, 0, 0], [0, 0, 0, -7]]


def test_io_errors_are_loud(tmp_path):
    from ugcore_b200.capi import UG4B200Error
    with pytest.raises(UG4B200Error):
        ugio.read_matrix(tmp_path / "missing.mat")
    (tmp_path / "bad.mat").write_text("2\n3\n1\n0 0 0\n1\n")
    with pytest.raises(UG4B200Error):
        ugio.read_matrix(tmp_path / "bad.mat")                      # version 2 is not ConnectionViewer v1
    (tmp_path / "oob.mat").write_text("1\n2\n2\n0 0\n1 1\n1\n0 5 1.0\n")
    with pytest.raises(UG4B200Error):
        ugio.read_matrix(tmp_path / "oob.mat")                      # connection outside the matrix
    (tmp_path / "dense.mtx").write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(UG4B200Error):
        ugio.read_matrix(tmp_path / "dense.mtx")                    # "Other than sparse ... not yet implemented"


def test_solve_from_dumped_files(tmp_path):
    """The point of the importer: a hierarchy dumped to files and read back (lossless precision) is
    the same hierarchy — the oracle solves both to identical histories."""
    prob = pr.Problem(dim=2, num_refs=3)
    lv = {}
    for l in range(0, 4):
        pos = _positions(prob, l)
        ugio.write_matrix(tmp_path / f"A{l}.mat", prob.matrix(l), pos, dim=2, precision=17)
        A, _, _ = ugio.read_matrix(tmp_path / f"A{l}.mat", keep_zeros=True)
        assert _crs_equal(A, prob.matrix(l))
        P = R = None
        if l:
            ugio.write_matrix(tmp_path / f"P{l}.mtx", prob.prolongation(l), precision=16)
            ugio.write_matrix(tmp_path / f"R{l}.mtx", prob.restriction(l), precision=16)
            P, _, _ = ugio.read_matrix(tmp_path / f"P{l}.mtx")
            R, _, _ = ugio.read_matrix(tmp_path / f"R{l}.mtx")
        lv[l] = (A, P, R)
    ugio.write_vector(tmp_path / "b.vec", np.array(prob.rhs()), _positions(prob, 3), dim=2, precision=17)
    b, _, _ = ugio.read_vector(tmp_path / "b.vec")
    assert np.array_equal(b, np.array(prob.rhs()))
    from helpers import gmg_desc, oracle_levels
    orc = oracle.Oracle("port")
    desc = gmg_desc(3)
    lv_file = {l: (orc.matrix(A), orc.matrix(P) if P is not None else None, orc.matrix(R) if R is not None else None)
               for l, (A, P, R) in lv.items()}
    lv_gen = oracle_levels(orc, prob)
    x1, ok1, h1 = oracle.OSolver(orc, desc, lv_file[3][0], lv_file).apply(b)
    x2, ok2, h2 = oracle.OSolver(orc, desc, lv_gen[3][0], lv_gen).apply(np.array(prob.rhs()))
    assert ok1 and ok2 and np.array_equal(h1, h2) and np.array_equal(x1, x2)


@pytest.mark.gpu
def test_gpu_cg_debug_writer_leaves_the_reference_file_set(tmp_path):
    """solver:set_debug(writer): CG_Residual_iterNNN.vec / CG_Solution_iterNNN.vec after every step (cg.h:124, 195,
    273-280) — readable by the ConnectionViewer reader, residual norms = the defect history, last solution = the result;
    the history equals the run without a writer (which uses the device-resident loop)."""
    import ugcore_b200 as ug
    from ugcore_b200 import io as ugio, problems as pr
    from helpers import gmg_desc
    prob = pr.Problem(dim=3, num_refs=3)
    desc = gmg_desc(3)
    s = ug.Solver.from_problem(desc, prob)
    s.set_debug_dir(tmp_path, precision=17)
    x, ok, h = s.apply(prob.rhs())
    x0, ok0, h0 = ug.Solver.from_problem(desc, prob).apply(prob.rhs())
    assert ok and ok0 and len(h) == len(h0) and np.allclose(h, h0, rtol=1e-10)
    steps = len(h) - 1
    names = sorted(p.name for p in tmp_path.iterdir())
    # the write before the loop and the one of the first step both carry the step count 0 (the reference's numbering)
    assert names == sorted([f"CG_{k}_iter{i:03d}.vec" for k in ("Residual", "Solution") for i in range(steps)])
    for i in range(steps):
        r = ugio.read_vector(str(tmp_path / f"CG_Residual_iter{i:03d}.vec"))[0]
        assert abs(np.linalg.norm(r) - h[i + 1]) <= 1e-12 * h[0]
    xs = ugio.read_vector(str(tmp_path / f"CG_Solution_iter{steps - 1:03d}.vec"))[0]
    assert np.array_equal(xs, x)
