"""CPU: the entry-stream formats of ug4b200_matrix_upload_crs, checked on the host through ug4b200_host_stream_plan (the
same builder code, no device): which 4-byte stream a matrix gets, and that the x-staged stream is a lossless re-encoding —
every word leads back to its column and its value, runs are 16-byte aligned, even-sized and inside the vector (up to its
16-byte padding), the header carries the bytes the kernel will wait for."""
import ctypes as C

import numpy as np
import pytest

from ugcore_b200 import capi, problems as pr
from ugcore_b200.solver import cuthill_mckee, permute_crs


def _plan(A, arrays=False):
    p = capi.StreamPlan()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rp, ci, va = np.ascontiguousarray(A.rowptr, np.int64), np.ascontiguousarray(A.cols, np.int32), np.ascontiguousarray(A.vals)
    capi.check(capi.dev.ug4b200_host_stream_plan(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), C.byref(p), None, None, None, None))
    if not arrays or not p.x_staged:
        return p, None
    xw = np.zeros(p.padded_nnz, np.uint32)
    hdr = np.zeros((p.num_slices, 4), np.int32)
    runs = np.zeros((p.num_slices, p.x_staged_runs, 2), np.int32)
    dic = np.zeros(p.num_distinct_values)
    capi.check(capi.dev.ug4b200_host_stream_plan(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), C.byref(p), vp(xw), vp(hdr), vp(runs), vp(dic)))
    return p, (xw, hdr, runs, dic)


def _check_lossless(A, p, arrs):
    xw, hdr, runs, dic = arrs
    n = A.nrows
    for s in range(p.num_slices):
        base, width, xbytes, nr = (int(v) for v in hdr[s])
        start = runs[s, :nr, 0].astype(np.int64)
        length = (runs[s, :nr, 1] & 0xffff).astype(np.int64)
        pos = (runs[s, :nr, 1] >> 16).astype(np.int64)
        assert np.all(start % 2 == 0) and np.all(length % 2 == 0) and np.all(length > 0)
        assert np.array_equal(pos, np.concatenate([[0], np.cumsum(length)[:-1]])) and int(length.sum()) * 8 == xbytes
        assert np.all(start + length <= A.ncols + (A.ncols & 1))          # at most the 16-byte padding behind the vector
        assert np.all(runs[s, nr:] == 0)
        stage = np.concatenate([np.arange(a, a + l) for a, l in zip(start, length)]) if nr else np.zeros(0, np.int64)
        for l in range(32):
            r = s * 32 + l
            if r >= n:
                break
            k = np.arange(A.rowptr[r + 1] - A.rowptr[r])
            assert k.size <= width
            w = xw[base * 32 + k * 32 + l]
            assert np.array_equal(stage[w >> 16], A.cols[A.rowptr[r]:A.rowptr[r + 1]])
            got = dic[(w & 0xffff) >> 3]
            assert np.array_equal(got.view(np.int64), np.asarray(A.vals[A.rowptr[r]:A.rowptr[r + 1]]).view(np.int64))
            assert np.all((w & 7) == 0)


@pytest.mark.parametrize("name", ["poisson3d_lex", "poisson2d", "convdiff3d", "multi_cell", "prolongation", "restriction"])
def test_x_staged_stream_is_a_lossless_reencoding(name):
    M = {"poisson3d_lex": lambda: pr.Problem(dim=3, num_refs=3).matrix(),
         "poisson2d": lambda: pr.Problem(dim=2, num_refs=5).matrix(),
         "convdiff3d": lambda: pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2).matrix(),
         "multi_cell": lambda: pr.Problem(dim=3, num_refs=2, base=(3, 1, 2)).matrix(),
         "prolongation": lambda: pr.Problem(dim=3, num_refs=3).prolongation(3),
         "restriction": lambda: pr.Problem(dim=3, num_refs=3).restriction(3)}[name]()
    p, arrs = _plan(M, arrays=True)
    assert p.value_indexed and p.x_staged, (name, p.num_distinct_values, p.x_staged_runs, p.x_staged_max_doubles)
    assert p.x_staged_runs <= 32 and p.x_staged_max_doubles <= 384
    _check_lossless(M, p, arrs)


def test_which_stream_a_numbering_gets():
    """The facts DESIGN.md §2 / §6 rest on: lexicographic 27-point operator = 9 runs per slice; 65^3 fits the 16-bit column
    window, 257^3 does not (two planes + a slice = 132 645 columns) but is x-stageable; ugcore's hierarchical numbering is neither;
    Cuthill-McKee of it is x-stageable with ~23 runs."""
    lex = pr.Problem(dim=3, num_refs=5).matrix()
    p, _ = _plan(lex)
    assert p.value_indexed and p.x_staged and p.x_staged_runs == 9 and p.max_column_window == 2 * 33 * 33 + 2 * 33 + 2 + 31
    assert 2 * 257 * 257 + 2 * 257 + 2 + 31 > 65535 > 2 * 129 * 129 + 2 * 129 + 2 + 31   # (two planes + 32 rows) 257^3: no; 129^3: fits
    hier = pr.Problem(dim=3, num_refs=5, order=pr.ORDER_HIER).matrix()
    p, _ = _plan(hier)
    assert not p.x_staged and p.num_distinct_values == 6
    perm = cuthill_mckee(hier, reverse=False)
    cmk = permute_crs(hier, perm, perm)
    p, arrs = _plan(cmk, arrays=True)
    assert p.x_staged and 9 < p.x_staged_runs <= 32 and p.x_staged_max_doubles <= 384
    _check_lossless(cmk, p, arrs)


def test_plain_stream_cases():
    """More than 256 distinct values: no x-staged stream; rows longer than 27 entries: neither; more than 65536 values:
    no dictionary at all."""
    from ugcore_b200.problems import Crs
    rng = np.random.default_rng(1)
    n = 400
    rp = np.arange(0, 3 * n + 1, 3, dtype=np.int64)
    ci = np.stack([np.maximum(np.arange(n) - 1, 0), np.arange(n), np.minimum(np.arange(n) + 1, n - 1)], 1)
    ci = np.sort(ci, 1)
    ci[0] = [0, 1, 2]; ci[-1] = [n - 3, n - 2, n - 1]
    A = Crs(n, n, 1, rp, ci.ravel().astype(np.int32), rng.standard_normal(3 * n))
    p, _ = _plan(A)
    assert p.num_distinct_values == 3 * n and p.value_indexed and not p.x_staged          # 1200 values: 16-bit dictionary only
    B = Crs(1, 64, 1, np.array([0, 40], np.int64), np.arange(40, dtype=np.int32), np.ones(40))
    p, _ = _plan(B)
    assert p.max_row_len == 40 and p.value_indexed and not p.x_staged
    m = 70000
    D = Crs(m, m, 1, np.arange(m + 1, dtype=np.int64), np.arange(m, dtype=np.int32), rng.standard_normal(m))
    p, _ = _plan(D)
    assert p.num_distinct_values == -1 and not p.value_indexed and not p.x_staged


# ---- value-indexed stream and the value dictionary (host builder shared with ug4b200_matrix_upload_crs) ----

def _vi(A):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rp, ci, va = np.ascontiguousarray(A.rowptr, np.int64), np.ascontiguousarray(A.cols, np.int32), np.ascontiguousarray(A.vals)
    p, _ = _plan(A)
    words = np.zeros(max(p.padded_nnz, 1), np.uint32)
    cb = np.zeros(max(p.num_slices, 1), np.int32)
    vs = C.c_int(-7)
    capi.check(capi.dev.ug4b200_host_value_indexed_stream(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), vp(words), vp(cb), C.byref(vs)))
    dic = np.zeros(max(p.num_distinct_values, 1))
    if p.num_distinct_values > 0:
        capi.check(capi.dev.ug4b200_host_stream_plan(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), C.byref(p), None, None, None, vp(dic)))
    return p, words[:p.padded_nnz], cb[:p.num_slices], vs.value, dic


def _first_occurrence_order(vals):
    bits = np.asarray(vals).view(np.int64)
    _, first = np.unique(bits, return_index=True)
    return np.asarray(vals)[np.sort(first)]


@pytest.mark.parametrize("name", ["poisson3d", "convdiff3d", "prolongation", "random_300_values", "random_2000_values", "empty_rows"])
def test_value_indexed_stream_is_a_lossless_reencoding(name):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_stream_plan_golden", os.path.join(os.path.dirname(__file__), "golden", "make_stream_plan_golden.py"))
    gold = importlib.util.module_from_spec(spec); spec.loader.exec_module(gold)
    M = {"poisson3d": lambda: pr.Problem(dim=3, num_refs=3).matrix(),
         "convdiff3d": lambda: pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2).matrix(),
         "prolongation": lambda: pr.Problem(dim=3, num_refs=3).prolongation(3),
         "random_300_values": lambda: gold.random_crs(700, 9, 300, 1),
         "random_2000_values": lambda: gold.random_crs(900, 12, 2000, 2),
         "empty_rows": lambda: gold.matrices()["empty_rows"]}[name]()
    p, words, cb, vshift, dic = _vi(M)
    assert p.value_indexed and vshift == (3 if p.num_distinct_values <= 1024 else 0)
    # the dictionary lists the values in the order a sequential scan of the CRS array first meets them (bit patterns:
    # 0.0 and -0.0 are two entries) — the numbering the parallel scan has to reproduce
    assert np.array_equal(dic.view(np.int64), _first_occurrence_order(M.vals).view(np.int64))
    rowlen = np.diff(M.rowptr)
    seen = np.zeros(p.padded_nnz, bool)
    off = 0
    for s in range(p.num_slices):
        rows = np.arange(s * 32, min(s * 32 + 32, M.nrows))
        width = int(rowlen[rows].max()) if rows.size else 0
        lo = min((int(M.cols[M.rowptr[r]]) for r in rows if rowlen[r]), default=0)
        assert cb[s] == lo
        for r in rows:
            k = np.arange(rowlen[r])
            pos = off + k * 32 + (r - s * 32)
            w = words[pos]
            seen[pos] = True
            assert np.array_equal((w >> 16).astype(np.int64) + lo, M.cols[M.rowptr[r]:M.rowptr[r + 1]])
            got = dic[(w & 0xffff) >> vshift]
            assert np.array_equal(got.view(np.int64), np.asarray(M.vals[M.rowptr[r]:M.rowptr[r + 1]]).view(np.int64))
            if vshift:
                assert np.all((w & 7) == 0)
        off += width * 32
    assert off == p.padded_nnz and np.all(words[~seen] == 0)          # padding: column base, dictionary entry 0


def test_no_value_indexed_stream_beyond_the_limits():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_stream_plan_golden", os.path.join(os.path.dirname(__file__), "golden", "make_stream_plan_golden.py"))
    gold = importlib.util.module_from_spec(spec); spec.loader.exec_module(gold)
    p, words, cb, vshift, dic = _vi(gold.random_crs(9000, 12, 70000, 3))       # > 65536 distinct values
    assert p.num_distinct_values == -1 and not p.value_indexed and vshift == -1 and not words.any()
    n = 70000                                                                   # one slice spans columns 0 .. 69999
    wide = pr.Crs(n, n, 1, (np.arange(n + 1) * 2).astype(np.int64),
                  np.stack([np.zeros(n, np.int32), np.full(n, n - 1, np.int32)], 1).ravel(), np.ones(2 * n))
    p, words, cb, vshift, dic = _vi(wide)
    assert p.max_column_window == n - 1 and not p.value_indexed and vshift == -1 and p.num_distinct_values == 1


def test_large_matrix_dictionary_matches_the_sequential_scan():
    """> 65536 entries: the scan runs in parallel chunks; values first met in late chunks must still be numbered in
    global first-occurrence order."""
    rng = np.random.default_rng(5)
    n, per = 40000, 6
    cols = (np.arange(n)[:, None] + np.arange(per)[None, :]) % n
    cols.sort(axis=1)
    pool = rng.standard_normal(5000)
    idx = np.minimum(rng.integers(0, 5000, n * per), np.arange(n * per) // 40)   # new values keep appearing along the array
    A = pr.Crs(n, n, 1, (np.arange(n + 1) * per).astype(np.int64), cols.ravel().astype(np.int32), pool[idx].copy())
    p, words, cb, vshift, dic = _vi(A)
    assert p.num_distinct_values == np.unique(idx).size and vshift == 0
    assert np.array_equal(dic.view(np.int64), _first_occurrence_order(A.vals).view(np.int64))


def test_stream_builder_reproduces_the_committed_hashes():
    """tests/golden/stream_plan_hashes.json was written by the builder that the GPU suite validated (r02i); any later
    change of the host-side builder has to reproduce every array bit for bit."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location("make_stream_plan_golden", os.path.join(here, "golden", "make_stream_plan_golden.py"))
    gold = importlib.util.module_from_spec(spec); spec.loader.exec_module(gold)
    with open(os.path.join(here, "golden", "stream_plan_hashes.json")) as f:
        want = json.load(f)
    got = {k: gold.describe(A) for k, A in gold.matrices().items()}
    assert got == want
