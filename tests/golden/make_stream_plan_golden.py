"""Regenerates tests/golden/stream_plan_hashes.json: SHA-256 of every array the host-side entry-stream builder
(ug4b200_host_stream_plan, ug4b200_host_value_indexed_stream — the code ug4b200_matrix_upload_crs uploads from) produces
for a fixed set of matrices.  The streams are fully specified (tests/test_stream_plan.py decodes them), the hashes pin the
remaining freedom — dictionary order = order of first occurrence, run grouping — across refactorings of the builder.

    python tests/golden/make_stream_plan_golden.py            # rewrite the fixture
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ugcore_b200 import capi, problems as pr  # noqa: E402


def random_crs(n, per_row, ndistinct, seed):
    """banded random matrix with exactly `ndistinct` distinct values (first occurrences in random order)"""
    rng = np.random.default_rng(seed)
    pool = rng.standard_normal(ndistinct)
    rows = []
    for i in range(n):
        c = np.unique(np.clip(i + rng.integers(-40, 41, per_row), 0, n - 1))
        rows.append(c)
    rowptr = np.concatenate([[0], np.cumsum([len(c) for c in rows])]).astype(np.int64)
    cols = np.concatenate(rows).astype(np.int32)
    idx = rng.integers(0, ndistinct, cols.size)
    idx[:ndistinct] = rng.permutation(ndistinct) if cols.size >= ndistinct else idx[:ndistinct]
    return pr.Crs(n, n, 1, rowptr, cols, pool[idx].copy())


def matrices():
    p3 = pr.Problem(dim=3, num_refs=4)
    cd = pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=1e-2)
    mc = pr.Problem(dim=3, num_refs=2, base=(3, 1, 2))
    hier = pr.Problem(dim=3, num_refs=4, order=pr.ORDER_HIER)
    out = {"poisson3d_17": p3.matrix(4), "poisson3d_9": p3.matrix(3), "poisson3d_P4": p3.prolongation(4), "poisson3d_R4": p3.restriction(4),
           "poisson2d_33": pr.Problem(dim=2, num_refs=5).matrix(), "convdiff3d_9": cd.matrix(), "multi_cell": mc.matrix(),
           "poisson3d_hier_17": hier.matrix(),
           "random_300_values": random_crs(700, 9, 300, 1), "random_2000_values": random_crs(900, 12, 2000, 2),
           "random_70000_values": random_crs(9000, 12, 70000, 3), "empty_rows": pr.Crs(5, 5, 1, np.array([0, 0, 2, 2, 3, 3], np.int64),
                                                                                  np.array([0, 4, 2], np.int32), np.array([1.0, -0.0, 0.0]))}
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe(A):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rp, ci, va = np.ascontiguousarray(A.rowptr, np.int64), np.ascontiguousarray(A.cols, np.int32), np.ascontiguousarray(A.vals)
    p = capi.StreamPlan()
    capi.check(capi.dev.ug4b200_host_stream_plan(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), C.byref(p), None, None, None, None))
    d = {f: int(getattr(p, f)) for f, _ in capi.StreamPlan._fields_}
    if p.num_distinct_values > 0:
        dic = np.zeros(p.num_distinct_values)
        xw = np.zeros(p.padded_nnz if p.x_staged else 0, np.uint32)
        hdr = np.zeros((p.num_slices if p.x_staged else 0, 4), np.int32)
        runs = np.zeros((p.num_slices if p.x_staged else 0, max(p.x_staged_runs, 1), 2), np.int32)
        capi.check(capi.dev.ug4b200_host_stream_plan(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), C.byref(p), vp(xw) if p.x_staged else None,
                                                     vp(hdr) if p.x_staged else None, vp(runs) if p.x_staged else None, vp(dic)))
        d["dict"] = sha(dic)
        if p.x_staged:
            d["xw"], d["hdr"], d["runs"] = sha(xw), sha(hdr), sha(runs)
    if hasattr(capi.dev, "ug4b200_host_value_indexed_stream") and p.value_indexed:
        words = np.zeros(p.padded_nnz, np.uint32)
        cb = np.zeros(p.num_slices, np.int32)
        vs = C.c_int(-1)
        capi.check(capi.dev.ug4b200_host_value_indexed_stream(A.nrows, A.ncols, vp(rp), vp(ci), vp(va), vp(words), vp(cb), C.byref(vs)))
        d["vi_words"], d["vi_colbase"], d["vi_vshift"] = sha(words), sha(cb), vs.value
    return d


if __name__ == "__main__":
    out = {k: describe(A) for k, A in matrices().items()}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "stream_plan_hashes.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, {a: b for a, b in v.items() if not isinstance(b, str)})
