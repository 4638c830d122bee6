#!/usr/bin/env python
"""Kernel micro-benchmark: top-level SpMV-family kernels timed alone with CUDA events
(used for tuning experiments; bench.py carries the judged numbers)."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ugcore_b200 import capi, problems as pr
from ugcore_b200.solver import host_ctx, DeviceBuffer

def main():
    refs = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = capi.dev; ctx = host_ctx()
    prob = pr.Problem(dim=3, num_refs=refs, order=order)
    A = prob.matrix(refs); n, nnz = A.nrows, A.nnz
    m = C.c_void_p()
    capi.check(dev.ug4b200_matrix_upload_crs(ctx, 1, n, n, A.rowptr.ctypes.data_as(C.c_void_p), A.cols.ctypes.data_as(C.c_void_p), A.vals.ctypes.data_as(C.c_void_p), 0, C.byref(m)), ctx)
    info = capi.MatrixInfo(); dev.ug4b200_matrix_get_info(m, C.byref(info))
    rng = np.random.default_rng(0)
    sd, st, st2, sc, dinv, q = (DeviceBuffer.from_numpy(rng.standard_normal(n)) for _ in range(6))
    S = DeviceBuffer(16)
    capi.check(dev.ug4b200_jacobi_prepare(ctx, m, 0.66, 1, dinv.ptr), ctx)
    e0, e1 = C.c_void_p(), C.c_void_p()
    dev.ug4b200_event_create(ctx, C.byref(e0)); dev.ug4b200_event_create(ctx, C.byref(e1))
    def timeit(fn, reps=30):
        for _ in range(5): fn()
        dev.ug4b200_sync(ctx); dev.ug4b200_event_record(ctx, e0)
        for _ in range(reps): fn()
        dev.ug4b200_event_record(ctx, e1); dev.ug4b200_event_sync(ctx, e1)
        ms = C.c_float(); dev.ug4b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms)); return ms.value / reps
    fin = capi.Fin(capi.FIN_STORE, S.ptr, None, None, None)
    res = {}
    b_apply = 12 * nnz + 4 * (n + 1) + 16 * n
    cases = {
      "fused_addin_jac": (lambda: dev.ug4b200_jacobi_smooth_fused(ctx, m, dinv.ptr, sd.ptr, st.ptr, st2.ptr, sc.ptr, 3), 12*nnz + 4*(n+1) + 64*n),
      "fused_addin": (lambda: dev.ug4b200_jacobi_smooth_fused(ctx, m, dinv.ptr, sd.ptr, st.ptr, None, sc.ptr, 1), 12*nnz + 4*(n+1) + 48*n),
      "matmul_minus": (lambda: dev.ug4b200_matrix_matmul_minus(ctx, m, sd.ptr, st.ptr, 1), b_apply + 8*n),
      "apply": (lambda: dev.ug4b200_matrix_apply(ctx, m, sd.ptr, st.ptr, 1), b_apply),
      "apply_dot": (lambda: dev.ug4b200_matrix_apply_dot_ds(ctx, m, q.ptr, st.ptr, fin), b_apply),
      "restrict_like_skip_empty": (lambda: dev.ug4b200_matrix_apply_ignore_zero_rows(ctx, m, sd.ptr, 1.0, st.ptr, 1), b_apply),
      "cg_update": (lambda: dev.ug4b200_cg_update_ds(ctx, n, sd.ptr, st.ptr, st2.ptr, sc.ptr, S.ptr, fin), 48*n),
      "dot": (lambda: dev.ug4b200_vec_dot_ds(ctx, n, sd.ptr, st.ptr, fin), 16*n),
      "axpy": (lambda: dev.ug4b200_vec_scale_add2(ctx, n, sd.ptr, 1.0, sd.ptr, 0.5, st.ptr), 24*n),
      "copy": (lambda: dev.ug4b200_vec_copy(ctx, n, sd.ptr, st.ptr), 16*n),
    }
    for k, (fn, nbytes) in cases.items():
        ms = timeit(fn)
        res[k] = {"us": round(ms * 1e3, 2), "GBs": round(nbytes / ms / 1e6, 1)}
    print(json.dumps({"variant": os.path.basename(os.environ.get("UG4B200_LIBDIR", "default").rstrip("/")), "n": n, "order": order,
                      "stream": "xs" if info.x_staged else ("vi" if info.value_indexed else "plain"), **res}))

main()
