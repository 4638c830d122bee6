#!/usr/bin/env python
"""BASELINE configs[0] stand-in (S1): 2-D Poisson, unit square, quads, numRefs = 7 (129^2 = 16 641 DoF), GMG V(2,2) Jacobi(0.66)
+ CG, StdConvCheck(100, 1e-12, 1e-10) — device-resident solve time on one GPU next to one serial CPU solve with the reference's
kernels, and the history comparison.  (bench.py carries the judged 3-D numbers; this fills the S1 row of BASELINE.md.)"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import ugcore_b200 as ug
from ugcore_b200 import capi, problems as pr
from ugcore_b200.solver import host_ctx, DeviceBuffer
from helpers import gmg_desc, oracle_levels

refs = 7
prob = pr.Problem(dim=2, num_refs=refs)
desc = gmg_desc(refs)
s = ug.Solver.from_problem(desc, prob).init()
b = np.array(prob.rhs())
bd, xd = DeviceBuffer.from_numpy(b), DeviceBuffer(b.size)
dev, ctx = capi.dev, host_ctx()
e0, e1 = C.c_void_p(), C.c_void_p()
dev.ug4b200_event_create(ctx, C.byref(e0)); dev.ug4b200_event_create(ctx, C.byref(e1))
def solve():
    dev.ug4b200_vec_set(ctx, b.size, xd.ptr, C.c_double(0.0))
    assert s.apply_device(xd, bd)
for _ in range(3): solve()
dev.ug4b200_sync(ctx); dev.ug4b200_event_record(ctx, e0)
K = 20
for _ in range(K): solve()
dev.ug4b200_event_record(ctx, e1); dev.ug4b200_event_sync(ctx, e1)
ms = C.c_float(); dev.ug4b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms))
h = s.history()
orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
lv = oracle_levels(orc, prob, 0, refs)
osol = oracle.OSolver(orc, desc, lv[refs][0], lv)
osol.apply(b)
t0 = time.perf_counter(); xo, oko, ho = osol.apply(b); tc = time.perf_counter() - t0
k = min(len(h), len(ho))
print(json.dumps({"config": "S1 2-D Poisson 129^2 GMG V(2,2) Jacobi + CG", "n": int(b.size), "gpu_ms_per_solve": ms.value / K,
                  "gpu_mdof_per_s": b.size / (ms.value / K * 1e-3) / 1e6, "iterations": len(h) - 1, "iterations_cpu": len(ho) - 1,
                  "cpu_solve_s": tc, "cpu_cores": 1, "cpu_kind": orc.kind,
                  "history_rel_err_vs_cpu": float(np.max(np.abs(h[:k] - ho[:k]) / np.abs(ho[:k])))}))
