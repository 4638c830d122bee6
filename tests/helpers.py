"""Shared helpers of the test-suite (oracle-side hierarchy wiring, device buffers)."""
import ctypes as C

import numpy as np


def oracle_levels(orc, prob, base=None, top=None):
    base = prob.base_lev if base is None else base
    top = prob.num_refs if top is None else top
    lv = {}
    for l in range(base, top + 1):
        A = orc.matrix(prob.matrix(l))
        P = R = None
        if l > base:
            P = orc.matrix(prob.prolongation(l))
            R = orc.matrix(prob.restriction(l))
        lv[l] = (A, P, R)
    return lv


def gmg_desc(top, solver="cg", smoother=None, nu=(2, 2), cycle="V", base=0, base_solver="lu",
             its=100, absolute=1e-12, reduction=1e-10):
    smoother = smoother or {"type": "jac", "damp": 0.66}
    return {"type": solver,
            "precond": {"type": "gmg", "topLevel": top, "baseLevel": base, "smoother": smoother, "cycle": cycle,
                        "preSmooth": nu[0], "postSmooth": nu[1], "baseSolver": base_solver},
            "convCheck": {"iterations": its, "absolute": absolute, "reduction": reduction}}


def rel_hist_err(h_gpu, h_ref):
    """max_k |h_gpu[k] - h_ref[k]| / |h_ref[k]| over the common prefix.

    An absolute allowance of 1e-14 * (start defect) is subtracted first: a defect norm cannot
    be known more accurately than fp64 round-off of the start defect (eps * ||b||), whatever
    the summation order of the dot products — on the CPU as well."""
    n = min(len(h_gpu), len(h_ref))
    if not n:
        return 0.0
    diff = np.maximum(np.abs(h_gpu[:n] - h_ref[:n]) - 1e-14 * abs(h_ref[0]), 0.0)
    return float(np.max(diff / np.abs(h_ref[:n])))


class Dev:
    """Raw device arrays through the kernel-level C ABI (no torch)."""

    def __init__(self, ctx):
        from ugcore_b200 import capi
        self.capi, self.dev, self.ctx = capi, capi.dev, ctx
        self._bufs = []

    def chk(self, rc):
        return self.capi.check(rc, self.ctx)

    def alloc(self, nbytes):
        p = C.c_void_p()
        self.chk(self.dev.ug4b200_alloc(self.ctx, max(int(nbytes), 8), C.byref(p)))
        self._bufs.append(p)
        return p

    def up(self, a, dtype=np.float64):
        a = np.ascontiguousarray(a, dtype=dtype)
        p = self.alloc(a.nbytes)
        self.chk(self.dev.ug4b200_h2d(self.ctx, p, a.ctypes.data_as(C.c_void_p), a.nbytes))
        self.chk(self.dev.ug4b200_sync(self.ctx))
        return p

    def down(self, p, n, dtype=np.float64):
        a = np.empty(n, dtype=dtype)
        self.chk(self.dev.ug4b200_d2h(self.ctx, a.ctypes.data_as(C.c_void_p), p, a.nbytes))
        return a

    def matrix(self, crs):
        m = C.c_void_p()
        self.chk(self.dev.ug4b200_matrix_upload_crs(
            self.ctx, crs.block, crs.nrows, crs.ncols, crs.rowptr.ctypes.data_as(C.c_void_p),
            crs.cols.ctypes.data_as(C.c_void_p), crs.vals.ctypes.data_as(C.c_void_p), 0, C.byref(m)))
        return m

    def free_all(self):
        for p in self._bufs:
            self.dev.ug4b200_free(self.ctx, p)
        self._bufs = []


def make_rhs(prob, seed=None):
    """Top-level right-hand side: the generator's (seed None) or a seeded random one that
    respects the Dirichlet rows (defect must vanish there)."""
    if seed is None:
        return np.array(prob.rhs())
    rng = np.random.default_rng(seed)
    b = rng.standard_normal(prob.num_dofs)
    mask = np.repeat(prob.dirichlet() != 0, prob.block)
    b[mask] = 0.0
    return b
