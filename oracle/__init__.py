"""CPU oracle bindings — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper around ``oracle/liboracle.so`` (backend "port": plain C++ restatement
of ugcore's arithmetic) and ``oracle/_ref/liboracle_ref.so`` (backend "ref": the real
ugcore templates compiled from /root/reference/ugbase, control flow restated).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package.  ``ugcore_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "liboracle_ref.so")

SOLVER = {"cg": 0, "bicgstab": 1, "linear": 2, "lu": 3, "gmres": 5}
PRECOND = {None: 0, "none": 0, "jac": 1, "jacobi": 1, "gs": 2, "bgs": 3, "sgs": 4, "gmg": 5, "ilu": 6}


class SolverDesc(C.Structure):
    _fields_ = [
        ("solver", C.c_int), ("precond", C.c_int), ("damp", C.c_double),
        ("max_steps", C.c_int), ("min_defect", C.c_double), ("rel_reduction", C.c_double),
        ("base_lev", C.c_int), ("top_lev", C.c_int), ("cycle", C.c_int),
        ("nu1", C.c_int), ("nu2", C.c_int), ("smoother", C.c_int), ("smoother_damp", C.c_double),
        ("base_solver", C.c_int), ("base_max_steps", C.c_int),
        ("base_min_defect", C.c_double), ("base_rel_reduction", C.c_double),
        ("restart", C.c_int), ("ilu_beta", C.c_double),
    ]


def build(verbose: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0 or verbose:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One loaded oracle library (``kind`` = "port" or "ref")."""

    def __init__(self, kind: str = "port"):
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            if kind == "port":
                build()
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        # RTLD_LOCAL: both builds export identical symbols and must not interpose
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.kind = kind
        L.oracle_backend_name.restype = C.c_char_p
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_mat_create.restype = C.c_void_p
        L.oracle_mat_create.argtypes = [C.c_int, C.c_int64, C.c_int64, _lp, _ip, _dp]
        L.oracle_mat_destroy.argtypes = [C.c_void_p]
        for f in ("oracle_mat_nnz", "oracle_mat_rows", "oracle_mat_cols"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_mat_export.argtypes = [C.c_void_p, _lp, _ip, _dp]
        L.oracle_mat_transpose.restype = C.c_void_p
        L.oracle_mat_transpose.argtypes = [C.c_void_p, C.c_int]
        L.oracle_mat_rap.restype = C.c_void_p
        L.oracle_mat_rap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_axpy.argtypes = [C.c_void_p, _dp, C.c_double, C.c_void_p, C.c_double, _dp, C.c_int]
        L.oracle_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.oracle_apply_transposed.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_vec_set_random.argtypes = [C.c_int64, C.c_int, C.c_uint, C.c_double, C.c_double, _dp, _dp]
        L.oracle_matmul_minus.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.oracle_apply_ignore_zero_rows.argtypes = [C.c_void_p, _dp, C.c_double, _dp, C.c_int]
        L.oracle_dot.restype = C.c_double
        L.oracle_dot.argtypes = [C.c_int64, C.c_int, _dp, _dp]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [C.c_int64, C.c_int, _dp]
        L.oracle_scale_add2.argtypes = [C.c_int64, _dp, C.c_double, _dp, C.c_double, _dp]
        L.oracle_scale_add3.argtypes = [C.c_int64, _dp, C.c_double, _dp, C.c_double, _dp, C.c_double, _dp]
        L.oracle_jacobi.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, _dp]
        L.oracle_gs.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _dp]
        L.oracle_lu_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_mat_script.restype = C.c_void_p
        L.oracle_mat_script.argtypes = [C.c_int64, _dp, C.c_void_p, C.c_int64]
        L.oracle_ilu_factorize.restype = C.c_void_p
        L.oracle_ilu_factorize.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_ilu_apply.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        L.oracle_cuthill_mckee.argtypes = [C.c_void_p, C.c_int, C.c_int, _lp]
        L.oracle_solver_create.restype = C.c_void_p
        L.oracle_solver_create.argtypes = [C.POINTER(SolverDesc)]
        L.oracle_solver_destroy.argtypes = [C.c_void_p]
        L.oracle_solver_set_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_solver_set_smoother_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_solver_set_precond_matrix.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_solver_set_surface_map.argtypes = [C.c_void_p, C.c_int64, _ip]
        L.oracle_solver_init.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_solver_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.oracle_solver_steps.argtypes = [C.c_void_p]
        L.oracle_solver_history.argtypes = [C.c_void_p, _dp, C.c_int]
        L.oracle_precond_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.oracle_set_reduction_mode.argtypes = [C.c_int]
        assert L.oracle_backend_name().decode() == kind

    def set_reduction_mode(self, mode: int):
        """0: ugcore's sequential dot / norm (the reference, default); 1: pairwise tree over the same products.
        Mode 1 only serves to measure the sensitivity of a residual history to the summation order."""
        self.lib.oracle_set_reduction_mode(int(mode))

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())
        return rc

    def matrix(self, crs) -> "OMat":
        """crs: object with nrows, ncols, block, rowptr(int64), cols(int32), vals(float64)."""
        return OMat(self, crs.block, crs.nrows, crs.ncols, crs.rowptr, crs.cols, crs.vals)

    def matrix_script(self, ops, max_rows=4096):
        """Run an op script against the reference's SparseMatrix (ref backend). Returns (OMat, isolated[nrows])."""
        ops = np.ascontiguousarray(ops, dtype=np.float64).reshape(-1, 4)
        iso = np.zeros(max_rows, np.uint8)
        h = self.lib.oracle_mat_script(ops.shape[0], ops.ravel(), iso.ctypes.data_as(C.c_void_p), max_rows)
        if not h:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())
        m = OMat(self, 1, self.lib.oracle_mat_rows(h), self.lib.oracle_mat_cols(h), None, None, None, handle=h)
        return m, iso[:m.nrows].copy()

    def set_random(self, nblocks, block, seed, lo, hi):
        """(values, maxnorm) of Vector::set_random(lo, hi) after srand(seed)."""
        out, mx = np.zeros(nblocks * block), np.zeros(1)
        self._chk(self.lib.oracle_vec_set_random(nblocks, block, seed, lo, hi, out, mx))
        return out, float(mx[0])

    def dot(self, a, b, block=1):
        a = _vec(a); b = _vec(b)
        return self.lib.oracle_dot(a.size // block, block, a, b)

    def norm(self, a, block=1):
        a = _vec(a)
        return self.lib.oracle_norm(a.size // block, block, a)

    def scale_add2(self, a1, v1, a2, v2):
        d = np.empty_like(_vec(v1))
        self._chk(self.lib.oracle_scale_add2(d.size, d, a1, _vec(v1), a2, _vec(v2)))
        return d

    def scale_add3(self, a1, v1, a2, v2, a3, v3):
        d = np.empty_like(_vec(v1))
        self._chk(self.lib.oracle_scale_add3(d.size, d, a1, _vec(v1), a2, _vec(v2), a3, _vec(v3)))
        return d


class OMat:
    def __init__(self, orc: Oracle, block, nrows, ncols, rowptr, cols, vals, handle=None):
        self.o = orc
        self.block, self.nrows, self.ncols = int(block), int(nrows), int(ncols)
        if handle is None:
            handle = orc.lib.oracle_mat_create(
                self.block, self.nrows, self.ncols,
                np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int32), _vec(vals))
            if not handle:
                raise RuntimeError("oracle: " + orc.lib.oracle_last_error().decode())
        self.h = handle

    def __del__(self):
        try:
            self.o.lib.oracle_mat_destroy(self.h)
        except Exception:
            pass

    @property
    def nnz(self):
        return self.o.lib.oracle_mat_nnz(self.h)

    def transpose(self, keep_zeros=True) -> "OMat":
        h = self.o.lib.oracle_mat_transpose(self.h, 1 if keep_zeros else 0)
        if not h:
            raise RuntimeError("oracle: " + self.o.lib.oracle_last_error().decode())
        return OMat(self.o, self.block, self.ncols, self.nrows, None, None, None, handle=h)

    def rap(self, R: "OMat", P: "OMat") -> "OMat":
        """R * self * P (AddMultiplyOf): the Galerkin coarse operator of gmg:set_rap(true)."""
        h = self.o.lib.oracle_mat_rap(R.h, self.h, P.h)
        if not h:
            raise RuntimeError("oracle: " + self.o.lib.oracle_last_error().decode())
        return OMat(self.o, self.block, R.nrows, P.ncols, None, None, None, handle=h)

    def export(self):
        nnz = self.nnz
        rp = np.zeros(self.nrows + 1, np.int64); ci = np.zeros(nnz, np.int32)
        va = np.zeros(nnz * self.block * self.block)
        self.o._chk(self.o.lib.oracle_mat_export(self.h, rp, ci, va))
        return rp, ci, va

    def _vb(self, vblock):
        return self.block if vblock is None else vblock

    def apply(self, x, vblock=None):
        vb = self._vb(vblock)
        y = np.zeros(self.nrows * vb)
        self.o._chk(self.o.lib.oracle_apply(self.h, y, _vec(x), vb))
        return y

    def apply_transposed(self, x):
        y = np.zeros(self.ncols * self.block)
        self.o._chk(self.o.lib.oracle_apply_transposed(self.h, y, _vec(x)))
        return y

    def matmul_minus(self, y, x, vblock=None):
        vb = self._vb(vblock)
        y = _vec(y).copy()
        self.o._chk(self.o.lib.oracle_matmul_minus(self.h, y, _vec(x), vb))
        return y

    def axpy(self, alpha, v, beta, w, vblock=None, dest=None):
        """dest = alpha*v + beta*A*w; pass dest (and v=None) for the in-place branch."""
        vb = self._vb(vblock)
        if dest is not None:
            d = _vec(dest).copy()
            self.o._chk(self.o.lib.oracle_axpy(self.h, d, alpha, d.ctypes.data_as(C.c_void_p), beta, _vec(w), vb))
            return d
        d = np.zeros(self.nrows * vb)
        vv = _vec(v) if v is not None else None
        vp = vv.ctypes.data_as(C.c_void_p) if vv is not None else None
        self.o._chk(self.o.lib.oracle_axpy(self.h, d, alpha, vp, beta, _vec(w), vb))
        return d

    def apply_ignore_zero_rows(self, dest, beta, w, vblock=None):
        vb = self._vb(vblock)
        d = _vec(dest).copy()
        self.o._chk(self.o.lib.oracle_apply_ignore_zero_rows(self.h, d, beta, _vec(w), vb))
        return d

    def jacobi(self, d, damp=1.0, block_inverse=True):
        c = np.zeros(self.nrows * self.block)
        self.o._chk(self.o.lib.oracle_jacobi(self.h, damp, int(block_inverse), c, _vec(d)))
        return c

    def gs(self, d, kind="ll", relax=1.0, c0=None):
        c = np.zeros(self.nrows * self.block) if c0 is None else _vec(c0).copy()
        self.o._chk(self.o.lib.oracle_gs(self.h, {"ll": 0, "ur": 1, "sgs": 2}[kind], relax, c, _vec(d)))
        return c

    def ilu(self, beta=0.0, sort_eps=1e-50) -> "OMat":
        """The ILU(0) / ILU(beta) factors of this matrix in one matrix (ilu.h:563-576)."""
        h = self.o.lib.oracle_ilu_factorize(self.h, beta, sort_eps)
        if not h:
            raise RuntimeError("oracle: " + self.o.lib.oracle_last_error().decode())
        return OMat(self.o, self.block, self.nrows, self.ncols, None, None, None, handle=h)

    def ilu_apply(self, d, inv_eps=1e-8):
        """c = U^-1 L^-1 d with self = the factor matrix (ILU::applyLU, ilu.h:593-599)."""
        c = np.zeros(self.nrows * self.block)
        self.o._chk(self.o.lib.oracle_ilu_apply(self.h, inv_eps, c, _vec(d)))
        return c

    def cuthill_mckee(self, reverse=True, preserve_consec=False):
        """new index of every old index (GetCuthillMcKeeOrder's defaults: reverse, not consecutive)."""
        ni = np.zeros(self.nrows, np.int64)
        self.o._chk(self.o.lib.oracle_cuthill_mckee(self.h, int(reverse), int(preserve_consec), ni))
        return ni

    def lu_solve(self, b):
        x = np.zeros(self.nrows * self.block)
        self.o._chk(self.o.lib.oracle_lu_solve(self.h, x, _vec(b)))
        return x


def make_desc(desc: dict) -> SolverDesc:
    """Translate a util.solver-style descriptor (scripts/util/solver_util.lua:423-575).

    {"type": "cg", "precond": {"type": "gmg", "smoother": {"type": "jac", "damp": 0.66},
      "cycle": "V", "preSmooth": 2, "postSmooth": 2, "baseLevel": 0, "baseSolver": "lu"},
     "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-10}}
    """
    d = SolverDesc()
    d.solver = SOLVER[desc.get("type", "cg")]
    cc = desc.get("convCheck", {})
    d.max_steps = cc.get("iterations", 100)
    d.min_defect = cc.get("absolute", 1e-12)
    d.rel_reduction = cc.get("reduction", 1e-6)
    # util.solver.defaults (solver_util.lua:423-575): linear / cg / bicgstab are preconditioned by ILU unless the
    # descriptor names a preconditioner; an explicit None (or "none") means no preconditioner at all
    pc = desc["precond"] if "precond" in desc else ("ilu" if desc.get("type", "cg") in ("linear", "cg", "bicgstab", "gmres") else None)
    if isinstance(pc, str):
        pc = None if pc == "none" else {"type": pc}
    d.precond = PRECOND[pc["type"] if pc else None]
    d.damp = 1.0
    # GMRES(restart): solver_util.lua:669-670 creates GMRES(5); BiCGStab: set_restart(n), 0 = never (bicgstab.h:161-163)
    d.restart = desc.get("restart", 0 if desc.get("type") == "bicgstab" else 5)
    d.ilu_beta = 0.0
    d.cycle, d.nu1, d.nu2 = 1, 2, 2
    d.smoother, d.smoother_damp = 1, 0.66
    d.base_solver, d.base_max_steps, d.base_min_defect, d.base_rel_reduction = 3, 1000, 1e-30, 1e-14
    if pc:
        if pc["type"] in ("jac", "jacobi"):
            d.damp = pc.get("damping", pc.get("damp", 0.66))
        elif pc["type"] in ("gs", "bgs", "sgs"):
            d.damp = pc.get("relax", 1.0)
        elif pc["type"] == "ilu":
            d.ilu_beta = pc.get("beta", 0.0)
        elif pc["type"] == "gmg":
            sm = pc.get("smoother", "gs")          # defaults.preconditioner.gmg: smoother "gs", preSmooth = postSmooth = 3
            if isinstance(sm, str):
                sm = {"type": sm}
            d.smoother = PRECOND[sm["type"]]
            d.ilu_beta = sm.get("beta", 0.0) if sm["type"] == "ilu" else 0.0
            d.smoother_damp = sm.get("damping", sm.get("damp", 0.66)) if d.smoother == 1 else sm.get("relax", 1.0)
            d.cycle = {"V": 1, "W": 2, "F": -1}[pc.get("cycle", "V")]
            d.nu1 = pc.get("preSmooth", 3)
            d.nu2 = pc.get("postSmooth", 3)
            d.base_lev = pc.get("baseLevel", 0)
            d.top_lev = pc["topLevel"]
            bs = pc.get("baseSolver", "lu")
            if isinstance(bs, str):
                bs = {"type": bs}
            d.base_solver = SOLVER[bs["type"]]
            bcc = bs.get("convCheck", {})
            d.base_max_steps = bcc.get("iterations", 1000)
            d.base_min_defect = bcc.get("absolute", 1e-30)
            d.base_rel_reduction = bcc.get("reduction", 1e-14)
    return d


class OSolver:
    """Solver built from a descriptor; ``levels`` maps level -> (A, P, R) OMat triples."""

    def __init__(self, orc: Oracle, desc: dict, A: OMat, levels: dict | None = None, smoother_matrices: dict | None = None,
                 precond_matrix: "OMat | None" = None, surface_map=None):
        self.o = orc
        self.desc = make_desc(desc)
        self.h = orc.lib.oracle_solver_create(C.byref(self.desc))
        if not self.h:
            raise RuntimeError("oracle: " + orc.lib.oracle_last_error().decode())
        self._keep = [A, levels]
        self.block = A.block
        if levels:
            for lev, (Al, Pl, Rl) in levels.items():
                orc._chk(orc.lib.oracle_solver_set_level(self.h, lev, Al.h, Pl.h if Pl else None, Rl.h if Rl else None))
        if smoother_matrices:   # level -> OMat the smoothers of that level are initialised with (parallel GS emulation)
            self._keep.append(smoother_matrices)
            for lev, S in smoother_matrices.items():
                orc._chk(orc.lib.oracle_solver_set_smoother_matrix(self.h, lev, S.h))
        if surface_map is not None:      # surface index of every top-level index (GMG only)
            sm = np.ascontiguousarray(surface_map, dtype=np.int32)
            orc._chk(orc.lib.oracle_solver_set_surface_map(self.h, sm.size, sm))
        if precond_matrix is not None:   # one-level preconditioner initialised with another matrix (parallel GS / ILU emulation)
            self._keep.append(precond_matrix)
            orc._chk(orc.lib.oracle_solver_set_precond_matrix(self.h, precond_matrix.h))
        orc._chk(orc.lib.oracle_solver_init(self.h, A.h))

    def __del__(self):
        try:
            self.o.lib.oracle_solver_destroy(self.h)
        except Exception:
            pass

    def apply(self, b, x0=None):
        """Returns (x, converged, defect history)."""
        b = _vec(b)
        x = np.zeros_like(b) if x0 is None else _vec(x0).copy()
        rc = self.o._chk(self.o.lib.oracle_solver_apply(self.h, x, b, self.block))
        hist = np.zeros(self.desc.max_steps + 2)
        n = self.o.lib.oracle_solver_history(self.h, hist, hist.size)
        return x, rc == 0, hist[:n].copy()

    def precond_apply(self, d):
        d = _vec(d)
        c = np.zeros_like(d)
        self.o._chk(self.o.lib.oracle_precond_apply(self.h, c, d, self.block))
        return c
