"""Python face of the descriptor-level driver (include/ug4b200_solver.h).

``Solver(desc, problem)`` is what ``util.solver.CreateSolver(desc)`` followed by
``solver:init(A, u)`` is in a ugcore Lua script (scripts/util/solver_util.lua:602,
:1182-1210); ``solver.apply(b)`` is ``solver:apply(u, b)``.  The descriptor uses the same
vocabulary as solver_util.lua (type, precond, smoother, cycle, preSmooth, postSmooth,
baseLevel, baseSolver, convCheck{iterations, absolute, reduction}).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, check_host, host, dev

SOLVER = {"cg": 0, "bicgstab": 1, "linear": 2, "lu": 3, "coarse_cg": 4, "gmres": 5}
PRECOND = {None: 0, "none": 0, "jac": 1, "jacobi": 1, "gs": 2, "bgs": 3, "sgs": 4, "gmg": 5, "ilu": 6}
ILU_ORDER = {None: 0, "natural": 0, "none": 0, "cmk": 1, "cuthill-mckee": 1, "multicolor": 2, "multicolour": 2}

_host_ready = False


def device_available() -> bool:
    """True if a CUDA device can be opened (never falls back to the CPU)."""
    ctx = C.c_void_p()
    rc = dev.ug4b200_ctx_create(0, None, C.byref(ctx))
    if rc == 0:
        dev.ug4b200_ctx_destroy(ctx)
    return rc == 0


def host_init(device: int = -1, stream=None) -> None:
    """Create the process-wide device context (GPUManager). One process per GPU."""
    global _host_ready
    if not _host_ready:
        check_host(host.ug4b200_host_init(device, stream))
        _host_ready = True


def host_ctx():
    host_init()
    return host.ug4b200_host_ctx()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _lookup(table: dict, key, what: str):
    try:
        return table[key]
    except KeyError:
        raise ValueError(f"unknown {what} {key!r}; known: {sorted(k for k in table if isinstance(k, str))}") from None


def make_desc(desc: dict, block: int = 1, flags: int = 0) -> capi.SolverDesc:
    d = capi.SolverDesc()
    d.block = block
    d.solver = _lookup(SOLVER, desc.get("type", "cg"), "linear solver")
    cc = desc.get("convCheck", {})
    d.max_steps = cc.get("iterations", 100)
    d.min_defect = cc.get("absolute", 1e-12)
    d.rel_reduction = cc.get("reduction", 1e-6)
    # util.solver.defaults (solver_util.lua:423-575): linear / cg / bicgstab are preconditioned by ILU unless the
    # descriptor names a preconditioner; an explicit None (or "none") means no preconditioner at all
    pc = desc["precond"] if "precond" in desc else ("ilu" if desc.get("type", "cg") in ("linear", "cg", "bicgstab", "gmres") else None)
    if isinstance(pc, str):
        pc = None if pc == "none" else {"type": pc}
    d.precond = _lookup(PRECOND, pc["type"] if pc else None, "preconditioner")
    d.damp = 1.0
    # GMRES(restart): solver_util.lua:669-670 creates GMRES(5); BiCGStab: set_restart(n), 0 = never (bicgstab.h:161-163)
    d.restart = desc.get("restart", 0 if desc.get("type") == "bicgstab" else 5)
    d.ilu_beta, d.ilu_order = 0.0, 0
    d.cycle, d.nu1, d.nu2 = 1, 2, 2
    d.smoother, d.smoother_damp = 1, 0.66
    d.base_solver, d.base_max_steps, d.base_min_defect, d.base_rel_reduction = 3, 1000, 1e-30, 1e-14
    if pc:
        if pc["type"] in ("jac", "jacobi"):
            d.damp = pc.get("damping", pc.get("damp", 0.66))
        elif pc["type"] in ("gs", "bgs", "sgs"):
            d.damp = pc.get("relax", 1.0)
        elif pc["type"] == "ilu":                # solver_util.lua: ilu = {beta, sort, ...}
            d.ilu_beta = pc.get("beta", 0.0)
            d.ilu_order = _lookup(ILU_ORDER, pc.get("ordering", "cmk" if pc.get("sort") else None), "ILU ordering")
        elif pc["type"] == "gmg":
            sm = pc.get("smoother", "gs")          # defaults.preconditioner.gmg: smoother "gs", preSmooth = postSmooth = 3
            if isinstance(sm, str):
                sm = {"type": sm}
            d.smoother = _lookup(PRECOND, sm["type"], "smoother")
            if d.smoother == PRECOND["gmg"]:
                raise ValueError("a GMG cannot smooth a GMG")
            if sm["type"] == "ilu":
                d.ilu_beta = sm.get("beta", 0.0)
                d.ilu_order = _lookup(ILU_ORDER, sm.get("ordering", "cmk" if sm.get("sort") else None), "ILU ordering")
            d.smoother_damp = sm.get("damping", sm.get("damp", 0.66)) if d.smoother == 1 else sm.get("relax", 1.0)
            d.cycle = {"V": 1, "W": 2, "F": -1}[pc.get("cycle", "V")]
            d.nu1 = pc.get("preSmooth", 3)
            d.nu2 = pc.get("postSmooth", 3)
            d.base_lev = pc.get("baseLevel", 0)
            d.top_lev = pc["topLevel"]
            d.gather_lev = pc.get("gatherLevel", d.base_lev)
            bs = pc.get("baseSolver", "lu")
            if isinstance(bs, str):
                bs = {"type": bs}
            d.base_solver = {"lu": 3, "cg": 4, "coarse_cg": 4}[bs["type"]]
            bcc = bs.get("convCheck", {})
            d.base_max_steps = bcc.get("iterations", 1000)
            d.base_min_defect = bcc.get("absolute", 1e-30)
            d.base_rel_reduction = bcc.get("reduction", 1e-14)
    d.flags = flags | desc.get("flags", 0)
    return d


class DeviceBuffer:
    """A device array of doubles owned by the host layer's context."""

    def __init__(self, n: int):
        self.ctx = host_ctx()
        self.n = int(n)
        p = C.c_void_p()
        check(dev.ug4b200_alloc(self.ctx, max(self.n, 1) * 8, C.byref(p)), self.ctx)
        self.ptr = p

    @classmethod
    def from_numpy(cls, a) -> "DeviceBuffer":
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = cls(a.size)
        check(dev.ug4b200_h2d(b.ctx, b.ptr, _ptr(a), a.size * 8), b.ctx)
        check(dev.ug4b200_sync(b.ctx), b.ctx)
        return b

    def to_numpy(self):
        a = np.empty(self.n)
        check(dev.ug4b200_d2h(self.ctx, _ptr(a), self.ptr, self.n * 8), self.ctx)
        return a

    def __del__(self):
        try:
            dev.ug4b200_free(self.ctx, self.ptr)
        except Exception:
            pass


def cuthill_mckee(A, reverse: bool = True, preserve_consec: bool = False):
    """new index of every old index: ugcore's native Cuthill-McKee on the matrix graph (GetCuthillMcKeeOrder,
    ugbase/lib_algebra/algebra_common/permutation_util.h:96-114; same result as the reference).  Host only."""
    ni = np.zeros(A.nrows, np.int64)
    rp, ci = np.ascontiguousarray(A.rowptr, np.int64), np.ascontiguousarray(A.cols, np.int32)
    check_host(host.ug4b200_host_cuthill_mckee(A.nrows, _ptr(rp), _ptr(ci), int(reverse), int(preserve_consec), _ptr(ni)))
    return ni


def permute_crs(A, prow, pcol):
    """B(prow[r], pcol[c]) = A(r, c) with sorted rows, explicit zeros kept (SetMatrixAsPermutation,
    permutation_util.h:50-64, generalised to rectangular transfers)."""
    from .problems import Crs
    bb = A.block * A.block
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
    pr_, pc_ = np.asarray(prow)[rows], np.asarray(pcol)[np.asarray(A.cols)]
    key = np.lexsort((pc_, pr_))
    rp = np.concatenate([[0], np.cumsum(np.bincount(pr_, minlength=A.nrows))]).astype(np.int64)
    vals = np.asarray(A.vals).reshape(-1, bb)[key].ravel().copy()
    return Crs(A.nrows, A.ncols, A.block, rp, pc_[key].astype(np.int32), vals)


def reorder_hierarchy(A, levels, order):
    """DoF reordering before upload: every level is renumbered by ``order`` —
    "cmk" / "rcmk" (Cuthill-McKee / reverse Cuthill-McKee of the level matrix) or a dict level -> permutation
    (new index of every old index).  Returns (A', levels', perms); the level matrices, P (rows: fine, columns:
    coarse numbering) and R are permuted consistently."""
    top = max(levels) if levels else None
    mats = {l: t[0] for l, t in levels.items()} if levels else {None: A}
    if levels and mats.get(top) is None:
        mats[top] = A
    perms = {}
    for l, M in mats.items():
        if isinstance(order, dict):
            perms[l] = np.asarray(order[l], dtype=np.int64)
        elif M is None:
            raise ValueError(f"reordering needs the level matrix of level {l} (not available with rap=True)")
        elif order in ("cmk", "rcmk"):
            perms[l] = cuthill_mckee(M, reverse=(order == "rcmk"))
        else:
            raise ValueError(f"unknown ordering {order!r}")
    if not levels:
        return permute_crs(A, perms[None], perms[None]), None, perms
    out = {}
    for l, (Al, Pl, Rl) in levels.items():
        out[l] = (permute_crs(Al, perms[l], perms[l]) if Al is not None else None,
                  permute_crs(Pl, perms[l], perms[l - 1]) if Pl is not None else None,
                  permute_crs(Rl, perms[l - 1], perms[l]) if Rl is not None else None)
    return permute_crs(A, perms[top], perms[top]), out, perms


class Solver:
    """CG / BiCGStab / LinearSolver / GMRES with Jacobi / GS / ILU / GMG preconditioning on the GPU.

    ``A`` is a host CRS (``problems.Crs``); ``levels`` maps level -> (A_l, P_l, R_l) for GMG
    (P_l, R_l None on the base level).  ``Solver.from_problem`` wires a synthetic hierarchy.
    ``order``: DoF reordering applied before upload (``reorder_hierarchy``); ``apply`` takes and returns
    vectors in the caller's numbering.  ``surface_map``: surface index of every top-level index when ``A`` (and the
    vectors) are in a surface numbering that differs from the level numbering of ``levels`` (ugcore's vSurfLevelMap).
    """

    def __init__(self, desc: dict, A, levels: dict | None = None, flags: int = 0, order=None, surface_map=None):
        host_init()
        if surface_map is not None and order is not None:
            raise ValueError("order= renumbers surface and levels alike; it cannot be combined with a surface map")
        self.perm = None
        if order is not None:
            A, levels, perms = reorder_hierarchy(A, levels, order)
            p = perms[max(levels)] if levels else perms[None]
            self.perm = np.repeat(p * A.block, A.block) + np.tile(np.arange(A.block), p.size)
        self.block = A.block
        pc = desc.get("precond")
        # gmg:set_rap(true): the level operators below the top level are Galerkin products computed at init
        rap = isinstance(pc, dict) and bool(pc.get("rap", False))
        if rap:
            flags |= capi.FLAG_RAP
        self.desc = make_desc(desc, self.block, flags)
        self.h = C.c_void_p()
        check_host(host.ug4b200_solver_create(C.byref(self.desc), C.byref(self.h)))
        self.n = A.nrows * A.block
        self._keep = [A, levels]
        check_host(host.ug4b200_solver_set_matrix(self.h, A.nrows, A.ncols, _ptr(A.rowptr), _ptr(A.cols), _ptr(A.vals)))
        if levels:
            for lev, (Al, Pl, Rl) in sorted(levels.items()):
                top = lev == self.desc.top_lev
                # the top level reuses the surface matrix — unless the surface is numbered differently (surface_map)
                skip = (top and surface_map is None) or (rap and not top) or Al is None
                nrows = Al.nrows if Al is not None else (Pl.nrows if Pl else levels[lev + 1][1].ncols)
                check_host(host.ug4b200_solver_set_level(
                    self.h, lev, nrows,
                    None if skip else _ptr(Al.rowptr), None if skip else _ptr(Al.cols), None if skip else _ptr(Al.vals),
                    Pl.ncols if Pl else 0,
                    _ptr(Pl.rowptr) if Pl else None, _ptr(Pl.cols) if Pl else None, _ptr(Pl.vals) if Pl else None,
                    _ptr(Rl.rowptr) if Rl else None, _ptr(Rl.cols) if Rl else None, _ptr(Rl.vals) if Rl else None))
        if surface_map is not None:
            self.set_surface_map(surface_map)
        self._inited = False

    @classmethod
    def from_problem(cls, desc: dict, prob, flags: int = 0, order=None) -> "Solver":
        desc = dict(desc)
        pc = desc.get("precond")
        levels = None
        if isinstance(pc, dict) and pc.get("type") == "gmg":
            pc = dict(pc)
            pc.setdefault("topLevel", prob.num_refs)
            pc.setdefault("baseLevel", prob.base_lev)
            desc["precond"] = pc
            levels = {}
            for lev in range(pc["baseLevel"], pc["topLevel"] + 1):
                base = lev == pc["baseLevel"]
                levels[lev] = (prob.matrix(lev), None if base else prob.prolongation(lev),
                               None if base else prob.restriction(lev))
        s = cls(desc, prob.matrix(desc["precond"]["topLevel"] if levels else None), levels, flags, order=order)
        s._keep.append(prob)
        return s

    def set_surface_map(self, surf_index_of_level_index):
        """GMG on a hierarchy whose SURFACE numbering of the top level differs from its LEVEL numbering
        (vSurfLevelMap): ``A`` and the vectors of ``apply`` are in surface numbering, ``levels`` in level numbering."""
        m = np.ascontiguousarray(surf_index_of_level_index, dtype=np.int32)
        check_host(host.ug4b200_solver_set_surface_map(self.h, m.size, _ptr(m)))

    def set_debug_dir(self, path, positions=None, dim: int = 3, precision: int = 0):
        """solver:set_debug(writer): CG then leaves CG_Residual_iterNNN.vec / CG_Solution_iterNNN.vec in ``path`` after
        every step, as a ugcore run with a debug writer does (cg.h:124, 195, 273-280).  ``path`` None: off."""
        if path is None:
            check_host(host.ug4b200_solver_set_debug_dir(self.h, None, None, 0, dim, precision))
            return
        pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        check_host(host.ug4b200_solver_set_debug_dir(self.h, str(path).encode(), _ptr(pos), 0 if pos is None else pos.shape[0], dim, precision))

    def set_coloring(self, perm, color_ptr, lev: int = -1):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        cp = np.ascontiguousarray(color_ptr, dtype=np.int64)
        check_host(host.ug4b200_solver_set_coloring(self.h, lev, perm.size, _ptr(perm), cp.size - 1, _ptr(cp)))

    def set_layouts(self, lev, neigh_rank, neigh_ptr, indices, nlocal):
        nr = np.ascontiguousarray(neigh_rank, dtype=np.int32)
        npt = np.ascontiguousarray(neigh_ptr, dtype=np.int64)
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        check_host(host.ug4b200_solver_set_layouts(self.h, lev, nr.size, _ptr(nr), _ptr(npt), _ptr(idx), nlocal))

    def set_smoother_matrix(self, lev, Acons):
        """Partitioned Gauss-Seidel: the level matrix made consistent on the interface rows
        (``dist.make_consistent``); the smoother sets the rows of the h-slaves to Dirichlet rows itself."""
        self._keep.append(Acons)
        check_host(host.ug4b200_solver_set_smoother_matrix(self.h, lev, Acons.nrows, _ptr(Acons.rowptr), _ptr(Acons.cols),
                                                           _ptr(Acons.vals)))

    def set_gathered_base(self, Aglobal, local_to_global):
        l2g = np.ascontiguousarray(local_to_global, dtype=np.int32)
        self._keep.append(Aglobal)
        check_host(host.ug4b200_solver_set_gathered_base(self.h, Aglobal.nrows, _ptr(Aglobal.rowptr), _ptr(Aglobal.cols),
                                                         _ptr(Aglobal.vals), l2g.size, _ptr(l2g)))

    def set_gathered_level(self, lev, A, P, R, nrows):
        """GLOBAL level matrix / transfers of a gathered level (A None on the gather level itself)."""
        self._keep += [A, P, R]
        check_host(host.ug4b200_solver_set_gathered_level(
            self.h, lev, nrows,
            _ptr(A.rowptr) if A else None, _ptr(A.cols) if A else None, _ptr(A.vals) if A else None,
            P.ncols if P else 0,
            _ptr(P.rowptr) if P else None, _ptr(P.cols) if P else None, _ptr(P.vals) if P else None,
            _ptr(R.rowptr) if R else None, _ptr(R.cols) if R else None, _ptr(R.vals) if R else None))

    def set_matrix(self, A, levels: dict | None = None):
        """A new operator for the same solver object — what ``solver:init(J, u)`` sees in every Newton or time step
        (the matrix was re-assembled).  The next ``apply`` re-runs ``init``: preconditioner preprocess, uploads, and
        the CUDA graphs of the Krylov loops are re-captured (they are tied to the generation of the device data)."""
        if self.perm is not None:
            raise ValueError("set_matrix: a solver created with order= keeps its permutation; create a new Solver")
        if A.nrows * A.block != self.n:
            raise ValueError("set_matrix: the number of unknowns must not change")
        self._keep[0] = A
        check_host(host.ug4b200_solver_set_matrix(self.h, A.nrows, A.ncols, _ptr(A.rowptr), _ptr(A.cols), _ptr(A.vals)))
        if levels:
            self._keep.append(levels)
            for lev, (Al, Pl, Rl) in sorted(levels.items()):
                skip = lev == self.desc.top_lev or Al is None
                nrows = Al.nrows if Al is not None else Pl.nrows
                check_host(host.ug4b200_solver_set_level(
                    self.h, lev, nrows,
                    None if skip else _ptr(Al.rowptr), None if skip else _ptr(Al.cols), None if skip else _ptr(Al.vals),
                    Pl.ncols if Pl else 0,
                    _ptr(Pl.rowptr) if Pl else None, _ptr(Pl.cols) if Pl else None, _ptr(Pl.vals) if Pl else None,
                    _ptr(Rl.rowptr) if Rl else None, _ptr(Rl.cols) if Rl else None, _ptr(Rl.vals) if Rl else None))
        self._inited = False
        return self

    def init(self):
        check_host(host.ug4b200_solver_init(self.h))
        self._inited = True
        return self

    def apply(self, b, x0=None):
        """solver:apply(u, b) with host vectors. Returns (x, converged, defect history)."""
        if not self._inited:
            self.init()
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        # no start vector given: it is 0 and is set on the device, not copied there
        fn = host.ug4b200_solver_apply_zero_guess if x0 is None else host.ug4b200_solver_apply
        if self.perm is not None:      # Pv[perm[i]] = v[i]  (SetVectorAsPermutation, permutation_util.h:72-78)
            bp, xp = np.empty_like(b), np.empty_like(x)
            bp[self.perm], xp[self.perm] = b, x
            rc = check_host(fn(self.h, _ptr(xp), _ptr(bp)))
            return xp[self.perm], rc == 0, self.history()
        rc = check_host(fn(self.h, _ptr(x), _ptr(b)))
        return x, rc == 0, self.history()

    def apply_pinned(self, x_ptr: int, b_ptr: int, zero_guess: bool = False) -> bool:
        """Same through raw host pointers (e.g. pinned torch tensors): x in/out, b in.  zero_guess: the start vector is 0
        (set on the device; x is output only)."""
        if self.perm is not None:
            raise ValueError("apply_pinned works in the solver's own numbering: permute the vectors (Solver.perm) or use apply")
        if not self._inited:
            self.init()
        fn = host.ug4b200_solver_apply_zero_guess if zero_guess else host.ug4b200_solver_apply
        return check_host(fn(self.h, C.c_void_p(x_ptr), C.c_void_p(b_ptr))) == 0

    def apply_device(self, x_dev, b_dev) -> bool:
        """Device-resident solve; x_dev / b_dev are DeviceBuffer or raw device pointers."""
        if not self._inited:
            self.init()
        xp = x_dev.ptr if isinstance(x_dev, DeviceBuffer) else C.c_void_p(x_dev)
        bp = b_dev.ptr if isinstance(b_dev, DeviceBuffer) else C.c_void_p(b_dev)
        return check_host(host.ug4b200_solver_apply_device(self.h, xp, bp)) == 0

    def precond_apply(self, d):
        if not self._inited:
            self.init()
        d = np.ascontiguousarray(d, dtype=np.float64)
        c = np.zeros(self.n)
        if self.perm is not None:
            dp = np.empty_like(d)
            dp[self.perm] = d
            check_host(host.ug4b200_solver_precond_apply(self.h, _ptr(c), _ptr(dp)))
            return c[self.perm]
        check_host(host.ug4b200_solver_precond_apply(self.h, _ptr(c), _ptr(d)))
        return c

    @property
    def steps(self) -> int:
        return host.ug4b200_solver_steps(self.h)

    @property
    def defect(self) -> float:
        return host.ug4b200_solver_defect(self.h)

    def history(self):
        buf = np.zeros(self.desc.max_steps + 2)
        n = host.ug4b200_solver_history(self.h, _ptr(buf), buf.size)
        return buf[:n].copy()

    def launch_count(self) -> int:
        n = C.c_int64()
        dev.ug4b200_launch_count(host_ctx(), C.byref(n))
        return n.value

    def __del__(self):
        try:
            host.ug4b200_solver_destroy(self.h)
        except Exception:
            pass
