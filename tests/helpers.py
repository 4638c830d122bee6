"""Shared helpers of the test-suite (oracle-side hierarchy wiring, device buffers)."""
import ctypes as C

import os

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


ROUND_OFF_FLOOR = 1e-12      # a defect below this fraction of the start defect is round-off of the start defect
ROUND_OFF_ALLOWANCE = 1e-14  # absolute allowance (x start defect) for steps below the floor


def rel_hist_err(h_gpu, h_ref, allowance=None, floor=None):
    """max_k |h_gpu[k] - h_ref[k]| / |h_ref[k]| over the common prefix — north_star's "residual history within
    1e-10 relative per iteration", strict for every step whose defect is at least 1e-12 x the start defect (that
    covers every step of the BASELINE configurations, which stop at a reduction of 1e-10 / 1e-8).

    Steps BELOW 1e-12 x start defect (a solver that converges to machine precision in one or two steps: an exact
    ILU, LU) are round-off of the start defect itself — eps * ||b|| * amplification — and have no meaningful relative
    accuracy in any implementation; for those, and only those, differences up to 1e-14 x start defect are ignored.
    (Round 1 applied that allowance to EVERY step: at the end of a history that has dropped by 1e-10 it is a relative
    1e-4 and hid everything but the first steps.)  What a history above the floor can be held to is measured:
    tests/test_reduction_order.py records how far the REFERENCE's own history moves when only the summation order of
    its dot products changes, and sens_tol() turns that into the tolerance.
    UG4B200_RECORD_HIST_ERR=<file>: every evaluation is appended to that file."""
    n = min(len(h_gpu), len(h_ref))
    if not n:
        return 0.0
    h_gpu, h_ref = np.asarray(h_gpu[:n], float), np.asarray(h_ref[:n], float)
    diff = np.abs(h_gpu - h_ref)
    below = np.abs(h_ref) < (ROUND_OFF_FLOOR if floor is None else floor) * abs(h_ref[0])
    allow = (ROUND_OFF_ALLOWANCE if allowance is None else allowance) * abs(h_ref[0])
    diff = np.where(below, np.maximum(diff - allow, 0.0), diff)
    with np.errstate(all="ignore"):
        q = np.where(diff == 0.0, 0.0, diff / np.abs(h_ref))
    err = float(np.max(q))
    rec = os.environ.get("UG4B200_RECORD_HIST_ERR")
    if rec:
        import json
        with open(rec, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "steps": n - 1, "err": err,
                                "strict": float(np.max(np.where(h_gpu == h_ref, 0.0, np.abs(h_gpu - h_ref) / np.abs(h_ref)))),
                                "reduction": float(h_ref[-1] / h_ref[0])}) + "\n")
    return err


def osolver_sensitivity(orc, osol, b):
    """Per-step relative movement of the history of the oracle solver `osol` when its dot products / norms are summed
    by a pairwise tree instead of ugcore's sequential loop (everything else bit-identical); plus both histories."""
    _, _, h0 = osol.apply(b)
    orc.set_reduction_mode(1)
    try:
        _, _, h1 = osol.apply(b)
    finally:
        orc.set_reduction_mode(0)
    k = min(len(h0), len(h1))
    with np.errstate(all="ignore"):
        return np.abs(h0[:k] - h1[:k]) / np.abs(h0[:k]), h0, h1


def reduction_order_sensitivity(orc, prob, desc, b):
    """The same for a solver descriptor on a generated problem (tests/test_reduction_order.py)."""
    import oracle
    pc = desc.get("precond")
    if isinstance(pc, dict) and pc.get("type") == "gmg":
        levels = oracle_levels(orc, prob, pc["baseLevel"], pc["topLevel"])
        osol = oracle.OSolver(orc, desc, levels[pc["topLevel"]][0], levels)
    else:
        osol = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
    return osolver_sensitivity(orc, osol, b)


def sens_tol(orc, osol, b, base=1e-10, factor=10.0):
    """Tolerance for a GPU-vs-reference history comparison: north_star's 1e-10 per iteration, or — where the
    REFERENCE's own history moves by more than a tenth of that when only the summation order of its reductions
    changes (BiCGStab, GMRES, long unpreconditioned runs) — `factor` times that measured movement.  The GPU's
    reduction tree is one more summation order; it cannot be expected to land closer to the sequential sum than
    another valid order does."""
    sens, h0, _ = osolver_sensitivity(orc, osol, b)
    ok = np.isfinite(sens) & (np.abs(h0[:sens.size]) >= ROUND_OFF_FLOOR * abs(h0[0]))
    sens = sens[ok]
    return max(base, factor * float(sens.max())) if sens.size else base


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


# ---- colour-sorted orderings and the serial emulation of ugcore's parallel Gauss-Seidel ----

def greedy_color_perm(A):
    """The ordering the library's Gauss-Seidel uses without a user colouring: greedy first-fit
    colouring of the stored pattern in row order (host helper of the C ABI), colours sorted,
    stable inside a colour.  Returns (perm old -> new, colour offsets)."""
    from ugcore_b200 import capi
    n = A.nrows
    color = np.zeros(max(n, 1), np.int32)
    nc = C.c_int()
    capi.dev.ug4b200_color_greedy(n, A.rowptr.ctypes.data_as(C.c_void_p), A.cols.ctypes.data_as(C.c_void_p),
                                  color.ctypes.data_as(C.c_void_p), C.byref(nc))
    color = color[:n]
    order = np.argsort(color, kind="stable")
    perm = np.empty(n, np.int64)
    perm[order] = np.arange(n)
    cptr = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=nc.value))]).astype(np.int64)
    return perm, cptr


def permute_crs(A, prow, pcol, keep=None):
    """B(prow[r], pcol[c]) = A(r, c); explicit zeros kept, columns sorted.  keep: boolean mask over
    the stored entries (dropped entries disappear from the pattern)."""
    from ugcore_b200.problems import Crs
    bb = A.block * A.block
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
    cols = np.asarray(A.cols)
    vals = np.asarray(A.vals).reshape(-1, bb)
    if keep is not None:
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    pr_, pc_ = np.asarray(prow)[rows], np.asarray(pcol)[cols]
    key = np.lexsort((pc_, pr_))
    rp = np.concatenate([[0], np.cumsum(np.bincount(pr_, minlength=A.nrows))]).astype(np.int64)
    return Crs(A.nrows, A.ncols, A.block, rp, pc_[key].astype(np.int32), vals[key].ravel().copy())


def with_dirichlet_rows(A, rows):
    """SetDirichletRow (sparsematrix_util.h:878-897): all blocks of the rows 0, diagonal block 1."""
    from ugcore_b200.problems import Crs
    b, bb = A.block, A.block * A.block
    vals = np.array(A.vals, dtype=np.float64).reshape(-1, bb)
    for r in rows:
        lo, hi = A.rowptr[r], A.rowptr[r + 1]
        vals[lo:hi] = 0.0
        d = lo + int(np.flatnonzero(np.asarray(A.cols[lo:hi]) == r)[0])
        for t in range(b):
            vals[d, t + b * t] = 1.0
    return Crs(A.nrows, A.ncols, A.block, np.array(A.rowptr), np.array(A.cols), vals.ravel().copy())


def parallel_gs_global_model(locals_, gprob, lev):
    """ugcore's parallel Gauss-Seidel (gauss_seidel.h:134-142, 204-215) on one partitioned level, in
    GLOBAL terms.  Every rank sweeps over its consistent matrix with the h-slave rows set to Dirichlet
    rows, on the unique defect; the slaves' corrections are exactly 0.  Hence one step is a sweep over the
    global matrix without the couplings between DoFs of different h-masters, in any global order that
    keeps, inside each master's block, the order of that rank's sweep (its colour-sorted local order).

    locals_: the ranks' local problems in rank order.  Returns (gperm: global id -> position,
    keep: mask over the stored entries of the global level matrix)."""
    from ugcore_b200 import dist as ugdist
    gA = gprob.matrix(lev)
    owner = np.full(gA.nrows, -1, np.int64)
    order = []
    for r, p in enumerate(locals_):
        gid = p.global_ids(lev)
        own = ugdist.owned_mask(p, lev, r)
        assert np.all(owner[gid[own]] == -1)
        owner[gid[own]] = r
        perm, _ = greedy_color_perm(p.matrix(lev))          # pattern only: the Dirichlet rows keep theirs
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        order.append(gid[inv][own[inv]])                    # owned DoFs in the order of the local sweep
    assert np.all(owner >= 0)
    order = np.concatenate(order)
    gperm = np.empty(gA.nrows, np.int64)
    gperm[order] = np.arange(gA.nrows)
    rows = np.repeat(np.arange(gA.nrows), np.diff(gA.rowptr))
    keep = owner[rows] == owner[np.asarray(gA.cols)]
    return gperm, keep


def partitioned_gs_oracle(orc, desc, refs, part, gather, problem=0, **kw):
    """Serial oracle of the partitioned GMG with Gauss-Seidel smoothing: levels above `gather` smooth
    like ugcore's parallel Gauss-Seidel (parallel_gs_global_model), the gathered levels below run the
    serial multicolour sweep on the global matrices.  Everything is permuted into the sweep order
    (the oracle's Gauss-Seidel is the reference's lexicographic gs_step over the matrix it is given).
    Returns (solve(b_global) -> (x_global, ok, history), global problem)."""
    import oracle
    from ugcore_b200 import dist as ugdist
    world = part[0] * part[1] * part[2]
    base = desc["precond"].get("baseLevel", 0)
    locals_ = [ugdist.local_problem(refs, part, r, problem=problem, **kw) for r in range(world)]
    gprob = ugdist.global_problem(refs, part, problem=problem, **kw)
    perms, keeps = {}, {}
    for l in range(base, refs + 1):
        if l > gather:
            perms[l], keeps[l] = parallel_gs_global_model(locals_, gprob, l)
        elif l > base:
            perms[l], _ = greedy_color_perm(gprob.matrix(l))
        else:
            perms[l] = np.arange(gprob.matrix(l).nrows)
    lv, sm = {}, {}
    for l in range(base, refs + 1):
        A = orc.matrix(permute_crs(gprob.matrix(l), perms[l], perms[l]))
        P = R = None
        if l > base:
            P = orc.matrix(permute_crs(gprob.prolongation(l), perms[l], perms[l - 1]))
            R = orc.matrix(permute_crs(gprob.restriction(l), perms[l - 1], perms[l]))
        lv[l] = (A, P, R)
        if l in keeps:
            sm[l] = orc.matrix(permute_crs(gprob.matrix(l), perms[l], perms[l], keep=keeps[l]))
    d = dict(desc)
    d["precond"] = dict(desc["precond"], topLevel=refs, baseLevel=base)
    osol = oracle.OSolver(orc, d, lv[refs][0], lv, smoother_matrices=sm)
    b = gprob.block
    top = np.repeat(perms[refs] * b, b) + np.tile(np.arange(b), perms[refs].size)   # component-wise permutation

    def solve(bg):
        bp = np.empty_like(bg)
        bp[top] = bg
        xp, ok, h = osol.apply(bp)
        return xp[top], ok, h

    return solve, gprob


def partitioned_onelevel_oracle(orc, desc, refs, part, problem=0, colored=False, **kw):
    """Serial oracle of a partitioned Krylov solve with a ONE-LEVEL Gauss-Seidel / ILU preconditioner in ugcore's
    parallel mode (gauss_seidel.h:134-142, ilu.h:536-543): the preconditioner works on the global matrix without
    the couplings between DoFs of different h-masters, every master's block in the order of that rank's sweep
    (colored: its greedy multicolour order — Gauss-Seidel, ILU with ordering "multicolor"; else the rank's own
    numbering — ILU in natural ordering).  Returns (solve(b_global) -> (x_global, ok, history), global problem)."""
    import oracle
    from ugcore_b200 import dist as ugdist
    world = part[0] * part[1] * part[2]
    locals_ = [ugdist.local_problem(refs, part, r, problem=problem, **kw) for r in range(world)]
    gprob = ugdist.global_problem(refs, part, problem=problem, **kw)
    gA = gprob.matrix(refs)
    if colored:
        gperm, keep = parallel_gs_global_model(locals_, gprob, refs)
    else:
        owner = np.full(gA.nrows, -1, np.int64)
        order = []
        for r, p in enumerate(locals_):
            gid = p.global_ids(refs)
            own = ugdist.owned_mask(p, refs, r)
            owner[gid[own]] = r
            order.append(gid[own])
        order = np.concatenate(order)
        gperm = np.empty(gA.nrows, np.int64)
        gperm[order] = np.arange(gA.nrows)
        rows = np.repeat(np.arange(gA.nrows), np.diff(gA.rowptr))
        keep = owner[rows] == owner[np.asarray(gA.cols)]
    A = orc.matrix(permute_crs(gA, gperm, gperm))
    M = orc.matrix(permute_crs(gA, gperm, gperm, keep=keep))
    osol = oracle.OSolver(orc, desc, A, precond_matrix=M)
    b = gprob.block
    top = np.repeat(gperm * b, b) + np.tile(np.arange(b), gperm.size)

    def solve(bg):
        bp = np.empty_like(bg)
        bp[top] = bg
        xp, ok, h = osol.apply(bp)
        return xp[top], ok, h

    return solve, gprob
