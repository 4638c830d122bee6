"""Partitioned (multi-GPU) set-up: one process per GPU, structured box partition.

ugcore's parallel model is kept (Model A): element-wise partition, interface
vertices duplicated on every rank that touches them, ADDITIVE matrices (each rank assembles
its own elements only, lib_algebra/parallelization/parallel_matrix_impl.h:88-114), vectors
carry consistent / additive / unique storage types.  SpMV needs no communication; smoother
corrections are made consistent by an interface exchange (AdditiveToConsistent,
parallelization_util.h:159-191), norms / dots all-reduce one double.

This module holds the host-side logic (pure numpy, unit-tested on CPU with gloo):
  * the process grid and each rank's box,
  * horizontal interface lists per level (IndexLayout: per neighbour the shared DoFs, in an
    order both sides agree on — ascending global id),
  * the gathered base solve map (mg_solver_impl.hpp:2003-2070).
torch.distributed is only plumbing: it carries the NCCL unique id to the other ranks.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import problems as pr


def rank_to_coord(rank: int, part):
    """x fastest."""
    return (rank % part[0], (rank // part[0]) % part[1], rank // (part[0] * part[1]))


def coord_to_rank(coord, part):
    return coord[0] + part[0] * (coord[1] + part[1] * coord[2])


def interfaces(prob: pr.Problem, lev: int):
    """Horizontal interfaces of one level of this rank's sub-box.

    Returns (neigh_rank[int32], neigh_ptr[int64], indices[int32]) with neighbours sorted by
    rank and, inside a neighbour, DoFs sorted by global id (identical order on both sides).
    """
    part, coord = prob.part, prob.coord
    dims = prob.dims(lev)
    d2l = prob.dof_to_lex(lev)
    gid = prob.global_ids(lev)
    lex_to_dof = np.empty(d2l.size, np.int64)
    lex_to_dof[d2l] = np.arange(d2l.size)
    ii = np.arange(dims[0])[:, None, None]
    jj = np.arange(dims[1])[None, :, None]
    kk = np.arange(dims[2])[None, None, :]
    lex = (ii + dims[0] * (jj + dims[1] * kk))
    out = []
    for off in itertools.product((-1, 0, 1), repeat=3):
        if off == (0, 0, 0):
            continue
        nc = tuple(coord[d] + off[d] for d in range(3))
        if any(nc[d] < 0 or nc[d] >= part[d] for d in range(3)):
            continue
        if any(off[d] != 0 and d >= prob.dim for d in range(3)):
            continue
        sl = []
        for d in range(3):
            if off[d] == -1:
                sl.append(slice(0, 1))
            elif off[d] == 1:
                sl.append(slice(dims[d] - 1, dims[d]))
            else:
                sl.append(slice(None))
        dofs = lex_to_dof[lex[tuple(sl)].ravel()]
        dofs = dofs[np.argsort(gid[dofs], kind="stable")]
        out.append((coord_to_rank(nc, part), dofs))
    out.sort(key=lambda t: t[0])
    ranks = np.array([r for r, _ in out], np.int32)
    ptr = np.concatenate([[0], np.cumsum([len(d) for _, d in out])]).astype(np.int64)
    idx = np.concatenate([d for _, d in out]).astype(np.int32) if out else np.zeros(0, np.int32)
    return ranks, ptr, idx


def multiplicity(prob: pr.Problem, lev: int):
    """Number of ranks holding a copy of each local DoF (1 in the interior)."""
    ranks, ptr, idx = interfaces(prob, lev)
    m = np.ones(prob.matrix(lev).nrows, np.int32)
    np.add.at(m, idx, 1)
    return m


def owned_mask(prob: pr.Problem, lev: int, rank: int):
    """True where this rank is the h-master (lowest rank among the sharers)."""
    ranks, ptr, idx = interfaces(prob, lev)
    own = np.ones(prob.matrix(lev).nrows, bool)
    for p, r in enumerate(ranks):
        if r < rank:
            own[idx[ptr[p]:ptr[p + 1]]] = False
    return own


def consistent_contributions(A: pr.Crs, ranks, ptr, idx) -> dict:
    """What this rank sends to each neighbour so that the neighbour can make its copy of the level
    matrix consistent: every stored entry (i, j) whose two DoFs both lie in the interface to that
    neighbour, addressed by the POSITIONS of i and j in the interface list (the order both sides
    agree on) — the role of the global AlgebraIDs in ComPol_MatAddRowsOverlap0
    (lib_algebra/parallelization/parallelization_util.h:100-150).
    Returns {neighbour rank: (pos_i[int32], pos_j[int32], vals[k, block*block])}."""
    bb = A.block * A.block
    rows = np.repeat(np.arange(A.nrows, dtype=np.int64), np.diff(A.rowptr))
    vals = np.asarray(A.vals).reshape(-1, bb)
    out = {}
    for q, r in enumerate(ranks):
        shared = np.asarray(idx[ptr[q]:ptr[q + 1]], dtype=np.int64)
        pos = np.full(A.nrows, -1, np.int64)
        pos[shared] = np.arange(shared.size)
        pi, pj = pos[rows], pos[A.cols]
        sel = (pi >= 0) & (pj >= 0)
        out[int(r)] = (pi[sel].astype(np.int32), pj[sel].astype(np.int32), vals[sel].copy())
    return out


def apply_contributions(A: pr.Crs, rank: int, ranks, ptr, idx, received: dict) -> pr.Crs:
    """The consistent level matrix of this rank: the additive local matrix plus the neighbours'
    entries of the shared rows (MatMakeConsistentOverlap0 / MakeConsistent with overlap 0,
    parallelization_util.h:126-150, parallel_matrix_overlap_impl.h:438-459).  Every copy of an entry is
    the sum of the contributions in ASCENDING RANK order (own contribution at the place of the own
    rank), so all copies of a row are bitwise identical — the convention of the device-side vector
    exchange (csrc/comm.cu).  A connection that only a neighbour stores is inserted
    (ComPol_MatAddRowsOverlap0 adds into mat(i, j))."""
    bb = A.block * A.block
    n = A.nrows
    rows = [np.repeat(np.arange(n, dtype=np.int64), np.diff(A.rowptr))]
    cols = [np.asarray(A.cols, dtype=np.int64)]
    vals = [np.asarray(A.vals).reshape(-1, bb)]
    src = [np.full(A.cols.size, rank, np.int64)]
    for q, r in enumerate(ranks):
        r = int(r)
        if r not in received:
            raise ValueError(f"rank {rank}: no contribution received from neighbour {r}")
        shared = np.asarray(idx[ptr[q]:ptr[q + 1]], dtype=np.int64)
        pi, pj, v = received[r]
        if pi.size and (int(pi.max()) >= shared.size or int(pj.max()) >= shared.size):
            raise ValueError(f"rank {rank}: neighbour {r} addresses positions outside the shared interface")
        rows.append(shared[pi]); cols.append(shared[pj]); vals.append(np.asarray(v).reshape(-1, bb))
        src.append(np.full(pi.size, r, np.int64))
    rows, cols, vals, src = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), np.concatenate(src)
    order = np.lexsort((src, cols, rows))            # by entry, contributions of an entry in ascending rank order
    rows, cols, vals = rows[order], cols[order], vals[order]
    first = np.ones(rows.size, bool)
    first[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    start = np.flatnonzero(first)
    count = np.diff(np.append(start, rows.size))
    acc = vals[start].copy()
    for k in range(1, int(count.max()) if count.size else 0):
        m = count > k
        acc[m] = acc[m] + vals[start[m] + k]         # left to right: ((c0 + c1) + c2) + ...
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, rows[start] + 1, 1)
    return pr.Crs(n, A.ncols, A.block, np.cumsum(rp), cols[start].astype(np.int32), acc.ravel().copy())


def make_consistent(A: pr.Crs, rank: int, ranks, ptr, idx, dist) -> pr.Crs:
    """MakeConsistent(A) over torch.distributed (host data, init time only): what ugcore's parallel
    Gauss-Seidel does in preprocess (gauss_seidel.h:134-142).  Collective: every rank of the
    process group must call it for the same level, also ranks without neighbours."""
    mine = consistent_contributions(A, ranks, ptr, idx)
    everything = [None] * dist.get_world_size()
    dist.all_gather_object(everything, mine)
    received = {int(r): everything[int(r)][rank] for r in ranks}
    return apply_contributions(A, rank, ranks, ptr, idx, received)


def _global_base(part, base_mult=1, base=None):
    """Base-grid elements per direction of the GLOBAL grid: ``base`` if given (strong scaling: the same grid for every
    process grid that divides it), else base_mult elements per rank and direction (weak scaling)."""
    if base is not None:
        base = tuple(int(b) for b in base)
        if any(b % p for b, p in zip(base, part)):
            raise ValueError(f"process grid {tuple(part)} does not divide the base grid {base}")
        return base
    return tuple(base_mult * p for p in part)


def local_problem(refs: int, part, rank: int, problem=pr.POISSON, order=pr.ORDER_LEX, dim=3, base_mult: int = 1,
                  base=None, **kw) -> pr.Problem:
    """This rank's sub-box of the global grid whose base grid has base_mult elements per rank and
    direction (base_mult = 3, refs = 5: 97 nodes per direction and rank) — or ``base`` elements per direction
    in total (base = (2, 2, 2), refs = 7: the 257^3 grid of BASELINE configs[2] for every process grid up to 2x2x2)."""
    coord = rank_to_coord(rank, part)
    return pr.Problem(dim=dim, num_refs=refs, problem=problem, base=_global_base(part, base_mult, base), base_lev=0,
                      order=order, part=tuple(part), coord=coord, **kw)


def global_problem(refs: int, part, problem=pr.POISSON, dim=3, base_mult: int = 1, base=None, **kw) -> pr.Problem:
    """The same grid assembled serially (parity target and gathered base matrix)."""
    return pr.Problem(dim=dim, num_refs=refs, problem=problem, base=_global_base(part, base_mult, base), base_lev=0, **kw)


_nccl_ready = False
_p2p_state = None  # None: not tried, True / False: outcome (identical on all ranks)


def nccl_bootstrap(dist) -> None:
    """Rank 0 draws the NCCL unique id, torch.distributed broadcasts it, every rank joins."""
    import ctypes as C

    import torch

    from .capi import check, check_host, dev, host
    from .solver import host_init
    global _nccl_ready
    host_init()
    if _nccl_ready:
        return
    _nccl_ready = True
    ident = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        check(dev.ug4b200_comm_unique_id(ident))
    t = torch.tensor(list(ident), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    ident = (C.c_ubyte * 128)(*t.cpu().tolist())
    check_host(host.ug4b200_host_comm_init(dist.get_world_size(), dist.get_rank(), ident))


def p2p_bootstrap(dist) -> bool:
    """Open the peer windows: every rank allocates its window, torch.distributed carries the
    64-byte CUDA IPC handles, every rank maps the windows of all others.  All ranks agree on the
    outcome; on failure (or UG4B200_P2P=0) the NCCL transport stays in use."""
    import ctypes as C
    import os

    import torch

    from .capi import dev
    from .solver import host_ctx
    global _p2p_state
    if _p2p_state is not None:
        return _p2p_state
    world, rank = dist.get_world_size(), dist.get_rank()
    cuda = dist.get_backend() == "nccl"
    ctx = host_ctx()

    def agree(ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32)
        if cuda:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    want = os.environ.get("UG4B200_P2P", "1") != "0" and world > 1
    handle = (C.c_ubyte * 64)()
    ok = want and dev.ug4b200_p2p_window_create(ctx, 0, handle, None) == 0
    created = ok
    mine = torch.tensor(list(handle), dtype=torch.uint8)
    if cuda:
        mine = mine.cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    ok = agree(ok)
    if ok:
        buf = (C.c_ubyte * (64 * world))(*[int(v) for t in allh for v in t.cpu().tolist()])
        rc = dev.ug4b200_p2p_window_open(ctx, world, rank, buf)
        if rc != 0 and rank == 0:
            msg = dev.ug4b200_last_error(ctx)
            print(f"ugcore_b200: peer windows unavailable ({msg.decode() if msg else rc}); using NCCL send/recv", flush=True)
        ok = agree(rc == 0)
    if not ok and created:
        dev.ug4b200_p2p_window_destroy(ctx)
    _p2p_state = ok
    return ok


def default_gather_level(refs: int, part, base: int = 0, max_local_rows: int = 40000, max_global_rows: int = 300000,
                         dim: int = 3, base_mult: int = 1, global_base=None) -> int:
    """Highest level that is kept (and cycled) redundantly on every rank instead of being partitioned:
    a level whose LOCAL box has at most max_local_rows rows is latency-bound — its four interface
    exchanges per cycle (~6-8 us each) cost more than smoothing the whole (still small) global level
    on every rank.  Measured at 129^3 per GPU: gathering level 5 (33^3 local, 65^3 global at N = 8)
    instead of level 4 saves 0.15-0.25 ms per solve at N = 2 and N = 8."""
    gb = _global_base(part, base_mult, global_base)
    lev = base
    for l in range(base, refs):
        loc, glob = 1, 1
        for d in range(dim):
            loc *= (gb[d] // part[d]) * 2 ** l + 1
            glob *= gb[d] * 2 ** l + 1
        if loc <= max_local_rows and glob <= max_global_rows:
            lev = l
    return lev


GS_KINDS = ("gs", "bgs", "sgs", "ilu")   # smoothers that sweep over the CONSISTENT level matrix in partitioned runs


def _smoother_kind(desc: dict):
    """type of the smoother that will run on partitioned levels (GMG smoother, or the preconditioner itself)"""
    pc = desc["precond"] if "precond" in desc else "ilu"
    if isinstance(pc, str):
        pc = None if pc == "none" else {"type": pc}
    if not pc:
        return None
    if pc.get("type") == "gmg":
        sm = pc.get("smoother", "gs")
        return sm if isinstance(sm, str) else sm.get("type")
    return pc.get("type")


def build_partitioned_solver(desc: dict, refs: int, part, rank: int, dist, problem=pr.POISSON, flags: int = 0,
                             gather_level=None, **kw):
    """Wire a GMG-preconditioned solver on this rank's sub-box: local additive level matrices and
    interface layouts for the partitioned levels gather_level+1..refs; the levels base..gather_level
    are gathered: every rank holds the global matrices and runs that part of the V-cycle
    redundantly (gather_level = base: only the base solve is gathered, as in
    mg_solver_impl.hpp:2003-2070).  gather_level None: UG4B200_GATHER_LEVEL or the default rule."""
    import os

    from .solver import Solver
    nccl_bootstrap(dist)
    p2p_bootstrap(dist)
    prob = local_problem(refs, part, rank, problem=problem, **kw)
    desc = dict(desc)
    pc = desc["precond"] if "precond" in desc else "ilu"
    if isinstance(pc, str):
        pc = None if pc == "none" else {"type": pc}
    if not pc or pc.get("type") != "gmg":
        # one-level preconditioner (Jacobi / Gauss-Seidel / ILU — util.solver's default is ILU) or none: the
        # Krylov vectors and the preconditioner live on the top level's layouts
        s = Solver(desc, prob.matrix(refs), None, flags)
        s._keep.append(prob)
        ranks, ptr, idx = interfaces(prob, refs)
        s.set_layouts(refs, ranks, ptr, idx, prob.matrix(refs).nrows)
        if _smoother_kind(desc) in GS_KINDS:
            s.set_smoother_matrix(refs, make_consistent(prob.matrix(refs), rank, ranks, ptr, idx, dist))
        return prob, s
    pc = dict(pc)
    pc["topLevel"], pc["baseLevel"] = refs, pc.get("baseLevel", 0)
    base = pc["baseLevel"]
    if gather_level is None:
        env = os.environ.get("UG4B200_GATHER_LEVEL")
        gather_level = int(env) if env is not None else default_gather_level(refs, part, base, dim=prob.dim,
                                                                             base_mult=kw.get("base_mult", 1),
                                                                             global_base=kw.get("base"))
    if pc.get("cycle", "V") != "V":
        gather_level = base
    gather = max(base, min(int(gather_level), refs - 1))
    pc["gatherLevel"] = gather
    desc["precond"] = pc
    levels = {}
    for lev in range(gather, refs + 1):
        levels[lev] = (prob.matrix(lev), None if lev == gather else prob.prolongation(lev),
                       None if lev == gather else prob.restriction(lev))
    s = Solver(desc, prob.matrix(refs), levels, flags)
    s._keep.append(prob)
    gs = _smoother_kind(desc) in GS_KINDS
    for lev in range(gather, refs + 1):
        ranks, ptr, idx = interfaces(prob, lev)
        s.set_layouts(lev, ranks, ptr, idx, prob.matrix(lev).nrows)
        if gs and lev > gather:
            # ugcore's parallel Gauss-Seidel smooths with the level matrix made consistent on the interface
            # rows (gauss_seidel.h:134-142); the rows are exchanged on the host, once
            s.set_smoother_matrix(lev, make_consistent(prob.matrix(lev), rank, ranks, ptr, idx, dist))
    gprob = global_problem(gather, part, problem=problem, **kw)
    s._keep.append(gprob)
    s.set_gathered_base(gprob.matrix(gather), prob.global_ids(gather).astype(np.int32))
    for lev in range(base, gather + 1):
        if gather == base:
            break
        s.set_gathered_level(lev, None if lev == gather else gprob.matrix(lev),
                             None if lev == base else gprob.prolongation(lev),
                             None if lev == base else gprob.restriction(lev), nrows=gprob.matrix(lev).nrows)
    return prob, s
