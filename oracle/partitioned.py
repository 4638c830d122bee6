"""oracle/partitioned.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The CPU arm of BASELINE.md §2 item 5: what P MPI ranks of ugcore do on the hot path, emulated by P processes of this
host.  Every process owns one sub-box of the global grid with its ADDITIVE level matrices and runs GMG V(nu1,nu2)
(Jacobi) preconditioned CG with ugcore's storage-type protocol:

  * vectors are consistent / additive / unique exactly where ugcore's are (ParallelVector,
    lib_algebra/parallelization/parallel_vector_impl.h:115-379; AdditiveToConsistent / AdditiveToUnique,
    parallelization_util.h:159-280): the defect stays additive, corrections are made consistent, `norm()` turns the
    defect unique in place and all-reduces the squares, `dotprod(consistent, additive)` all-reduces the local products;
  * the parallel Jacobi uses the CONSISTENT diagonal and makes its correction consistent (jacobi.h:170-232);
  * transfers are local (additive defects restrict to additive defects, consistent corrections prolongate to
    consistent ones: std_transfer_impl.h:719-806);
  * the levels base..gather are gathered onto one process, solved there (serial V-cycle of the oracle, LU on the base
    level) and the correction is broadcast — the gathered base solver of mg_solver_impl.hpp:2003-2070.

All matrix kernels, dot products and norms are the compiled reference kernels (oracle/_ref) or the port, called in place
on numpy memory.  The interface exchange and the all-reduce go through ONE shared-memory file mapped by all ranks
(ShmComm: a mailbox per ordered pair of ranks with sequence counters, spin waits) — the stand-in for pcl over
shared-memory MPI, which this image does not have; gloo was measured first and rejected (0.4 ms per all-reduce at 2
ranks, 4 ms at 8 on this host — two orders of magnitude above an MPI shared-memory transport, which would have made the
baseline unfairly slow).  The control flow restates cg.h:103-242 and mg_solver_impl.hpp:174-275, 1685-2136 like
oracle/solvers.cpp does for one process.

Checked against the serial oracle on the same global grid by tests/test_partitioned_cpu.py (same iteration count,
history to round-off).  Used by `bench.py --impl reference --gpus N` (N > 1): the same global grid the N GPUs solve,
partitioned over CPU processes.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ShmComm:
    """Message layer of one partitioned job over a file in /dev/shm mapped by every rank.

    Layout (float64 words unless noted): int64 counters seq[src, dst], ack[src, dst], red_seq[rank], bar[rank],
    red[2, rank], a mailbox of ``cap`` doubles per ordered pair, and ``world`` + 1 areas of ``nglob`` doubles for the
    gathered base solve.  A sender waits until its previous message to that neighbour was consumed, writes, bumps seq;
    the receiver waits for seq, reads, bumps ack.  x86 keeps stores (and loads) in program order, so data written before
    the counter is visible once the counter is."""

    def __init__(self, path, rank, world, cap, nglob):
        self.rank, self.world, self.cap, self.nglob = rank, world, cap, nglob
        ni = 2 * world * world + 2 * world
        import mmap
        with open(path, "r+b") as f:
            self._mm = mmap.mmap(f.fileno(), 0)            # MAP_SHARED
        m = self.mem = np.frombuffer(self._mm, dtype=np.int64)   # plain ndarray views: np.memmap's indexing is 3x slower
        self.seq = m[0:world * world].reshape(world, world)
        self.ack = m[world * world:2 * world * world].reshape(world, world)
        self.red_seq = m[2 * world * world:2 * world * world + world]
        self.bar = m[2 * world * world + world:ni]
        f = m[ni:].view(np.float64)
        self.red = f[0:2 * world].reshape(2, world)
        o = 2 * world
        self.box = f[o:o + world * world * cap].reshape(world, world, cap)
        o += world * world * cap
        self.garea = f[o:o + (world + 1) * nglob].reshape(world + 1, nglob)
        self.nred = 0
        self.nbar = 0

    @staticmethod
    def words(world, cap, nglob):
        return 2 * world * world + 2 * world + 2 * world + world * world * cap + (world + 1) * nglob

    @staticmethod
    def _wait(arr, i, want):
        spins = 0
        while arr[i] < want:
            spins += 1
            if spins > 2000:            # oversubscribed host: let the rank we wait for run
                time.sleep(0)

    def exchange_add(self, v, ranks, segs):
        """AdditiveToConsistent (parallelization_util.h:159-200): every copy of an interface DoF becomes the sum of all
        copies.  One message per neighbour and direction."""
        me = self.rank
        for q, r in enumerate(ranks):
            self._wait(self.ack[me], r, self.seq[me, r])
            self.box[me, r, :segs[q].size] = v[segs[q]]
            self.seq[me, r] += 1
        for q, r in enumerate(ranks):
            self._wait(self.seq[r], me, self.ack[r, me] + 1)
            v[segs[q]] += self.box[r, me, :segs[q].size]
            self.ack[r, me] += 1

    def allsum(self, s):
        """Sum of one double over all ranks, added in rank order on every rank (all ranks get the same bits)."""
        self.nred += 1
        slot = self.nred & 1
        self.red[slot, self.rank] = s
        self.red_seq[self.rank] = self.nred
        for r in range(self.world):
            self._wait(self.red_seq, r, self.nred)
        t = 0.0
        for r in range(self.world):
            t += float(self.red[slot, r])
        return t

    def barrier(self):
        self.nbar += 1
        self.bar[self.rank] = self.nbar
        for r in range(self.world):
            self._wait(self.bar, r, self.nbar)

    def sum_to_root(self, buf):
        """Additive vectors of all ranks summed (rank order) into area[world] on rank 0."""
        self.garea[self.rank, :] = buf
        self.barrier()
        if self.rank == 0:
            np.copyto(self.garea[self.world], self.garea[0])
            for r in range(1, self.world):
                self.garea[self.world] += self.garea[r]
        return self.garea[self.world]

    def bcast_from_root(self):
        """After rank 0 has written the result into area[world]."""
        self.barrier()
        return self.garea[self.world]


class _Level:
    __slots__ = ("A", "P", "R", "n", "block", "dinv", "ranks", "segs", "own", "sc", "sd", "st")


class PartitionedCG:
    """One rank of the partitioned GMG-CG solve.  ``comm_of(cap, nglob)`` returns this rank's ShmComm once the
    mailbox capacity and the size of the gathered level are known (identical on all ranks)."""

    def __init__(self, orc, comm_of, prob, gprob_of, global_base, rank, world, desc, gather=None):
        from ugcore_b200 import dist as ugdist
        self.orc, self.L = orc, orc.lib
        self.rank, self.world = rank, world
        pc = desc["precond"]
        sm = pc.get("smoother", {"type": "jac", "damp": 0.66})
        if desc.get("type", "cg") != "cg" or pc.get("type") != "gmg" or sm.get("type") not in ("jac", "jacobi"):
            raise ValueError("partitioned CPU arm: CG + GMG with Jacobi smoothing only")
        if pc.get("cycle", "V") != "V":
            raise ValueError("partitioned CPU arm: V-cycle only")
        self.nu1, self.nu2 = int(pc.get("preSmooth", 2)), int(pc.get("postSmooth", 2))
        damp = float(sm.get("damp", 0.66))
        cc = desc.get("convCheck", {})
        self.max_steps = int(cc.get("iterations", 100))
        self.min_defect, self.reduction = float(cc.get("absolute", 1e-12)), float(cc.get("reduction", 1e-10))
        self.top, self.base = prob.num_refs, int(pc.get("baseLevel", 0))
        if gather is None:
            # the CPU pays ~0.1 ms per exchange and smooths a 17^3 level in about the same time: gather below that
            gather = self.base
            for l in range(self.base, self.top):
                if np.prod([global_base[d] * 2 ** l + 1 for d in range(prob.dim)]) <= 5000:
                    gather = l
        self.gather = max(self.base, min(int(gather), self.top - 1))
        self.block = B = prob.matrix(self.top).block
        gp = gprob_of(self.gather)
        self.nglob = gp.matrix(self.gather).nrows * B
        # mailbox capacity: the largest face of any rank's box on the top level, B*B values per DoF (the diagonal
        # blocks travel once); every rank computes the same number
        dims = [global_base[d] // prob.part[d] * 2 ** self.top + 1 for d in range(3)]
        if prob.dim == 2:
            dims[2] = 1
        dims.sort()
        self.comm = comm_of(int(dims[1] * dims[2] * B * B), self.nglob)
        self.lev = {}
        for l in range(self.gather, self.top + 1):
            lv = _Level()
            crs = prob.matrix(l)
            lv.A = orc.matrix(crs)
            lv.P = orc.matrix(prob.prolongation(l)) if l > self.gather else None
            lv.R = orc.matrix(prob.restriction(l)) if l > self.gather else None
            lv.n, lv.block = crs.nrows, B
            ranks, ptr, idx = ugdist.interfaces(prob, l)
            lv.ranks = [int(r) for r in ranks]
            lv.segs = []
            for q in range(len(lv.ranks)):
                s = np.asarray(idx[ptr[q]:ptr[q + 1]], dtype=np.int64)
                lv.segs.append((s[:, None] * B + np.arange(B)[None, :]).ravel() if B > 1 else s)
            lv.own = np.repeat(ugdist.owned_mask(prob, l, rank), B).astype(np.float64)
            lv.sc, lv.sd, lv.st = (np.zeros(lv.n * B) for _ in range(3))
            # consistent diagonal, inverted and damped once (jacobi.h:170-194)
            rows = np.repeat(np.arange(crs.nrows), np.diff(crs.rowptr))
            dsel = np.asarray(crs.cols) == rows
            diag = np.zeros((crs.nrows, B * B))
            diag[rows[dsel]] = np.asarray(crs.vals).reshape(-1, B * B)[dsel]
            flat = diag.ravel()
            # blocks travel as B*B values per DoF: exchange with block-sized segments
            segs_bb = [(np.asarray(idx[ptr[q]:ptr[q + 1]], dtype=np.int64)[:, None] * B * B + np.arange(B * B)[None, :]).ravel()
                       for q in range(len(lv.ranks))]
            self.comm.exchange_add(flat, lv.ranks, segs_bb)
            if B == 1:
                lv.dinv = damp / flat
            else:
                # FixedArray2 blocks are stored column-major (common.cuh / small_algebra): (r, c) at r + B c
                lv.dinv = damp * np.linalg.inv(flat.reshape(-1, B, B).transpose(0, 2, 1))
            self.lev[l] = lv
        # gathered part: global ids of this rank's DoFs on the gather level; rank 0 holds the serial hierarchy
        gid = np.asarray(prob.global_ids(self.gather), dtype=np.int64)
        self.gid = (gid[:, None] * B + np.arange(B)[None, :]).ravel() if B > 1 else gid
        self.gbuf = np.zeros(self.nglob)
        self.gsolver = None
        if rank == 0:
            import oracle
            if self.gather > self.base:
                lvs = {}
                for l in range(self.base, self.gather + 1):
                    lvs[l] = (orc.matrix(gp.matrix(l)), orc.matrix(gp.prolongation(l)) if l > self.base else None,
                              orc.matrix(gp.restriction(l)) if l > self.base else None)
                gpc = dict(pc)
                gpc["topLevel"], gpc["baseLevel"] = self.gather, self.base
                self.gsolver = oracle.OSolver(orc, dict(desc, precond=gpc), lvs[self.gather][0], lvs)
                self._gkeep = lvs
            else:
                self.gmat = orc.matrix(gp.matrix(self.base))
        self._gp = gp
        n = self.lev[self.top].n * B
        self.q, self.z, self.p = np.zeros(n), np.zeros(n), np.zeros(n)
        self.history = []

    # ---- communication -------------------------------------------------------------------------------------------
    def _a2c(self, lv, v):
        self.comm.exchange_add(v, lv.ranks, lv.segs)

    def _allsum(self, s):
        return self.comm.allsum(s)

    def _norm_unique(self, lv, r):
        """ParallelVector::norm of an additive vector (parallel_vector_impl.h:300-330): r becomes UNIQUE in place
        (the h-master copy holds the sum, the others 0), then sqrt(allreduce(local sum of squares))."""
        self._a2c(lv, r)
        r *= lv.own
        nl = self.L.oracle_norm(r.size // lv.block, lv.block, r)
        return float(np.sqrt(self._allsum(nl * nl)))

    def _dot(self, lv, a, b):
        return self._allsum(self.L.oracle_dot(a.size // lv.block, lv.block, a, b))

    # ---- smoother ------------------------------------------------------------------------------------------------
    def _jacobi(self, lv, c, d):
        if lv.block == 1:
            np.multiply(d, lv.dinv, out=c)
        else:
            B = lv.block
            c[:] = np.einsum("nij,nj->ni", lv.dinv, d.reshape(-1, B)).ravel()
        self._a2c(lv, c)                                   # c additive -> consistent (jacobi.h:226-230)

    # ---- multigrid cycle (mg_solver_impl.hpp:1685-2136) ----------------------------------------------------------------
    def _lmgc(self, l):
        L = self.L
        lf = self.lev[l]
        B = lf.block
        for nu in range(self.nu1):
            self._jacobi(lf, lf.st, lf.sd)
            L.oracle_matmul_minus(lf.A.h, lf.sd, lf.st, B)
            if nu < self.nu1 - 1:
                lf.sc += lf.st
        if self.nu1 > 0:
            lf.sc += lf.st
        lc = self.lev[l - 1]
        lc.sc[:] = 0.0
        L.oracle_apply_ignore_zero_rows(lf.R.h, lc.sd, 1.0, lf.sd, B)
        if l - 1 == self.gather:
            self._gathered_solve()
        else:
            self._lmgc(l - 1)
        L.oracle_axpy(lf.P.h, lf.st, 0.0, lf.st.ctypes.data, 1.0, lc.sc, B)
        lf.sc += lf.st
        for nu in range(self.nu2):
            L.oracle_matmul_minus(lf.A.h, lf.sd, lf.st, B)
            self._jacobi(lf, lf.st, lf.sd)
            lf.sc += lf.st
        if l >= self.top:
            L.oracle_matmul_minus(lf.A.h, lf.sd, lf.st, B)

    def _gathered_solve(self):
        """Gathered base solver (mg_solver_impl.hpp:2003-2070): additive defects are summed onto rank 0, the serial
        cycle below runs there, the consistent correction goes back to everyone."""
        g = self.lev[self.gather]
        buf = self.gbuf
        buf[:] = 0.0
        buf[self.gid] = g.sd
        tot = self.comm.sum_to_root(buf)
        if self.rank == 0:
            tot[:] = self.gsolver.precond_apply(tot) if self.gsolver is not None else self.gmat.lu_solve(tot)
        res = self.comm.bcast_from_root()
        g.sc[:] = res[self.gid]
        self.comm.barrier()                                # the areas are rewritten by the next gathered solve

    def precond(self, c, d):
        """AssembledMultiGridCycle::apply (mg_solver_impl.hpp:174-275): c consistent, d additive (untouched)."""
        t = self.lev[self.top]
        t.sd[:] = d
        t.sc[:] = 0.0
        self._lmgc(self.top)
        c[:] = t.sc

    # ---- CG (cg.h:103-242) ---------------------------------------------------------------------------------------
    def apply(self, b_additive):
        """x = 0; returns (x consistent, converged, defect history)."""
        L, t = self.L, self.lev[self.top]
        B = t.block
        n = t.n
        x = np.zeros(n * B)
        r = np.array(b_additive, dtype=np.float64)        # r = b - A*0
        q, z, p = self.q, self.z, self.p
        self.precond(z, r)
        h = [self._norm_unique(t, r)]
        p[:] = z
        rho_old = self._dot(t, z, r)
        ended = lambda: (not np.isfinite(h[-1])) or len(h) - 1 >= self.max_steps or h[-1] < self.min_defect or h[-1] / h[0] < self.reduction
        while not ended():
            L.oracle_apply(t.A.h, q, p, B)
            lam = self._dot(t, q, p)
            if lam == 0.0:
                break
            alpha = rho_old / lam
            L.oracle_scale_add2(x.size, x, 1.0, x, alpha, p)
            L.oracle_scale_add2(r.size, r, 1.0, r, -alpha, q)
            h.append(self._norm_unique(t, r))
            if ended():
                break
            self.precond(z, r)
            rho = self._dot(t, z, r)
            beta = rho / rho_old
            L.oracle_scale_add2(p.size, p, beta, p, 1.0, z)
            rho_old = rho
        self.history = h
        ok = h[-1] < self.min_defect or h[-1] / h[0] < self.reduction
        return x, bool(ok), np.array(h)


def _worker(rank, part, refs, desc, store, steps, warmup, q, problem, kw, want_x):
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, ROOT)
    import oracle
    from ugcore_b200 import dist as ugdist
    world = int(np.prod(part))
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    prob = ugdist.local_problem(refs, part, rank, problem=problem, **kw)
    gprob_of = lambda l: ugdist.global_problem(l, part, problem=problem, **kw)     # serial grid up to the gather level
    gbase = ugdist._global_base(part, kw.get("base_mult", 1), kw.get("base"))

    def comm_of(cap, nglob):
        # rank 0 sizes the file (sparse, zero-filled = all counters 0), the others wait for it to have its full size
        nbytes = 8 * ShmComm.words(world, cap, nglob)
        if rank == 0:
            with open(store + ".tmp", "wb") as f:
                f.truncate(nbytes)
            os.rename(store + ".tmp", store)
        else:
            t0 = time.time()
            while not os.path.exists(store):
                time.sleep(0.01)
                if time.time() - t0 > 600:
                    raise RuntimeError("partitioned CPU arm: rank 0 never created the exchange file")
        return ShmComm(store, rank, world, cap, nglob)

    s = PartitionedCG(orc, comm_of, prob, gprob_of, gbase, rank, world, desc)
    b = np.array(prob.rhs())
    for _ in range(warmup):
        s.apply(b)
    s.comm.barrier()
    t0 = time.perf_counter()
    x = ok = h = None
    for _ in range(steps):
        x, ok, h = s.apply(b)
    s.comm.barrier()
    dt = time.perf_counter() - t0
    out = {"rank": rank, "dt": dt, "ok": ok, "hist": [float(v) for v in h], "n_local": int(prob.matrix(refs).nrows) * s.block,
           "gather": s.gather, "kind": orc.kind}
    if want_x:
        out["x"] = np.array(x, copy=True)
        out["gid"] = np.array(prob.global_ids(refs), copy=True)   # the queue pickles later, in a feeder thread
    q.put(out)


def run(part, refs, desc, steps=1, warmup=0, jobs=1, problem=0, want_x=False, timeout=3600, **kw):
    """``jobs`` concurrent partitioned solves (each prod(part) processes sharing one exchange file) of the global grid
    local_problem(refs, part, ...) describes.  Returns a list (one entry per job) of per-rank result lists."""
    import multiprocessing as mp
    import shutil
    import tempfile
    mpc = mp.get_context("spawn")
    world = int(np.prod(part))
    groups = []
    # exchange files: /dev/shm when it has room (a container may cap it at 64 MB: a sparse file there would end in SIGBUS),
    # else an ordinary temporary directory (MAP_SHARED on the page cache behaves the same)
    shm = None
    try:
        st = os.statvfs("/dev/shm")
        if os.access("/dev/shm", os.W_OK) and st.f_bavail * st.f_frsize > (2 << 30):
            shm = "/dev/shm"
    except OSError:
        pass
    tmp = tempfile.mkdtemp(prefix="ug4b200_part_", dir=shm)
    out = []
    try:
        for j in range(jobs):
            store = os.path.join(tmp, f"exchange{j}")
            qu = mpc.Queue()
            procs = [mpc.Process(target=_worker, args=(r, tuple(part), refs, desc, store, steps, warmup, qu, problem, kw, want_x))
                     for r in range(world)]
            for p in procs:
                p.start()
            groups.append((procs, qu))
        for procs, qu in groups:
            res = []
            deadline = time.time() + timeout
            while len(res) < world:
                try:
                    res.append(qu.get(timeout=2))
                except Exception:
                    if any(p.exitcode not in (None, 0) for p in procs):
                        raise RuntimeError("partitioned CPU arm: a rank died")
                    if time.time() > deadline:
                        raise RuntimeError("partitioned CPU arm: timeout")
            out.append(sorted(res, key=lambda r: r["rank"]))
    finally:
        for procs, _ in groups:
            for p in procs:
                p.join(timeout=5 if not out else 30)
                if p.is_alive():
                    p.terminate()       # our own children, by handle (a rank spinning on a dead neighbour)
        shutil.rmtree(tmp, ignore_errors=True)
    return out
