"""GPU parity of the peer-window transport (comm.cu), on ONE GPU: R contexts of this process
play the ranks ("loop-back": ug4b200_p2p_window_attach hands every context the raw base
pointers of the others' windows instead of CUDA IPC handles; the kernels are the ones the
multi-GPU path runs, on R streams of the same device).

Checked bit for bit against the definition: every copy of an interface DoF ends up with the
sum over all copies in ascending rank order (AdditiveToConsistent,
ugbase/lib_algebra/parallelization/parallelization_util.h:159-191); all-reduce = sum over
ranks in ascending rank order, identical on every rank.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PART = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


class Ranks:
    def __init__(self, world, window_bytes=8 << 20):
        from ugcore_b200 import capi
        self.capi, self.dev, self.world = capi, capi.dev, world
        self.ctx, bases = [], (C.c_void_p * world)()
        for r in range(world):
            c = C.c_void_p()
            capi.check(self.dev.ug4b200_ctx_create(0, None, C.byref(c)))
            self.ctx.append(c)
            b = C.c_void_p()
            capi.check(self.dev.ug4b200_p2p_window_create(c, window_bytes, None, C.byref(b)), c)
            bases[r] = b.value
        for r in range(world):
            capi.check(self.dev.ug4b200_p2p_window_attach(self.ctx[r], world, r, bases), self.ctx[r])
            assert self.dev.ug4b200_p2p_enabled(self.ctx[r]) == 1
        self.bufs = []

    def chk(self, r, rc):
        return self.capi.check(rc, self.ctx[r])

    def up(self, r, a, dtype=np.float64):
        a = np.ascontiguousarray(a, dtype=dtype)
        p = C.c_void_p()
        self.chk(r, self.dev.ug4b200_alloc(self.ctx[r], max(a.nbytes, 8), C.byref(p)))
        self.chk(r, self.dev.ug4b200_h2d(self.ctx[r], p, a.ctypes.data_as(C.c_void_p), a.nbytes))
        self.chk(r, self.dev.ug4b200_sync(self.ctx[r]))
        self.bufs.append((r, p))
        return p

    def down(self, r, p, n):
        a = np.empty(n)
        self.chk(r, self.dev.ug4b200_d2h(self.ctx[r], a.ctypes.data_as(C.c_void_p), p, a.nbytes))
        return a

    def sync_all(self):
        for r in range(self.world):
            self.chk(r, self.dev.ug4b200_sync(self.ctx[r]))

    def close(self, ifaces=()):
        for r, I in ifaces:
            self.dev.ug4b200_interface_destroy(self.ctx[r], I)
        for r, p in self.bufs:
            self.dev.ug4b200_free(self.ctx[r], p)
        for c in self.ctx:
            self.dev.ug4b200_ctx_destroy(c)


def _interfaces(R, refs, lev):
    """Per rank: (problem, interface handle) of one level of the box partition."""
    from ugcore_b200 import dist as ugdist
    part = PART[R.world]
    probs, handles = [], []
    for r in range(R.world):
        prob = ugdist.local_problem(refs, part, r)
        ranks, ptr, idx = ugdist.interfaces(prob, lev)
        I = C.c_void_p()
        R.chk(r, R.dev.ug4b200_interface_create(R.ctx[r], ranks.size, ranks.ctypes.data_as(C.c_void_p),
                                                ptr.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
                                                prob.matrix(lev).nrows, C.byref(I)))
        probs.append(prob)
        handles.append(I)
    for r in range(R.world):  # all published: resolve the neighbours' receive regions
        R.chk(r, R.dev.ug4b200_interface_commit(R.ctx[r], handles[r]))
    return probs, handles


def _expected_consistent(vals, gids, nglobal, block):
    """ascending-rank sum of all copies, per global DoF component"""
    acc = np.zeros(nglobal * block)
    seen = np.zeros(nglobal * block, bool)
    for v, g in zip(vals, gids):
        gi = (g[:, None] * block + np.arange(block)[None, :]).ravel()
        acc[gi] = np.where(seen[gi], acc[gi] + v, v)
        seen[gi] = True
    return [acc[(g[:, None] * block + np.arange(block)[None, :]).ravel()] for g in gids]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("block", [1, 3])
def test_additive_to_consistent_loopback_bit_exact(world, block):
    refs, lev = 4, 4  # 17^3 boxes: faces of 289 DoFs, edges, corners
    R = Ranks(world)
    ifaces = []
    try:
        probs, H = _interfaces(R, refs, lev)
        ifaces = list(enumerate(H))
        gids = [p.global_ids(lev) for p in probs]
        nglobal = int(max(g.max() for g in gids)) + 1
        rng = np.random.default_rng(world * 10 + block)
        # several exchanges back to back without a host sync in between: exercises the epoch
        # parity double buffering (a rank may run one epoch ahead of a neighbour)
        rounds = 5
        vals = [[rng.standard_normal(g.size * block) for g in gids] for _ in range(rounds)]
        dv = [[R.up(r, vals[k][r]) for r in range(world)] for k in range(rounds)]
        for k in range(rounds):
            for r in range(world):
                R.chk(r, R.dev.ug4b200_additive_to_consistent(R.ctx[r], H[r], dv[k][r], block))
        R.sync_all()
        for k in range(rounds):
            exp = _expected_consistent(vals[k], gids, nglobal, block)
            for r in range(world):
                got = R.down(r, dv[k][r], gids[r].size * block)
                assert np.array_equal(got, exp[r]), (k, r)
    finally:
        R.close(ifaces)


def test_exchange_multi_cta_and_graph_replay():
    """33^3 boxes (faces of 1089 DoFs x block 3 -> several CTAs per exchange) and the exchange
    captured into a CUDA graph that is replayed: epochs live on the device."""
    world, refs, lev, block = 2, 5, 5, 3
    R = Ranks(world)
    ifaces = []
    try:
        probs, H = _interfaces(R, refs, lev)
        ifaces = list(enumerate(H))
        gids = [p.global_ids(lev) for p in probs]
        nglobal = int(max(g.max() for g in gids)) + 1
        rng = np.random.default_rng(3)
        v0 = [rng.standard_normal(g.size * block) for g in gids]
        dv = [R.up(r, v0[r]) for r in range(world)]
        graphs = []
        for r in range(world):
            R.chk(r, R.dev.ug4b200_graph_begin(R.ctx[r]))
            R.chk(r, R.dev.ug4b200_additive_to_consistent(R.ctx[r], H[r], dv[r], block))
            g = C.c_void_p()
            R.chk(r, R.dev.ug4b200_graph_end(R.ctx[r], C.byref(g)))
            graphs.append(g)
        reps = 3
        for _ in range(reps):
            for r in range(world):
                R.chk(r, R.dev.ug4b200_graph_launch(R.ctx[r], graphs[r]))
        R.sync_all()
        exp = v0
        for _ in range(reps):  # consistent -> "additive" again: every replay sums the copies once more
            exp = _expected_consistent(exp, gids, nglobal, block)
        for r in range(world):
            assert np.array_equal(R.down(r, dv[r], gids[r].size * block), exp[r])
        for r in range(world):
            R.dev.ug4b200_graph_destroy(R.ctx[r], graphs[r])
    finally:
        R.close(ifaces)


@pytest.mark.parametrize("world", [2, 8])
def test_allreduce_and_fused_dot_loopback(world):
    R = Ranks(world)
    try:
        rng = np.random.default_rng(world)
        n = 37
        x = [rng.standard_normal(n) for _ in range(world)]
        dx = [R.up(r, x[r]) for r in range(world)]
        for rep in range(3):  # epochs advance; result of rep k feeds rep k+1
            for r in range(world):
                R.chk(r, R.dev.ug4b200_allreduce_sum(R.ctx[r], dx[r], n))
        R.sync_all()
        exp = x
        for rep in range(3):
            s = exp[0].copy()
            for r in range(1, world):
                s = s + exp[r]
            exp = [s.copy() for _ in range(world)]
        for r in range(world):
            assert np.array_equal(R.down(r, dx[r], n), exp[r])
        # fused local dot + all-reduce + finaliser (sqrt) in one kernel per rank
        m = 100_003
        a = [rng.standard_normal(m) for _ in range(world)]
        b = [rng.standard_normal(m) for _ in range(world)]
        da = [R.up(r, a[r]) for r in range(world)]
        db = [R.up(r, b[r]) for r in range(world)]
        out = [R.up(r, np.zeros(2)) for r in range(world)]
        for r in range(world):
            fin = R.capi.Fin(R.capi.FIN_STORE, out[r], None, None, None)
            R.chk(r, R.dev.ug4b200_vec_dot_allreduce_ds(R.ctx[r], m, da[r], db[r], fin, None))
        R.sync_all()
        got = [R.down(r, out[r], 1)[0] for r in range(world)]
        ref = sum(float(np.dot(a[r], b[r])) for r in range(world))
        assert all(g == got[0] for g in got), "all ranks must hold the identical sum"
        assert abs(got[0] - ref) <= 1e-12 * max(1.0, sum(float(np.abs(a[r] * b[r]).sum()) for r in range(world)))
    finally:
        R.close()


@pytest.mark.parametrize("world,block", [(2, 1), (8, 1), (4, 3)])
def test_gather_sum_loopback_bit_exact(world, block):
    """Gathered level: local additive vectors -> global sum on every rank (ascending rank order)."""
    from ugcore_b200 import dist as ugdist
    refs = lev = 3
    R = Ranks(world)
    G = []
    try:
        part = PART[world]
        probs = [ugdist.local_problem(refs, part, r) for r in range(world)]
        gids = [p.global_ids(lev).astype(np.int32) for p in probs]
        nglobal = int(max(g.max() for g in gids)) + 1
        for r in range(world):
            h = C.c_void_p()
            R.chk(r, R.dev.ug4b200_gather_create(R.ctx[r], nglobal, gids[r].size, gids[r].ctypes.data_as(C.c_void_p), block, C.byref(h)))
            G.append(h)
        for r in range(world):
            R.chk(r, R.dev.ug4b200_gather_commit(R.ctx[r], G[r]))
        rng = np.random.default_rng(7 + world)
        rounds = 4
        vals = [[rng.standard_normal(g.size * block) for g in gids] for _ in range(rounds)]
        dl = [[R.up(r, vals[k][r]) for r in range(world)] for k in range(rounds)]
        dg = [[R.up(r, np.full(nglobal * block, np.nan)) for r in range(world)] for k in range(rounds)]
        for k in range(rounds):
            for r in range(world):
                R.chk(r, R.dev.ug4b200_gather_sum(R.ctx[r], G[r], dg[k][r], dl[k][r]))
        R.sync_all()
        for k in range(rounds):
            acc = None
            for r in range(world):
                slot = np.zeros(nglobal * block)
                gi = (gids[r][:, None].astype(np.int64) * block + np.arange(block)[None, :]).ravel()
                slot[gi] = vals[k][r]
                acc = slot if acc is None else acc + slot
            for r in range(world):
                assert np.array_equal(R.down(r, dg[k][r], nglobal * block), acc), (k, r)
    finally:
        for r, h in enumerate(G):
            R.dev.ug4b200_gather_destroy(R.ctx[r], h)
        R.close()
