"""CPU suite, world_size 2 over gloo: host-side logic of the partitioned path — interface
lists, additive/consistent/unique protocol, gathered-base map — checked against the serial
assembly of the same global grid.  (The device kernels need a GPU; their exchange is the same
protocol over NCCL, covered by tests/test_multi_gpu.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ugcore_b200 import dist as ugdist, problems as pr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part, refs = (2, 1, 1), 3
        prob = ugdist.local_problem(refs, part, rank)
        gprob = ugdist.global_problem(refs, part)
        for lev in range(0, refs + 1):
            A, gA = prob.matrix(lev).to_scipy(), gprob.matrix(lev).to_scipy()
            gid = prob.global_ids(lev)
            ranks, ptr, idx = ugdist.interfaces(prob, lev)
            assert list(ranks) == [1 - rank]
            # both sides list the same global DoFs in the same order
            mine = torch.from_numpy(gid[idx].copy())
            theirs = torch.empty_like(mine)
            reqs = [dist.isend(mine, 1 - rank), dist.irecv(theirs, 1 - rank)]
            [r.wait() for r in reqs]
            assert torch.equal(mine, theirs)
            nodes = prob.dims(lev)
            assert idx.size == nodes[1] * nodes[2]

            def to_consistent(v):
                """AdditiveToConsistent, summed in ascending rank order (comm.cu semantics)."""
                send = torch.from_numpy(v[idx].copy()); recv = torch.empty_like(send)
                reqs = [dist.isend(send, 1 - rank), dist.irecv(recv, 1 - rank)]
                [r.wait() for r in reqs]
                out = v.copy()
                out[idx] = (v[idx] + recv.numpy()) if rank == 0 else (recv.numpy() + v[idx])
                return out

            # A_additive * x_consistent made consistent == serial A x (away from Dirichlet rows,
            # which are identity on every copy: additive sum = multiplicity * identity)
            rng = np.random.default_rng(lev)
            xg = rng.standard_normal(gA.shape[0])
            y = to_consistent(A @ xg[gid])
            interior = prob.dirichlet(lev) == 0
            assert np.allclose(y[interior], (gA @ xg)[gid][interior], rtol=1e-13, atol=1e-13)
            mult = ugdist.multiplicity(prob, lev)
            assert np.allclose(y[~interior], (mult * xg[gid])[~interior])
            # unique dot: sum over h-master copies of a consistent vector == serial dot
            own = ugdist.owned_mask(prob, lev, rank)
            t = torch.tensor([float(np.dot(xg[gid][own], xg[gid][own]))], dtype=torch.float64)
            dist.all_reduce(t)
            assert abs(t.item() - float(xg @ xg)) < 1e-10 * float(xg @ xg)
            # additive rhs sums to the serial rhs
            if lev == refs:
                b = to_consistent(np.array(prob.rhs()))
                assert np.allclose(b, np.array(gprob.rhs())[gid], rtol=1e-13, atol=1e-15)
            # P is consistent (same rows as the serial P), R additive sums to serial R
            if lev > 0:
                P, gP = prob.prolongation(lev).to_scipy(), gprob.prolongation(lev).to_scipy()
                cg = prob.global_ids(lev - 1)
                xc = rng.standard_normal(gP.shape[1])
                assert np.allclose(P @ xc[cg], (gP @ xc)[gid], rtol=1e-14, atol=1e-14)
        # MakeConsistent over the real transport (all_gather_object): interior rows of the consistent
        # matrix are the rows of the serially assembled matrix restricted to the local DoFs
        for lev in range(1, refs + 1):
            ranks, ptr, idx = ugdist.interfaces(prob, lev)
            Ac = ugdist.make_consistent(prob.matrix(lev), rank, ranks, ptr, idx, dist)
            gid = prob.global_ids(lev)
            G = gprob.matrix(lev).to_scipy()[gid][:, gid].toarray()
            interior = prob.dirichlet(lev) == 0
            assert np.array_equal(Ac.to_scipy().toarray()[interior], G[interior])
        # gathered base: local level-0 DoFs map into the global base numbering
        l2g = prob.global_ids(0)
        assert l2g.size == 8 and set(l2g) <= set(range(12))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_partition_protocol_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=240) for _ in ps]
    [p.join(timeout=60) for p in ps]
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_interfaces_2x2x2_symmetry():
    """All 8 ranks in one process: every pair agrees on its shared DoFs; corner DoF has 8 copies."""
    sys.path.insert(0, ROOT)
    from ugcore_b200 import dist as ugdist
    part, refs = (2, 2, 2), 2
    probs = [ugdist.local_problem(refs, part, r) for r in range(8)]
    lists = {}
    for r, p in enumerate(probs):
        ranks, ptr, idx = ugdist.interfaces(p, refs)
        assert len(ranks) == 7
        gid = p.global_ids(refs)
        for k, s in enumerate(ranks):
            lists[(r, int(s))] = gid[idx[ptr[k]:ptr[k + 1]]]
        mult = ugdist.multiplicity(p, refs)
        assert mult.max() == 8 and (mult == 8).sum() == 1
    for (r, s), g in lists.items():
        assert np.array_equal(g, lists[(s, r)])
    owned = sum(int(ugdist.owned_mask(p, refs, r).sum()) for r, p in enumerate(probs))
    assert owned == (2 * 2 ** refs + 1) ** 3


@pytest.mark.parametrize("part", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_strong_scaling_partitions_of_one_global_grid(part):
    """bench.py --scaling strong: the SAME global grid (2x2x2 base cells) on every process grid.  All ranks in one process:
    the local boxes tile the global grid (owned DoFs sum to its size), the additive right-hand sides sum to the serial
    one, and — the point of the round-2 generator change — the solution is O(1) on the partition planes."""
    sys.path.insert(0, ROOT)
    from ugcore_b200 import dist as ugdist
    refs, base = 3, (2, 2, 2)
    world = part[0] * part[1] * part[2]
    gprob = ugdist.global_problem(refs, part, base=base)
    assert gprob.dims(refs) == (17, 17, 17)
    probs = [ugdist.local_problem(refs, part, r, base=base) for r in range(world)]
    acc = np.zeros(gprob.num_dofs)
    owned = 0
    for r, p in enumerate(probs):
        gid = p.global_ids(refs)
        np.add.at(acc, gid, np.array(p.rhs()))
        owned += int(ugdist.owned_mask(p, refs, r).sum())
        assert p.dims(refs) == tuple(16 // part[d] + 1 for d in range(3))
    assert owned == 17 ** 3
    assert np.allclose(acc, np.array(gprob.rhs()), rtol=1e-13, atol=1e-18)
    ex = np.array(gprob.exact()).reshape(17, 17, 17)            # [k][j][i]
    assert np.abs(ex[:, :, 8]).max() > 0.4 * np.abs(ex).max()   # the plane x = 1 between the two boxes in x
    with pytest.raises(ValueError):
        ugdist.local_problem(refs, (3, 1, 1), 0, base=base)     # 3 does not divide 2 base cells
    # the gather rule sees the real local box: 129^3 per rank at refs 7 on 2x2x2 -> level 5, one rank with 257^3 -> level 4
    assert ugdist.default_gather_level(7, (2, 2, 2), global_base=base) == 5
    assert ugdist.default_gather_level(7, (2, 1, 1), global_base=base) <= 5


def _comp(idx, b):
    idx = np.asarray(idx)
    return np.repeat(idx * b, b) + np.tile(np.arange(b), idx.size)


@pytest.mark.parametrize("problem", [0, 1, 2])
@pytest.mark.parametrize("part", [(2, 1, 1), (2, 2, 2)])
def test_make_consistent_and_parallel_gs_model(problem, part):
    """Host logic of the partitioned Gauss-Seidel smoother, all ranks emulated in one process:
    (1) dist.consistent_contributions / apply_contributions (ugcore: MakeConsistent) reproduce the rows of
        the serially assembled matrix; Dirichlet rows carry the multiplicity like in ugcore;
    (2) the global model the multi-GPU parity test solves with the serial oracle (helpers.
        parallel_gs_global_model) equals the rank-by-rank sweep of gauss_seidel.h:134-142, 204-215 —
        bit for bit for scalar problems."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import greedy_color_perm, parallel_gs_global_model, permute_crs, with_dirichlet_rows
    from ugcore_b200 import dist as ugdist
    orc = oracle.Oracle("port")
    refs = lev = 2
    world = part[0] * part[1] * part[2]
    locs = [ugdist.local_problem(refs, part, r, problem=problem) for r in range(world)]
    gprob = ugdist.global_problem(refs, part, problem=problem)
    b = gprob.block
    gA = gprob.matrix(lev)
    G = gA.to_scipy()
    ifs = [ugdist.interfaces(p, lev) for p in locs]
    sent = [ugdist.consistent_contributions(p.matrix(lev), *ifs[r]) for r, p in enumerate(locs)]
    cons = []
    for r, p in enumerate(locs):
        ranks, ptr, idx = ifs[r]
        A = p.matrix(lev)
        Ac = ugdist.apply_contributions(A, r, ranks, ptr, idx, {int(s): sent[int(s)][r] for s in ranks})
        assert np.array_equal(Ac.rowptr, A.rowptr) and np.array_equal(Ac.cols, A.cols)
        g = _comp(p.global_ids(lev), b)
        S, Gl = Ac.to_scipy().toarray(), G[g][:, g].toarray()
        dirn = np.repeat(p.dirichlet(lev) != 0, b)
        assert np.allclose(S[~dirn], Gl[~dirn], rtol=1e-13, atol=1e-15)
        mult = np.repeat(ugdist.multiplicity(p, lev), b)
        assert np.array_equal(S[dirn], (np.eye(S.shape[0]) * mult)[dirn])
        cons.append(Ac)
    gperm, keep = parallel_gs_global_model(locs, gprob, lev)
    rng = np.random.default_rng(1)
    d = rng.standard_normal(gA.nrows * b)
    d[np.repeat(gprob.dirichlet(lev) != 0, b)] = 0.0     # a defect vanishes in Dirichlet rows
    gp = _comp(gperm, b)
    S = orc.matrix(permute_crs(gA, gperm, gperm, keep=keep))
    for kind in ("ll", "ur", "sgs"):
        dp = np.empty_like(d); dp[gp] = d
        c_model = S.gs(dp, kind, 0.9)[gp]
        c_ranks = np.zeros_like(d)
        for r, p in enumerate(locs):
            own = ugdist.owned_mask(p, lev, r)
            Ad = with_dirichlet_rows(cons[r], np.flatnonzero(~own))            # SetDirichletRow on the h-slaves
            perm, _ = greedy_color_perm(Ad)
            gid = p.global_ids(lev)
            du = np.where(np.repeat(own, b), d[_comp(gid, b)], 0.0)           # unique defect
            pp = _comp(perm, b)
            dl = np.empty_like(du); dl[pp] = du
            c = orc.matrix(permute_crs(Ad, perm, perm)).gs(dl, kind, 0.9)[pp]
            assert not c[np.repeat(~own, b)].any()                            # the correction is unique
            c_ranks[_comp(gid[own], b)] = c[np.repeat(own, b)]
        if b == 1:
            assert np.array_equal(c_model, c_ranks), kind
        else:   # block values are not dyadic: the rank-ordered sums of the interface rows round differently
            assert np.allclose(c_model, c_ranks, rtol=1e-12, atol=1e-14), kind


def test_partitioned_gs_oracle_converges():
    """The serial oracle of the partitioned BiCGStab + GMG(GS) solve (what tests/test_multi_gpu.py compares
    the GPUs with) converges like the serial solver, a little slower (block Jacobi over the ranks)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import gmg_desc, partitioned_gs_oracle
    orc = oracle.Oracle("port")
    desc = gmg_desc(3, solver="bicgstab", smoother={"type": "gs", "relax": 1.0}, reduction=1e-8)
    solve, gprob = partitioned_gs_oracle(orc, desc, 3, (2, 2, 1), 1, problem=1, eps=1e-1)
    b = np.array(gprob.rhs())
    x, ok, h = solve(b)
    assert ok and h[-1] < 1e-8 * h[0] and len(h) < 20
    A = gprob.matrix(3).to_scipy()
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)
