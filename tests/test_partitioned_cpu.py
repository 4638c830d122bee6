"""CPU: the partitioned CPU baseline arm (oracle/partitioned.py — ugcore's MPI path emulated by processes of this
host: additive matrices, consistent / additive / unique vectors, gathered coarse levels) against the SERIAL oracle on the
same global grid.  It is baseline infrastructure for `bench.py --impl reference --gpus N`; what is pinned here is that
it solves the same problem the same way: identical iteration counts, the residual history to round-off (a partitioned
sum is another summation order), the solution to 1e-12."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, oracle_levels
from oracle import partitioned
from ugcore_b200 import dist as ugdist, problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _serial(orc, refs, part, desc, problem=pr.POISSON, **kw):
    gp = ugdist.global_problem(refs, part, problem=problem, **kw)
    lv = oracle_levels(orc, gp)
    s = oracle.OSolver(orc, desc, lv[refs][0], lv)
    return s.apply(np.array(gp.rhs())), s.block


@pytest.mark.parametrize("part,refs,problem,kw", [
    ((2, 1, 1), 3, pr.POISSON, {}),
    ((2, 2, 1), 3, pr.POISSON, {}),
    ((2, 2, 2), 3, pr.POISSON, {}),
    ((2, 2, 1), 3, pr.POISSON, {"base": (2, 2, 2)}),          # strong-scaling grid: 2x2x2 base cells on a 2x2x1 process grid
    ((2, 2, 2), 2, pr.ELASTICITY, {}),                        # 3x3 blocks: block Jacobi with the consistent diagonal blocks
])
def test_partitioned_cpu_solve_equals_the_serial_oracle(part, refs, problem, kw):
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    desc = gmg_desc(refs) if problem == pr.POISSON else gmg_desc(refs, reduction=1e-8, its=200)
    res = partitioned.run(part, refs, desc, problem=problem, want_x=True, **kw)[0]
    (xo, oko, ho), B = _serial(orc, refs, part, desc, problem=problem, **kw)
    assert len(res) == int(np.prod(part)) and oko
    for r in res:
        h = np.array(r["hist"])
        assert r["ok"] and len(h) == len(ho)
        assert np.max(np.abs(h - ho) / ho) < 1e-11
        assert r["hist"] == res[0]["hist"]                     # every rank sees the same bits (rank-ordered all-reduce)
        g = (r["gid"][:, None] * B + np.arange(B)[None, :]).ravel()
        assert np.linalg.norm(r["x"] - xo[g]) <= 1e-12 * np.linalg.norm(xo[g])


def test_gather_level_rule_and_concurrent_jobs():
    """Levels whose GLOBAL grid has at most 5000 nodes are gathered; two jobs side by side use separate exchange files."""
    res = partitioned.run((2, 1, 1), 4, gmg_desc(4), jobs=2)
    assert len(res) == 2 and all(len(job) == 2 for job in res)
    assert res[0][0]["gather"] == 3                            # 17 x 9 x 9 = 1377 <= 5000 < 33 x 17 x 17
    assert res[0][0]["hist"] == res[1][0]["hist"]


def _comm_worker(rank, world, path, q):
    sys.path.insert(0, ROOT)
    from oracle.partitioned import ShmComm
    c = ShmComm(path, rank, world, 8, 4)
    # ring of interfaces: every rank shares DoFs 0..2 with the next and 3..5 with the previous rank
    ranks = sorted({(rank + 1) % world, (rank - 1) % world})
    segs = [np.arange(0, 3) if r == (rank + 1) % world else np.arange(3, 6) for r in ranks]
    if world == 2:
        segs = [np.arange(0, 6)]
    out = []
    for it in range(50):                                       # mailboxes are reused: sequence counters must hold
        v = np.full(6, float(rank + 1 + it))
        c.exchange_add(v, ranks, segs)
        out.append(v.copy())
        s = c.allsum(float(rank) + 0.5 * it)
        assert s == sum(float(r) + 0.5 * it for r in range(world))
    buf = np.full(4, float(rank + 1))
    tot = c.sum_to_root(buf)
    if rank == 0:
        tot[:] = 2.0 * tot
    res = np.array(c.bcast_from_root())
    c.barrier()
    q.put((rank, out[-1], res))


@pytest.mark.parametrize("world", [2, 3])
def test_shared_memory_message_layer(world, tmp_path):
    import multiprocessing as mp
    path = str(tmp_path / "exchange")
    with open(path, "wb") as f:
        f.truncate(8 * partitioned.ShmComm.words(world, 8, 4))
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    ps = [mpc.Process(target=_comm_worker, args=(r, world, path, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    [p.join(timeout=30) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    it = 49
    for rank, v, res in got:
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        if world == 2:
            assert np.array_equal(v, np.full(6, (rank + 1 + it) + (nxt + 1 + it)))
        else:
            # each side lists ITS copies of the shared DoFs: my 0..2 pair with the next rank's 3..5
            assert np.array_equal(v[:3], np.full(3, (rank + 1 + it) + (nxt + 1 + it)))
            assert np.array_equal(v[3:], np.full(3, (rank + 1 + it) + (prv + 1 + it)))
        assert np.array_equal(res, np.full(4, 2.0 * sum(range(1, world + 1))))


def test_bench_reference_arm_partitions_the_multi_gpu_grid():
    """`bench.py --impl reference --gpus 2`: the global grid of the 2-GPU run (33 x 17 x 17 at numRefs 4), partitioned;
    `--cpu-arm replicas` keeps round 1's exchange-free replicas of one GPU's box."""
    def run(extra):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--refs", "4", "--steps", "1",
                            "--warmup", "0", "--cpu-procs", "2"] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        return json.loads(lines[0])
    d = run([])
    assert d["config"]["cpu_arm"] == "partitioned" and d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 2
    assert "33x17x17 nodes (9537 DoF" in d["cpu_baseline"]["sample"] and d["config"]["iterations"] == 7
    assert d["value"] == pytest.approx(9537 / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    e = run(["--cpu-arm", "replicas"])
    assert e["config"]["cpu_arm"] == "replicas" and "17^3 nodes (4913 DoF" in e["cpu_baseline"]["sample"]
