"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): partitioned solve vs serial oracle."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# p2p = 1: peer-window transport (direct NVLink stores + flags), 0: NCCL send/recv + all-reduce
# gather: levels 0..gather are held by every rank and cycled redundantly (-1: default rule = 2 here)
# p2p = 2: peer windows + interface rows pushed by the smoothing kernels themselves (UG4B200_FUSED_PUSH=1)
@pytest.mark.parametrize("world,flags,p2p,gather", [(2, 0, 1, -1), (2, 2, 1, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 0, 0),
                                                    (2, 0, 2, 0), (2, 2, 2, -1),
                                                    (4, 0, 1, -1), (8, 0, 1, -1), (8, 0, 1, 0), (8, 0, 0, 1), (8, 0, 2, 0)])
def test_partitioned_gmg_cg_matches_serial_oracle(world, flags, p2p, gather):
    _run(world, flags, p2p, gather, "poisson", 1e-10, 1e-9)


@pytest.mark.parametrize("world,flags,p2p,gather", [(2, 0, 1, -1), (2, 0, 0, 0), (4, 0, 1, 1), (8, 0, 1, -1)])
def test_partitioned_bicgstab_gmg_gauss_seidel_matches_oracle(world, flags, p2p, gather):
    """BASELINE configs[3] partitioned: BiCGStab + GMG with ugcore's parallel Gauss-Seidel (multicolour inside
    a rank) vs the serial oracle of the same method."""
    _run(world, flags, p2p, gather, "convdiff_gs", 1e-10, 1e-7)


@pytest.mark.parametrize("world,flags,p2p,gather", [(2, 0, 1, -1), (4, 0, 0, 0)])
def test_partitioned_bicgstab_gmg_ilu_matches_oracle(world, flags, p2p, gather):
    """The same with ILU(0) smoothing in the multicolour ordering (parallel ILU of ilu.h:536-543, 640-652)."""
    _run(world, flags, p2p, gather, "convdiff_ilu", 1e-10, 1e-7)


@pytest.mark.parametrize("world,p2p,case,tol", [(2, 1, "cg_ilu", 1e-10), (4, 0, "cg_ilu", 1e-10), (2, 1, "bicgstab_gs", 1e-10),
                                                (8, 1, "bicgstab_gs", 1e-10)])
def test_partitioned_one_level_preconditioners_match_oracle(world, p2p, case, tol):
    """No multigrid: CG + ILU(0) (util.solver's default solver) and BiCGStab + Gauss-Seidel on a partitioned grid,
    ugcore's parallel mode of both (consistent matrix, Dirichlet rows on the h-slaves, unique defect)."""
    _run(world, 0, p2p, -1, case, tol, 1e-7)


@pytest.mark.parametrize("world,flags,p2p,gather", [(2, 0, 1, -1), (8, 0, 1, -1), (2, 0, 0, 0)])
def test_partitioned_elasticity_block3_matches_serial_oracle(world, flags, p2p, gather):
    """BASELINE configs[4] partitioned: 3x3-block GMG-CG, block interface exchange."""
    _run(world, flags, p2p, gather, "elasticity", 1e-10, 1e-9)


@pytest.mark.parametrize("world,case,tol", [(2, "poisson_sgs", 1e-10), (4, "elasticity_sgs", 1e-10)])
def test_partitioned_symmetric_gauss_seidel_matches_oracle(world, case, tol):
    """CG + GMG with symmetric Gauss-Seidel smoothing (scalar and 3x3 blocks) in ugcore's parallel mode."""
    _run(world, 0, 1, -1, case, tol, 1e-8)


@pytest.mark.parametrize("world,cycle", [(2, "W"), (4, "F")])
def test_partitioned_w_and_f_cycles_match_serial_oracle(world, cycle):
    """W- and F-cycles visit the coarse levels several times: everything below the top level stays partitioned down
    to the gathered base solve (no replicated coarse cycle)."""
    _run(world, 0, 1, -1, "poisson", 1e-10, 1e-9, extra=[cycle])


@pytest.mark.parametrize("world,p2p,case,tol", [(2, 1, "poisson", 1e-10), (2, 0, "poisson", 1e-10), (4, 1, "poisson", 1e-10),
                                                (8, 1, "poisson", 1e-10), (2, 1, "elasticity", 1e-10), (8, 1, "elasticity", 1e-10),
                                                (2, 1, "convdiff_gs", 1e-10), (4, 1, "poisson_sgs", 1e-10)])
def test_partitioned_solves_with_random_rhs_match_serial_oracle(world, p2p, case, tol):
    """Seeded random global right-hand side: values on the two sides of every partition plane are unrelated and O(1)
    on the interfaces, so summation order, master / slave choice and the 4- and 8-way sharers all show up in the
    history (parallel_vector_impl.h:269-379, parallelization_util.h:159-280)."""
    _run(world, 0, p2p, -1, case, tol, 1e-7 if tol > 1e-9 else 1e-9, seed=20261017)


def _run(world, flags, p2p, gather, case, hist_tol, sol_tol, extra=(), seed=None):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, UG4B200_P2P=str(min(p2p, 1)), UG4B200_FUSED_PUSH="1" if p2p == 2 else "0")
    env.pop("UG4B200_TEST_RHS_SEED", None)
    if seed is not None:
        env["UG4B200_TEST_RHS_SEED"] = str(seed)
    env.pop("UG4B200_GATHER_LEVEL", None)
    if gather >= 0:
        env["UG4B200_GATHER_LEVEL"] = str(gather)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + flags + 20 * p2p + 40 * (gather + 1)),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), "3", str(flags), case] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert line, r.stdout[-3000:] + r.stderr[-3000:]
    for res in json.loads(line[0][len("MGPU_RESULT "):]):
        assert res["ok"] and res["oracle_ok"]
        assert abs(res["its"] - res["its_oracle"]) <= 1
        # north_star's tolerance, or 10 x the movement of the reference's own history under a reordered sum
        assert res["hist_err"] < max(hist_tol, 10 * res["ref_reorder_sensitivity"]) and res["sol_err"] < sol_tol, res
        assert res["p2p"] == bool(p2p), res
        # the interfaces must carry real data (round 1's Poisson rhs was antisymmetric about every partition plane)
        assert res["iface_rel"] > 1e-3, res
