"""torchrun entry of the multi-GPU parity tests: a partitioned solve on N GPUs vs the SERIAL oracle
on the same global grid.

  argv: refs flags [case [cycle]]
  case "poisson" (default)  GMG V(2,2) Jacobi + CG on 3-D Poisson: same math as the serial solver up to
                            summation order -> compared with the plain serial oracle
  case "elasticity"         the same with 3x3 blocks (block Jacobi, interface exchange of blocks)
  case "convdiff_gs"        BiCGStab + GMG with Gauss-Seidel smoothing on upwind convection-diffusion:
                            ugcore's parallel Gauss-Seidel is a block Jacobi over the ranks with GS inside
                            (gauss_seidel.h:134-142, 204-215) -> compared with the serial oracle of exactly
                            that method (helpers.partitioned_gs_oracle)
  case "convdiff_ilu"       the same with ILU(0) smoothing (ilu.h:536-543, 640-652: same parallel structure)
  case "poisson_sgs", "elasticity_sgs"   CG + GMG with symmetric Gauss-Seidel smoothing, scalar and 3x3 blocks
  case "cg_ilu"             no multigrid: CG + ILU(0) (util.solver's default), natural ordering inside a rank
  case "bicgstab_gs"        no multigrid: BiCGStab + one (multicolour) Gauss-Seidel sweep
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle
    from helpers import gmg_desc, oracle_levels, partitioned_gs_oracle, partitioned_onelevel_oracle, rel_hist_err
    from ugcore_b200 import dist as ugdist, problems as pr, solver as S

    refs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    case = sys.argv[3] if len(sys.argv) > 3 else "poisson"
    cycle = sys.argv[4] if len(sys.argv) > 4 else "V"          # "V" | "W" | "F" (poisson / elasticity cases)
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    S.host_init(lr, None)
    part = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    block = 1
    if case == "poisson":
        problem, kw, desc = pr.POISSON, {}, gmg_desc(refs, cycle=cycle)
    elif case == "elasticity":
        problem, kw, desc, block = pr.ELASTICITY, {}, gmg_desc(refs, reduction=1e-8, its=200), 3
    elif case == "convdiff_gs":
        problem, kw = pr.CONVDIFF, {"eps": 1e-1}
        desc = gmg_desc(refs, solver="bicgstab", smoother={"type": "gs", "relax": 1.0}, reduction=1e-8)
    elif case == "convdiff_ilu":   # ILU(0) smoothing, multicolour ordering inside a rank; parallel mode of ilu.h:536-543, 640-652
        problem, kw = pr.CONVDIFF, {"eps": 1e-1}
        desc = gmg_desc(refs, solver="bicgstab", smoother={"type": "ilu", "ordering": "multicolor"}, reduction=1e-8)
    elif case == "poisson_sgs":    # CG + GMG with symmetric Gauss-Seidel smoothing (forward, diagonal, backward sweep)
        problem, kw, desc = pr.POISSON, {}, gmg_desc(refs, smoother={"type": "sgs", "relax": 1.0})
    elif case == "elasticity_sgs": # the same with 3x3 blocks
        problem, kw, desc, block = pr.ELASTICITY, {}, gmg_desc(refs, smoother={"type": "sgs", "relax": 1.0}, reduction=1e-8, its=200), 3
    elif case == "cg_ilu":         # util.solver's default configuration: CG preconditioned by ILU(0), one level
        problem, kw = pr.POISSON, {}
        desc = {"type": "cg", "precond": {"type": "ilu"}, "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}}
    elif case == "bicgstab_gs":    # BiCGStab preconditioned by one Gauss-Seidel sweep
        problem, kw = pr.CONVDIFF, {"eps": 1e-1}
        desc = {"type": "bicgstab", "precond": {"type": "gs"}, "convCheck": {"iterations": 200, "absolute": 1e-12, "reduction": 1e-8}}
    else:
        raise SystemExit(f"unknown case {case}")
    prob, s = ugdist.build_partitioned_solver(desc, refs, part, rank, dist, problem=problem, flags=flags, **kw)
    gid = prob.global_ids(refs)
    g = np.repeat(gid * block, block) + np.tile(np.arange(block), gid.size)
    seed = os.environ.get("UG4B200_TEST_RHS_SEED")
    if seed:
        # seeded random global right-hand side (0 on Dirichlet DoFs), handed to the ranks in UNIQUE form (the h-master
        # holds the value, the other copies 0 — a valid additive vector whose sum over the copies is exactly b_glob):
        # interface values are O(1) and unrelated on the two sides of every partition plane
        gtop = ugdist.global_problem(refs, part, problem=problem, **kw)
        b_glob = np.random.default_rng(int(seed)).standard_normal(gtop.num_dofs)
        b_glob[np.repeat(np.asarray(gtop.dirichlet(refs), bool), block)] = 0.0
        own = np.repeat(ugdist.owned_mask(prob, refs, rank), block)
        b_loc = np.where(own, b_glob[g], 0.0)
        rhs_of = lambda gp: b_glob
    else:
        b_loc = prob.rhs()
        rhs_of = lambda gp: np.array(gp.rhs())
    x, ok, h = s.apply(b_loc)

    if case in ("cg_ilu", "bicgstab_gs"):
        solve, gprob = partitioned_onelevel_oracle(orc, desc, refs, part, problem=problem, colored=(case == "bicgstab_gs"), **kw)
        osolve = lambda: solve(rhs_of(gprob))
    elif case in ("convdiff_gs", "convdiff_ilu", "poisson_sgs", "elasticity_sgs"):
        solve, gprob = partitioned_gs_oracle(orc, desc, refs, part, s.desc.gather_lev, problem=problem, **kw)
        osolve = lambda: solve(rhs_of(gprob))
    else:
        gprob = ugdist.global_problem(refs, part, problem=problem, **kw)
        lv = oracle_levels(orc, gprob)
        osol = oracle.OSolver(orc, desc, lv[refs][0], lv)
        osolve = lambda: osol.apply(rhs_of(gprob))
    xo, oko, ho = osolve()
    # how far the reference's own history moves when only the summation order of its reductions changes
    # (tests/test_reduction_order.py): the partitioned sum IS another summation order
    orc.set_reduction_mode(1)
    try:
        _, _, h1 = osolve()
    finally:
        orc.set_reduction_mode(0)
    k = min(len(ho), len(h1))
    sens = float(np.max(np.abs(ho[:k] - h1[:k]) / np.abs(ho[:k])))
    from ugcore_b200 import capi
    res = {"rank": rank, "p2p": bool(capi.dev.ug4b200_p2p_enabled(S.host_ctx())), "ok": bool(ok), "oracle_ok": bool(oko), "its": len(h) - 1, "its_oracle": len(ho) - 1,
           "hist_err": rel_hist_err(h, ho), "sol_err": float(np.linalg.norm(x - xo[g]) / np.linalg.norm(xo[g])),
           "iface_rel": float(np.max(np.abs(x[np.repeat(ugdist.multiplicity(prob, refs) > 1, block)])) / max(np.max(np.abs(x)), 1e-300)),
           "final_reduction": float(h[-1] / h[0]), "ref_reorder_sensitivity": sens}
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
