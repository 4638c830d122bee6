"""torchrun entry of the multi-GPU parity test: partitioned GMG-CG on N GPUs vs the SERIAL
oracle on the same global grid (same math up to summation order)."""
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
    from helpers import gmg_desc, oracle_levels, rel_hist_err
    from ugcore_b200 import dist as ugdist, solver as S

    refs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    S.host_init(lr, None)
    part = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    desc = gmg_desc(refs)
    prob, s = ugdist.build_partitioned_solver(desc, refs, part, rank, dist, flags=flags)
    x, ok, h = s.apply(prob.rhs())

    gprob = ugdist.global_problem(refs, part)
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    lv = oracle_levels(orc, gprob)
    xo, oko, ho = oracle.OSolver(orc, desc, lv[refs][0], lv).apply(gprob.rhs())
    gid = prob.global_ids(refs)
    from ugcore_b200 import capi
    res = {"rank": rank, "p2p": bool(capi.dev.ug4b200_p2p_enabled(S.host_ctx())), "ok": bool(ok), "oracle_ok": bool(oko), "its": len(h) - 1, "its_oracle": len(ho) - 1,
           "hist_err": rel_hist_err(h, ho), "sol_err": float(np.linalg.norm(x - xo[gid]) / np.linalg.norm(xo[gid]))}
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
