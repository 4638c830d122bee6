"""Generate tests/golden/residual_histories.json with the COMPILED REFERENCE kernels
(oracle/_ref, built from /root/reference/ugbase) driving the restated solver loops.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
                                                     python tests/golden/make_golden.py ilu   (ilu_gmres_histories.json)
The fixtures pin (i) the oracle port (CPU suite) and (ii) the CUDA path (GPU suite) on the
GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import gmg_desc, make_rhs, oracle_levels  # noqa: E402
from ugcore_b200 import problems as pr  # noqa: E402

CASES = [
    ("S1_poisson2d_33_gmg_jac_cg", dict(dim=2, num_refs=5), gmg_desc(5), None),
    ("S2_poisson3d_17_gmg_jac_cg", dict(dim=3, num_refs=4), gmg_desc(4), None),
    ("poisson3d_17_gmg_jac_cg_random_rhs", dict(dim=3, num_refs=4), gmg_desc(4), 11),
    ("S2_poisson3d_33_gmg_jac_cg", dict(dim=3, num_refs=5), gmg_desc(5), None),
    ("poisson3d_17_hier_gmg_jac_cg", dict(dim=3, num_refs=4, order=1), gmg_desc(4), 12),
    ("poisson3d_17_gmg_W_linear", dict(dim=3, num_refs=4), gmg_desc(4, solver="linear", cycle="W", reduction=1e-8), 13),
    ("S5_elasticity3d_9_gmg_blockjac_cg", dict(dim=3, num_refs=3, problem=2), gmg_desc(3, reduction=1e-8, its=200), None),
    ("poisson2d_17_cg_jacobi", dict(dim=2, num_refs=4),
     {"type": "cg", "precond": {"type": "jac", "damp": 0.66}, "convCheck": {"iterations": 400, "absolute": 1e-12, "reduction": 1e-8}}, 14),
]


_CC8 = {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}
# second fixture file: ILU(0) / ILU(beta) (the reference's own FactorizeILUSorted / FactorizeILUBeta / invert_L /
# invert_U, compiled from operator/preconditioner/ilu.h) under CG, BiCGStab, LinearSolver, GMRES and as GMG smoother
CASES_ILU = [
    ("poisson3d_17_cg_ilu", dict(dim=3, num_refs=4), {"type": "cg", "precond": {"type": "ilu"}, "convCheck": _CC8}, None),
    ("convdiff3d_17_bicgstab_ilu_beta03", dict(dim=3, num_refs=4, problem=1, eps=0.1),
     {"type": "bicgstab", "precond": {"type": "ilu", "beta": 0.3}, "convCheck": _CC8}, None),
    ("convdiff3d_17_gmres5_ilu", dict(dim=3, num_refs=4, problem=1, eps=0.1),
     {"type": "gmres", "restart": 5, "precond": {"type": "ilu"}, "convCheck": _CC8}, 21),
    ("convdiff3d_9_gmres20_noprecond", dict(dim=3, num_refs=3, problem=1, eps=0.1),
     {"type": "gmres", "restart": 20, "precond": None, "convCheck": {"iterations": 200, "absolute": 1e-12, "reduction": 1e-6}}, None),
    ("poisson2d_33_linear_ilu", dict(dim=2, num_refs=5),
     {"type": "linear", "precond": {"type": "ilu"}, "convCheck": {"iterations": 400, "absolute": 1e-12, "reduction": 1e-6}}, 22),
    ("elasticity3d_5_cg_block_ilu", dict(dim=3, num_refs=2, problem=2), {"type": "cg", "precond": {"type": "ilu"}, "convCheck": _CC8}, None),
    ("poisson3d_17_gmg_ilu_smoother_cg", dict(dim=3, num_refs=4), gmg_desc(4, smoother={"type": "ilu"}), 23),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "base"
    cases, fname = (CASES, "residual_histories.json") if which == "base" else (CASES_ILU, "ilu_gmres_histories.json")
    orc = oracle.Oracle("ref")
    out = {"generator": "tests/golden/make_golden.py" + ("" if which == "base" else " ilu"), "oracle_backend": orc.kind,
           "reference": "UG4/ugcore kernels compiled from /root/reference/ugbase (oracle/_ref)", "cases": []}
    for name, pargs, desc, seed in cases:
        prob = pr.Problem(**pargs)
        pc = desc.get("precond")
        if isinstance(pc, dict) and pc.get("type") == "gmg":
            lv = oracle_levels(orc, prob, pc["baseLevel"], pc["topLevel"])
            s = oracle.OSolver(orc, desc, lv[pc["topLevel"]][0], lv)
        else:
            s = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
        x, ok, h = s.apply(make_rhs(prob, seed))
        out["cases"].append({"name": name, "problem": pargs, "desc": desc, "rhs_seed": seed, "converged": bool(ok),
                             "history": [float(v) for v in h], "solution_norm": float(np.linalg.norm(x)),
                             "solution_sample": [float(v) for v in x[:: max(1, x.size // 16)][:16]]})
        print(name, ok, len(h) - 1, h[-1] / h[0])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), fname), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
