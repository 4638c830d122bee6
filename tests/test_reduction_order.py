"""CPU: how far does the REFERENCE's own residual history move when nothing but the summation order of its dot
products and norms changes?

ugcore sums strictly left to right (vector_impl.h:72-79, 323-329); a GPU reduction is a tree.  north_star asks for
histories "within 1e-10 relative per iteration" — whether a given solver can be held to that by ANY implementation
with another (equally valid) summation order is a property of the reference algorithm, and it is measured here
with the compiled reference kernels: the oracle's solver loop (oracle/solvers.cpp) runs once with ugcore's
sequential reductions and once with a pairwise tree over the same products (oracle.set_reduction_mode(1)),
everything else bit-identical.  The measured movement is the yardstick for the tolerances in the GPU tests
(tests/test_gpu_solver.py: hist_tol) and is committed as tests/golden/reduction_order_sensitivity.json.

Findings (see the JSON): CG + GMG moves by <= 2e-11 per step up to 65^3, 3.7e-11 at 129^3 and 1.04e-10 at 257^3 (the
error of a sequential sum grows with its length) -> 1e-10 holds with margin up to configs[1] and is exactly the
reference's own noise level at configs[2].  BiCGStab + GMG(Jacobi) moves by 4e-10 at 33^3 and 1.5e-8 at 65^3 on
Poisson: the reference itself cannot hold 1e-10 under a reordered sum there, so histories are compared with
max(1e-10, 10 x measured movement of the reference on the same problem) (helpers.sens_tol)."""
import json
import os

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, oracle_levels, reduction_order_sensitivity
from ugcore_b200 import problems as pr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reduction_order_sensitivity.json")


def _cases(refs):
    return {
        "cg_gmg_jacobi_poisson": (pr.Problem(dim=3, num_refs=refs), gmg_desc(refs)),
        "bicgstab_gmg_gs_convdiff": (pr.Problem(dim=3, num_refs=refs, problem=pr.CONVDIFF, eps=0.1),
                                     gmg_desc(refs, solver="bicgstab", smoother={"type": "gs", "relax": 1.0}, reduction=1e-8)),
        "bicgstab_gmg_jacobi_poisson": (pr.Problem(dim=3, num_refs=refs), gmg_desc(refs, solver="bicgstab")),
        "linear_gmg_jacobi_poisson": (pr.Problem(dim=3, num_refs=refs), gmg_desc(refs, solver="linear", reduction=1e-8)),
    }


def _measure(refs_list=(3, 4, 5)):
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    out = {}
    for refs in refs_list:
        for name, (prob, desc) in _cases(refs).items():
            sens, h0, h1 = reduction_order_sensitivity(orc, prob, desc, np.array(prob.rhs()))
            out[f"{name}@{2 ** refs + 1}^3"] = {"steps": len(h0) - 1, "steps_reordered": len(h1) - 1, "max_rel_move": float(sens.max()),
                                               "per_step": [float(v) for v in sens], "backend": orc.kind}
    return out


def test_pairwise_mode_only_changes_reductions():
    """mode 1 must leave every non-reduction operation alone: one V-cycle (no dot products inside) is bit-identical,
    and the mode is reset to the reference's sequential sums afterwards."""
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    prob = pr.Problem(dim=3, num_refs=3)
    desc = gmg_desc(3)
    lv = oracle_levels(orc, prob, 0, 3)
    s = oracle.OSolver(orc, desc, lv[3][0], lv)
    d = np.array(prob.rhs())
    c0 = s.precond_apply(d)
    orc.set_reduction_mode(1)
    try:
        c1 = s.precond_apply(d)
        x1, ok1, h1 = s.apply(d)
    finally:
        orc.set_reduction_mode(0)
    x0, ok0, h0 = s.apply(d)
    assert np.array_equal(c0, c1)
    assert ok0 and ok1 and len(h0) == len(h1)
    assert not np.array_equal(h0, h1), "a pairwise tree that reproduces the sequential sum bit for bit is no test"
    assert orc.lib.oracle_reduction_mode() == 0


def test_reference_history_sensitivity_matches_the_committed_measurement():
    """The numbers DESIGN.md quotes are the ones the compiled reference produces here (they are deterministic)."""
    got = _measure((3, 4))
    with open(GOLDEN) as f:
        want = json.load(f)
    for k, v in got.items():
        assert k in want, k
        if v["backend"] == want[k]["backend"]:
            assert v["steps"] == want[k]["steps"]
            assert v["max_rel_move"] == pytest.approx(want[k]["max_rel_move"], rel=1e-6), k


def test_cg_holds_1e10_and_bicgstab_does_not_under_reordered_sums():
    with open(GOLDEN) as f:
        want = json.load(f)
    cg = [v["max_rel_move"] for k, v in want.items() if k.startswith("cg_") and "257^3" not in k]
    bi = [v["max_rel_move"] for k, v in want.items() if k.startswith("bicgstab_gmg_jacobi")]
    assert max(cg) < 5e-11          # CG up to 129^3: 1e-10 per iteration is attainable with margin (257^3: see below)
    assert max(bi) > 1e-10          # BiCGStab: the reference itself moves by more than north_star's tolerance


def test_full_size_entries_explain_the_gpu_deviation():
    """The two full-size entries (129^3: 16 s, 257^3: 136 s and 20 GB on the CPU — generated once by
    `python tests/test_reduction_order.py --full`, not recomputed here): the reference's own CG history moves by 3.66e-11
    at 129^3 and 1.04e-10 at 257^3 when only its summation order changes — the very numbers by which the B200 history
    differs from it (bench.py history_rel_err_vs_cpu: 3.67e-11 / 1.04e-10, profiles/r02h, r02f): the deviation is the
    rounding of ugcore's sequential sums, not of the device path."""
    with open(GOLDEN) as f:
        want = json.load(f)
    a, b = want["cg_gmg_jacobi_poisson@129^3"], want["cg_gmg_jacobi_poisson@257^3 (2x2x2 base cells)"]
    assert a["steps"] == 8 and b["steps"] == 8
    assert 3e-11 < a["max_rel_move"] < 5e-11 and 0.9e-10 < b["max_rel_move"] < 1.2e-10


if __name__ == "__main__":   # regenerate the committed measurement: python tests/test_reduction_order.py [--full]
    import sys
    full = {}
    if "--full" in sys.argv:
        orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
        for name, mk in (("cg_gmg_jacobi_poisson@129^3", lambda: pr.Problem(dim=3, num_refs=7)),
                         ("cg_gmg_jacobi_poisson@257^3 (2x2x2 base cells)", lambda: pr.Problem(dim=3, num_refs=7, base=(2, 2, 2)))):
            p = mk()
            sens, h0, h1 = reduction_order_sensitivity(orc, p, gmg_desc(7), np.array(p.rhs()))
            full[name] = {"steps": len(h0) - 1, "steps_reordered": len(h1) - 1, "max_rel_move": float(sens.max()),
                          "per_step": [float(v) for v in sens], "backend": orc.kind}
    else:
        with open(GOLDEN) as f:
            full = {k: v for k, v in json.load(f).items() if "129^3" in k or "257^3" in k}
    res = _measure((3, 4, 5, 6))
    res.update(full)
    with open(GOLDEN, "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(f"{k:45s} steps {v['steps']:3d}  max rel move {v['max_rel_move']:.2e}")
