"""The oracle's port against the compiled reference kernels on WHOLE SOLVES, over the grid of solver x preconditioner
combinations the descriptor layer offers (scalar and 3x3-block problems): identical convergence flags, histories and
iterates, bit for bit.  (Kernel by kernel the same is shown in tests/test_oracle.py; the solver control flow is one
piece of code for both backends, so what this pins is the port's arithmetic under every code path the solvers take.)"""
import numpy as np
import pytest

import oracle
from helpers import gmg_desc, oracle_levels
from ugcore_b200 import problems as pr

CC = {"iterations": 60, "absolute": 1e-12, "reduction": 1e-8}
ONE_LEVEL = [None, {"type": "jac", "damping": 0.7}, {"type": "gs"}, {"type": "bgs"}, {"type": "sgs"}, {"type": "ilu"},
             {"type": "ilu", "beta": 0.5}]
SMOOTHERS = [{"type": "jac", "damp": 0.66}, {"type": "gs", "relax": 1.0}, {"type": "sgs", "relax": 0.9}, {"type": "ilu"}]


def _problem(name):
    return {"poisson": lambda: pr.Problem(dim=3, num_refs=3), "convdiff": lambda: pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=0.1),
            "elasticity": lambda: pr.Problem(dim=3, num_refs=2, problem=pr.ELASTICITY)}[name]()


def _both(desc, prob, orc, orc_ref):
    out = []
    for o in (orc, orc_ref):
        pc = desc.get("precond")
        if isinstance(pc, dict) and pc.get("type") == "gmg":
            lv = oracle_levels(o, prob, pc["baseLevel"], pc["topLevel"])
            s = oracle.OSolver(o, desc, lv[pc["topLevel"]][0], lv)
        else:
            s = oracle.OSolver(o, desc, o.matrix(prob.matrix()))
        out.append(s.apply(np.array(prob.rhs())))
    (x0, ok0, h0), (x1, ok1, h1) = out
    assert ok0 == ok1 and np.array_equal(h0, h1) and np.array_equal(x0, x1)
    return ok1, h1


@pytest.mark.parametrize("solver", ["cg", "bicgstab", "linear", "gmres"])
@pytest.mark.parametrize("pc", range(len(ONE_LEVEL)))
@pytest.mark.parametrize("problem", ["poisson", "convdiff"])
def test_port_equals_reference_one_level(solver, pc, problem, orc, orc_ref):
    desc = {"type": solver, "restart": 8, "precond": ONE_LEVEL[pc], "convCheck": CC}
    ok, h = _both(desc, _problem(problem), orc, orc_ref)
    assert np.isfinite(h).all()
    if solver in ("bicgstab", "gmres") and ONE_LEVEL[pc] and ONE_LEVEL[pc]["type"] == "ilu":
        assert ok and h[-1] < 1e-8 * h[0]


@pytest.mark.parametrize("solver", ["cg", "bicgstab", "linear", "gmres"])
@pytest.mark.parametrize("sm", range(len(SMOOTHERS)))
@pytest.mark.parametrize("cycle", ["V", "W"])
def test_port_equals_reference_gmg(solver, sm, cycle, orc, orc_ref):
    prob = _problem("convdiff" if solver in ("bicgstab", "gmres") else "poisson")
    desc = gmg_desc(3, solver=solver, smoother=SMOOTHERS[sm], cycle=cycle, reduction=1e-8)
    desc["restart"] = 4
    ok, h = _both(desc, prob, orc, orc_ref)
    assert ok and h[-1] < 1e-8 * h[0]


@pytest.mark.parametrize("desc", [gmg_desc(2, reduction=1e-8, its=200),
                                  gmg_desc(2, smoother={"type": "sgs", "relax": 1.0}, reduction=1e-8, its=200),
                                  {"type": "cg", "precond": {"type": "jac", "damping": 0.8}, "convCheck": dict(CC, iterations=300)},
                                  {"type": "bicgstab", "precond": {"type": "sgs"}, "convCheck": dict(CC, iterations=300)}])
def test_port_equals_reference_blocks(desc, orc, orc_ref):
    """3x3 blocks (the port has no block ILU: covered by the ref backend alone in tests/test_ilu.py)."""
    ok, h = _both(desc, _problem("elasticity"), orc, orc_ref)
    assert ok
