"""GPU: the grid of solver x preconditioner x cycle combinations against the oracle (history within 1e-10 relative per
step — or 10 x what the reference's own history moves under a reordered sum, whichever is larger: helpers.sens_tol —,
same step count, same convergence flag).  Gauss-Seidel-type preconditioners are left to their own tests: the
device sweeps in a multicolour ordering, so the oracle has to be run on the colour-permuted problem
(tests/test_gpu_solver.py, tests/test_ilu.py)."""
import os

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, oracle_levels, rel_hist_err, sens_tol
from ugcore_b200 import problems as pr

pytestmark = pytest.mark.gpu
CC = {"iterations": 60, "absolute": 1e-12, "reduction": 1e-8}


def _prob(name):
    return pr.Problem(dim=3, num_refs=3) if name == "poisson" else pr.Problem(dim=3, num_refs=3, problem=pr.CONVDIFF, eps=0.1)


def _orc():
    return oracle.Oracle("ref" if oracle.have_ref() else "port")


@pytest.mark.parametrize("problem", ["poisson", "convdiff"])
@pytest.mark.parametrize("solver", ["cg", "bicgstab", "linear", "gmres"])
def test_one_level_preconditioners(problem, solver):
    import ugcore_b200 as ug
    prob, orc = _prob(problem), _orc()
    for pc in (None, {"type": "jac", "damping": 0.7}, {"type": "ilu"}, {"type": "ilu", "beta": 0.5}):
        desc = {"type": solver, "restart": 8, "precond": pc, "convCheck": CC}
        x, ok, h = ug.Solver(desc, prob.matrix()).apply(prob.rhs())
        osol = oracle.OSolver(orc, desc, orc.matrix(prob.matrix()))
        xo, oko, ho = osol.apply(np.array(prob.rhs()))
        if not np.isfinite(ho).all():        # CG on the non-symmetric operator may break down: both must say so
            assert not ok and not oko, (solver, pc)
            continue
        assert ok == oko and abs(len(h) - len(ho)) <= 1, (solver, pc)
        if oko:
            tol = sens_tol(orc, osol, np.array(prob.rhs()))
            assert rel_hist_err(h, ho) < tol, (solver, pc, rel_hist_err(h, ho), tol)


@pytest.mark.parametrize("problem", ["poisson", "convdiff"])
@pytest.mark.parametrize("solver", ["cg", "bicgstab", "linear", "gmres"])
def test_gmg_cycles_and_smoothers(problem, solver):
    import ugcore_b200 as ug
    prob, orc = _prob(problem), _orc()
    lv = oracle_levels(orc, prob)
    for sm in ({"type": "jac", "damp": 0.66}, {"type": "ilu"}):
        for cycle in ("V", "W", "F"):
            desc = gmg_desc(3, solver=solver, smoother=sm, cycle=cycle, reduction=1e-8)
            desc["restart"] = 4
            x, ok, h = ug.Solver.from_problem(desc, prob).apply(prob.rhs())
            osol = oracle.OSolver(orc, desc, lv[3][0], lv)
            xo, oko, ho = osol.apply(np.array(prob.rhs()))
            # (CG is not a solver for the non-symmetric operator: it need not converge, but must fail the same way)
            assert ok == oko and abs(len(h) - len(ho)) <= 1, (solver, sm, cycle)
            assert oko or (solver == "cg" and problem == "convdiff"), (solver, sm, cycle)
            if oko:
                tol = sens_tol(orc, osol, np.array(prob.rhs()))
                assert rel_hist_err(h, ho) < tol, (solver, sm, cycle, rel_hist_err(h, ho), tol)
