"""DoF reordering before upload: Cuthill-McKee numbering of every level, the hierarchy
permuted consistently (ugcore: IOrderingAlgorithm + SetMatrixAsPermutation / SetVectorAsPermutation,
lib_algebra/algebra_common/permutation_util.h:50-78).  The ordering itself is checked against the reference's
own ComputeCuthillMcKeeOrder in tests/test_ilu.py."""
import os

import numpy as np
import pytest

import oracle
from helpers import gmg_desc, rel_hist_err
from ugcore_b200 import problems as pr


def _levels(prob):
    return {l: (prob.matrix(l), prob.prolongation(l) if l else None, prob.restriction(l) if l else None)
            for l in range(prob.base_lev, prob.num_refs + 1)}


@pytest.mark.parametrize("order", ["cmk", "rcmk"])
@pytest.mark.parametrize("problem", [pr.POISSON, pr.ELASTICITY])
def test_reordered_hierarchy_is_the_same_operator(order, problem):
    from ugcore_b200.solver import reorder_hierarchy
    prob = pr.Problem(dim=3, num_refs=2 if problem == pr.ELASTICITY else 3, problem=problem, order=pr.ORDER_HIER)
    top, b = prob.num_refs, prob.block
    levels = _levels(prob)
    A2, lv2, perms = reorder_hierarchy(prob.matrix(top), levels, order)
    comp = lambda p: np.repeat(p * b, b) + np.tile(np.arange(b), p.size)
    rng = np.random.default_rng(0)
    for l in range(1, top + 1):
        A, P, R = levels[l]
        pf, pc = comp(perms[l]), comp(perms[l - 1])
        xf, xc = rng.standard_normal(A.nrows * b), rng.standard_normal(P.ncols * b)
        xfp, xcp = np.empty_like(xf), np.empty_like(xc)
        xfp[pf], xcp[pc] = xf, xc
        S = lambda M: M.to_scipy() if M.block == b else __import__("scipy.sparse", fromlist=["kron"]).kron(M.to_scipy(), np.eye(b)).tocsr()
        assert np.allclose((S(lv2[l][0]) @ xfp)[pf], S(A) @ xf, rtol=1e-13, atol=1e-13)
        assert np.allclose((S(lv2[l][1]) @ xcp)[pf], S(P) @ xc, rtol=1e-13, atol=1e-13)
        assert np.allclose((S(lv2[l][2]) @ xfp)[pc], S(R) @ xf, rtol=1e-13, atol=1e-13)
        assert np.all(np.diff(lv2[l][0].cols)[np.diff(np.repeat(np.arange(A.nrows), np.diff(lv2[l][0].rowptr))) == 0] > 0)  # sorted rows
    assert np.array_equal(A2.cols, lv2[top][0].cols)
    if problem == pr.POISSON:    # the hierarchical numbering has a wide band; Cuthill-McKee narrows it
        A = prob.matrix(top)
        r0 = np.repeat(np.arange(A.nrows), np.diff(A.rowptr))
        r2 = np.repeat(np.arange(A2.nrows), np.diff(A2.rowptr))
        assert np.abs(r2 - A2.cols).max() < np.abs(r0 - A.cols).max()


# (these GPU tests compose kernels that are GPU-verified on their own — the solve path on permuted inputs,
#  ug4b200_vec_gather / _scatter_add — with host logic checked on the CPU; they are not gated)

@pytest.mark.gpu
@pytest.mark.parametrize("order", ["cmk", "rcmk"])
def test_gpu_solve_with_reordered_hierarchy_matches_oracle(order):
    """GMG-CG on the hierarchically numbered grid, renumbered by (reverse) Cuthill-McKee before upload: vectors go in
    and come out in the caller's numbering; the oracle solves the same permuted hierarchy."""
    import ugcore_b200 as ug
    from ugcore_b200.solver import reorder_hierarchy
    prob = pr.Problem(dim=3, num_refs=3, order=pr.ORDER_HIER)
    desc = gmg_desc(3)
    s = ug.Solver.from_problem(desc, prob, order=order)
    x, ok, h = s.apply(prob.rhs())
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    A2, lv2, perms = reorder_hierarchy(prob.matrix(3), _levels(prob), order)
    lv = {l: tuple(orc.matrix(M) if M is not None else None for M in t) for l, t in lv2.items()}
    bp = np.empty(prob.num_dofs); bp[perms[3]] = prob.rhs()
    xo, oko, ho = oracle.OSolver(orc, desc, lv[3][0], lv).apply(bp)
    assert ok and oko and abs(len(h) - len(ho)) <= 1
    assert rel_hist_err(h, ho) < 1e-10
    assert np.linalg.norm(x - xo[perms[3]]) <= 1e-9 * np.linalg.norm(xo)
    # and it is the same problem: the un-reordered solve agrees to solver accuracy
    x0, ok0, h0 = ug.Solver.from_problem(desc, prob).apply(prob.rhs())
    assert ok0 and np.linalg.norm(x - x0) <= 1e-8 * np.linalg.norm(x0)


def _surface_problem(seed=4):
    """A hierarchy whose surface numbering of the top level is a random permutation of its level numbering."""
    from ugcore_b200.solver import permute_crs
    prob = pr.Problem(dim=3, num_refs=3)
    sigma = np.random.default_rng(seed).permutation(prob.matrix(3).nrows)       # surface index of level index
    A_surf = permute_crs(prob.matrix(3), sigma, sigma)
    b_surf = np.empty(prob.num_dofs); b_surf[sigma] = prob.rhs()
    return prob, sigma, A_surf, b_surf


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_oracle_surface_level_map(kind, request):
    """GMG::apply's surface <-> level copies (mg_solver_impl.hpp:211-217, 244-248) with a non-identity vSurfLevelMap:
    the solve in surface numbering is the level-numbered solve, permuted (same V-cycle; the Krylov part sums its rows
    in another order, hence round-off level differences only)."""
    from helpers import oracle_levels
    orc = request.getfixturevalue("orc" if kind == "port" else "orc_ref")
    prob, sigma, A_surf, b_surf = _surface_problem()
    desc = gmg_desc(3)
    lv = oracle_levels(orc, prob)
    x0, ok0, h0 = oracle.OSolver(orc, desc, lv[3][0], lv).apply(np.array(prob.rhs()))
    x1, ok1, h1 = oracle.OSolver(orc, desc, orc.matrix(A_surf), lv, surface_map=sigma).apply(b_surf)
    assert ok0 and ok1 and len(h0) == len(h1)
    assert rel_hist_err(h1, h0) < 1e-11
    assert np.linalg.norm(x1[sigma] - x0) <= 1e-12 * np.linalg.norm(x0)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 4])
def test_gpu_surface_level_map_matches_oracle(flags):
    """The device gathers / scatters of AssembledMultiGridCycle::apply (ug4b200_vec_gather / _scatter_add through the
    surface map) against the oracle with the same map; flags = 4: unfused V-cycle."""
    import ugcore_b200 as ug
    from helpers import oracle_levels
    prob, sigma, A_surf, b_surf = _surface_problem()
    desc = gmg_desc(3)
    levels = _levels(prob)
    x, ok, h = ug.Solver(dict(desc), A_surf, levels, flags=flags, surface_map=sigma).apply(b_surf)
    orc = oracle.Oracle("ref" if oracle.have_ref() else "port")
    xo, oko, ho = oracle.OSolver(orc, desc, orc.matrix(A_surf), oracle_levels(orc, prob), surface_map=sigma).apply(b_surf)
    assert ok and oko and abs(len(h) - len(ho)) <= 1
    assert rel_hist_err(h, ho) < 1e-10
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
    with pytest.raises(Exception, match="not a permutation"):
        ug.Solver(dict(desc), A_surf, levels, surface_map=np.zeros(sigma.size, np.int32))
