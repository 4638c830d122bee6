/*
 * oracle/solvers.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Restated control flow; see solvers.h for the reference line map.
 */
#include "solvers.h"
#include <cmath>
#include <stdexcept>

namespace oracle {

// ---- reductions ------------------------------------------------------------------------
// Mode 0 (default, THE reference): the backend's dot / norm = ugcore's strictly sequential left-to-right sums
// (vector_impl.h:72-79, 323-329).  Mode 1: the same products summed by a pairwise tree.  Mode 1 is not ugcore's
// arithmetic; it exists only so that a test can MEASURE how far the reference's own residual history moves when
// nothing but the summation order of its reductions changes (tests/test_reduction_order.py) — the yardstick for
// the tolerance a GPU reduction tree can be held to.
static int g_reduction_mode = 0;
void set_reduction_mode(int mode) { g_reduction_mode = mode; }
int reduction_mode() { return g_reduction_mode; }
namespace {
double pairwise(const double* a, const double* b, int64_t n)
{
	if (n <= 8) { double s = 0.0; for (int64_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
	const int64_t h = n / 2;
	return pairwise(a, b, h) + pairwise(a + h, b + h, n - h);
}
double rdot(Backend& bk, const Vec& a, const Vec& b) { return g_reduction_mode == 0 ? bk.dot(a, b) : pairwise(a.data(), b.data(), a.len()); }
double rnorm(Backend& bk, const Vec& a) { return g_reduction_mode == 0 ? bk.norm(a) : std::sqrt(pairwise(a.data(), a.data(), a.len())); }
} // namespace

// ---- StdConvCheck (convergence_check_impl.h:85-169, 246-257) --------------------

void StdConvCheck::start_defect(double d)
{
	history.clear();
	initialDefect = d; currentDefect = d; currentStep = 0; ratesProduct = 1;
	history.push_back(d);
}
void StdConvCheck::update_defect(double d)
{
	lastDefect = currentDefect; currentDefect = d; currentStep++;
	ratesProduct *= d / lastDefect;
	history.push_back(d);
}
bool StdConvCheck::is_valid_number(double v)
{
	if (v == 0.0) return true;
	return v >= std::numeric_limits<double>::min() && v <= std::numeric_limits<double>::max() && v == v && v >= 0.0;
}
bool StdConvCheck::iteration_ended() const
{
	if (!is_valid_number(currentDefect)) return true;
	if (step() >= maxSteps) return true;
	if (defect() < minDefect) return true;
	if (reduction() < relReduction) return true;
	return false;
}
bool StdConvCheck::post() const
{
	bool success = false;
	if (defect() < minDefect) success = true;
	if (reduction() < relReduction) success = true;
	return success;
}

// ---- IPreconditioner (preconditioner.h:239-348) ---------------------------------

bool Preconditioner::init(const Mat& A_)
{
	A = &A_;
	if (!preprocess()) return false;
	inited = true;
	return true;
}
bool Preconditioner::apply(Vec& c, const Vec& d)
{
	if (!inited) return false;
	if (!step(c, d)) return false;
	const double kappa = damping;
	if (kappa != 1.0) bk.scale(c, kappa);
	return true;
}
bool Preconditioner::apply_update_defect(Vec& c, Vec& d)
{
	if (!apply(c, d)) return false;
	bk.matmul_minus(*A, d, c);
	return true;
}

// ---- Jacobi (jacobi.h:155-300) --------------------------------------------------

bool Jacobi::preprocess()
{
	if (A->nrows != A->ncols) return false;
	diagInv.reset(bk.jacobi_prepare(*A, damping, block));
	return true;
}
bool Jacobi::step(Vec& c, const Vec& d) { bk.jacobi_step(*diagInv, c, d); return true; }
bool Jacobi::apply(Vec& c, const Vec& d)
{
	if (!inited) return false;
	// constant damping is already contained in diagInv (jacobi.h:283-290)
	return step(c, d);
}

// ---- Gauss-Seidel (gauss_seidel.h:221-230, 270-380) ------------------------------

bool GaussSeidel::step(Vec& c, const Vec& d)
{
	switch (kind) {
		case FORWARD: bk.gs_step_LL(*A, c, d, relax); break;
		case BACKWARD: bk.gs_step_UR(*A, c, d, relax); break;
		case SYMMETRIC: bk.sgs_step(*A, c, d, relax); break;
	}
	return true;
}

// ---- ILU (ilu.h) ---------------------------------------------------------------------------

bool ILU::preprocess()
{
	factors.reset(bk.ilu_factorize(*A, beta, sortEps));   // m_ILU = mat; Factorize...(m_ILU) :563-576
	h.reset(bk.vector(A->ncols, A->block));               // m_h.resize(m_ILU.num_cols()) :556
	return true;
}
bool ILU::step(Vec& c, const Vec& d)
{
	bk.ilu_invert_L(*factors, *h, d);                     // h := L^-1 d      :596
	bk.ilu_invert_U(*factors, c, *h, invEps);             // c := U^-1 h      :598
	return true;
}

// ---- ILinearOperatorInverse::apply (preconditioned_linear_operator_inverse.h:152-160)

bool InverseOperator::apply(Vec& x, const Vec& b)
{
	VecP bTmp(bk.vector(b.n, b.block));
	bk.assign(*bTmp, b);
	return apply_return_defect(x, *bTmp);
}

// ---- LU (lu.h:236-263, 322-380) ---------------------------------------------------

bool LU::init(const Mat& A_)
{
	A = &A_;
	if (A_.nrows == 0) return true;
	lu.reset(bk.lu_init(A_));
	if (!lu) throw std::runtime_error("LU: matrix is singular");
	return true;
}
bool LU::apply(Vec& x, const Vec& b)
{
	if (A->nrows == 0) return true;
	bk.lu_apply(*lu, x, b);
	return true;
}
bool LU::apply_return_defect(Vec& x, Vec& b)
{
	if (!apply(x, b)) return false;
	bk.matmul_minus(*A, b, x);
	return true;
}

// ---- CG (cg.h:103-242) -------------------------------------------------------------

bool CG::apply_return_defect(Vec& x, Vec& b)
{
	Vec& r = b;
	bk.matmul_minus(*A, r, x);                 // r := b - A x
	VecP q(bk.vector(r.n, r.block)), z(bk.vector(x.n, x.block)), p(bk.vector(x.n, x.block));
	if (precond) { if (!precond->apply(*z, r)) return false; }
	else bk.assign(*z, r);
	conv.start_defect(rnorm(bk, r));
	bk.assign(*p, *z);
	double rhoOld = rdot(bk, *z, r), rho;
	while (!conv.iteration_ended()) {
		bk.apply(*A, *q, *p);                  // q = A p
		double lambda = rdot(bk, *q, *p);
		if (lambda == 0.0) {
			if (p->len()) return false;
			lambda = 1.0;
		}
		const double alpha = rhoOld / lambda;
		bk.scale_add2(x, 1.0, x, alpha, *p);
		bk.scale_add2(r, 1.0, r, -alpha, *q);
		conv.update_defect(rnorm(bk, r));
		if (conv.iteration_ended()) break;
		if (precond) { if (!precond->apply(*z, r)) return false; }
		else bk.assign(*z, r);
		rho = rdot(bk, *z, r);
		const double beta = rho / rhoOld;
		bk.scale_add2(*p, beta, *p, 1.0, *z);
		rhoOld = rho;
	}
	return conv.post();
}

// ---- BiCGStab (bicgstab.h:112-383) ---------------------------------------------------

bool BiCGStab::apply_return_defect(Vec& x, Vec& b)
{
	bk.matmul_minus(*A, b, x);
	Vec& r = b;
	VecP r0(bk.vector(r.n, r.block)), p(bk.vector(r.n, r.block)), v(bk.vector(r.n, r.block)),
	     t(bk.vector(r.n, r.block)), s(bk.vector(r.n, r.block)), q(bk.vector(x.n, x.block));
	conv.start_defect(rnorm(bk, r));
	double rho = 1, alpha = 1, omega = 1, norm_r0 = 0.0;
	bool bRestart = true;
	while (!conv.iteration_ended()) {
		if (numRestarts > 0 && (conv.step() % numRestarts == 0)) bRestart = true;
		if (bRestart) {
			bk.assign(*r0, r);
			bk.set(*p, 0.0); alpha = 0.0;
			bk.set(*v, 0.0); omega = 1.0;
			rho = 1.0;
			norm_r0 = conv.defect();
			bRestart = false;
		}
		const double rhoOld = rho;
		if (!r.len()) rho = 1.0; else rho = rdot(bk, *r0, r);
		const double norm_r = conv.defect();
		if (std::fabs(rho) / (norm_r * norm_r0) <= minOrtho) bRestart = true;
		if (rhoOld == 0.0) return false;
		const double beta = (rho / rhoOld) * (alpha / omega);
		bk.scale_add3(*p, 1.0, r, beta, *p, -beta * omega, *v);
		if (precond) { if (!precond->apply(*q, *p)) return false; }
		else bk.assign(*q, *p);
		bk.apply(*A, *v, *q);
		if (!v->len()) alpha = 1.0; else alpha = rdot(bk, *v, *r0);
		if (alpha == 0.0) return false;
		alpha = rho / alpha;
		bk.scale_add2(x, 1.0, x, alpha, *q);
		bk.scale_add2(*s, 1.0, r, -alpha, *v);
		conv.update_defect(rnorm(bk, *s));
		if (conv.iteration_ended()) { bk.assign(r, *s); break; }
		if (precond) { if (!precond->apply(*q, *s)) return false; }
		else bk.assign(*q, *s);
		bk.apply(*A, *t, *q);
		double tt;
		if (!t->len()) tt = 1.0; else tt = rdot(bk, *t, *t);
		if (!s->len()) omega = 1.0; else omega = rdot(bk, *s, *t);
		if (tt == 0.0) return false;
		omega = omega / tt;
		bk.scale_add2(x, 1.0, x, omega, *q);
		bk.scale_add2(r, 1.0, *s, -omega, *t);
		conv.update_defect(rnorm(bk, r));
		if (omega == 0.0) return false;
	}
	return conv.post();
}

// ---- LinearSolver (linear_solver.h:114-196) --------------------------------------------

bool LinearSolver::apply_return_defect(Vec& x, Vec& b)
{
	Vec& d = b;
	bk.matmul_minus(*A, d, x);
	VecP c(bk.vector(x.n, x.block));
	conv.start_defect(rnorm(bk, d));
	while (!conv.iteration_ended()) {
		if (precond) { if (!precond->apply_update_defect(*c, d)) return false; }
		else { bk.assign(*c, d); bk.matmul_minus(*A, d, *c); }
		bk.add(x, *c);
		conv.update_defect(rnorm(bk, d));
	}
	return conv.post();
}

// ---- GMRES (gmres.h:104-278) -----------------------------------------------------------------

bool GMRES::apply_return_defect(Vec& x, Vec& b)
{
	VecP spR(bk.vector(b.n, b.block));
	bk.assign(*spR, b);                                   // spR = b.clone()            :113
	bk.matmul_minus(*A, *spR, x);                         // b - A x                    :116
	conv.start_defect(rnorm(bk, *spR));                     //                            :122
	const size_t m = restart;
	std::vector<VecP> v(m + 1);
	std::vector<std::vector<double> > h(m + 1, std::vector<double>(m + 1, 0.0));
	std::vector<double> gamma(m + 1), c(m + 1), s(m + 1);
	while (!conv.iteration_ended()) {
		if (!v[0]) v[0].reset(bk.vector(x.n, x.block));
		if (precond) { if (!precond->apply(*v[0], *spR)) return false; }        // :141-146
		else std::swap(v[0], spR);                                                // :148-150
		gamma[0] = rnorm(bk, *v[0]);                                                // :162
		bk.scale(*v[0], 1. / gamma[0]);                                           // :165
		size_t numIter = 0;
		for (size_t j = 0; j < m; ++j) {
			numIter = j;
			if (!v[j + 1]) v[j + 1].reset(bk.vector(x.n, x.block));
			bk.apply(*A, *spR, *v[j]);                                            // :182
			if (precond) { if (!precond->apply(*v[j + 1], *spR)) return false; }  // :185-190
			else std::swap(v[j + 1], spR);                                        // :192-194
			for (size_t i = 0; i <= j; ++i) {
				h[i][j] = rdot(bk, *v[j + 1], *v[i]);                               // :211
				bk.scale_add2(*v[j + 1], 1.0, *v[j + 1], (-1) * h[i][j], *v[i]);  // VecScaleAppend :214, :341-345
			}
			h[j + 1][j] = rnorm(bk, *v[j + 1]);                                     // :218
			for (size_t i = 0; i < j; ++i) {                                      // :221-228
				const double hij = h[i][j], hi1j = h[i + 1][j];
				h[i][j] = c[i + 1] * hij + s[i + 1] * hi1j;
				h[i + 1][j] = s[i + 1] * hij - c[i + 1] * hi1j;
			}
			const double alpha = std::sqrt(h[j][j] * h[j][j] + h[j + 1][j] * h[j + 1][j]);   // :231
			s[j + 1] = h[j + 1][j] / alpha;
			c[j + 1] = h[j][j] / alpha;
			h[j][j] = alpha;
			gamma[j + 1] = s[j + 1] * gamma[j];                                   // :239-240
			gamma[j] = c[j + 1] * gamma[j];
			if (!precond) conv.update_defect(gamma[j + 1]);                       // :242-252
			bk.scale(*v[j + 1], 1. / (h[j + 1][j]));                              // :255
		}
		for (size_t i = numIter;; --i) {                                          // :259-269
			for (size_t j = i + 1; j <= numIter; ++j) gamma[i] -= h[i][j] * gamma[j];
			gamma[i] /= h[i][i];
			bk.scale_add2(x, 1.0, x, gamma[i], *v[i]);
			if (i == 0) break;
		}
		bk.assign(*spR, b);                                                       // :272-273
		bk.matmul_minus(*A, *spR, x);
		if (precond) conv.update_defect(rnorm(bk, *spR));                           // :275-276
	}
	return conv.post();
}

// ---- GMG (mg_solver_impl.hpp) -----------------------------------------------------------

void GMG::set_level(int level, const Mat* A, const Mat* P, const Mat* R)
{
	if (level < baseLev || level > topLev) throw std::runtime_error("GMG::set_level: level out of range");
	if ((int)lev.size() != topLev - baseLev + 1) lev.resize(topLev - baseLev + 1);
	LevData& ld = L(level);
	ld.A = A; ld.P = P; ld.R = R;
}

void GMG::set_level_smoother_matrix(int level, const Mat* S)
{
	if (level < baseLev || level > topLev) throw std::runtime_error("GMG::set_level_smoother_matrix: level out of range");
	if ((int)lev.size() != topLev - baseLev + 1) lev.resize(topLev - baseLev + 1);
	L(level).S = S;
}

// init(): mg_solver_impl.hpp:378-496 — level memory, smoother clones + init
// (:1135-1169), base solver init (:1171-1229); level operators are handed in
// re-discretised (assemble_level_operator :526-752, rap = false).
bool GMG::init(const Mat& A_)
{
	surfaceMat = &A_;
	if (baseLev > topLev) throw std::runtime_error("GMG::init: Base Level greater than Surface level.");
	if (!smootherProto) throw std::runtime_error("GMG::init: PreSmoother not set.");
	if (!baseSolver) throw std::runtime_error("GMG::init: Base Solver not set.");
	for (int l = baseLev; l <= topLev; ++l) {
		LevData& ld = L(l);
		if (!ld.A) throw std::runtime_error("GMG::init: level operator missing");
		ld.sc.reset(bk.vector(ld.A->nrows, ld.A->block));
		ld.sd.reset(bk.vector(ld.A->nrows, ld.A->block));
		ld.st.reset(bk.vector(ld.A->nrows, ld.A->block));
		if (l > baseLev) {
			if (!ld.P || !ld.R) throw std::runtime_error("GMG::init: transfer missing");
			ld.pre.reset(smootherProto->clone());
			ld.post.reset(smootherProto->clone());
			const Mat& S = ld.S ? *ld.S : *ld.A;
			if (!ld.pre->init(S) || !ld.post->init(S)) return false;
		}
	}
	return baseSolver->init(*L(baseLev).A);
}

// apply(): mg_solver_impl.hpp:174-275
bool GMG::apply(Vec& c, const Vec& d)
{
	LevData& top = L(topLev);
	const int B = d.block;
	if (surfMap.empty()) bk.assign(*top.sd, d);   // surface -> level copy (:211-217), identity map
	else {
		if ((int64_t)surfMap.size() != top.sd->n) throw std::runtime_error("GMG: surface map has the wrong size");
		for (int64_t i = 0; i < top.sd->n; ++i) for (int t = 0; t < B; ++t) top.sd->data()[i * B + t] = d.data()[(int64_t)surfMap[i] * B + t];
	}
	bk.set(c, 0.0);                    // :231
	bk.set(*top.sc, 0.0);              // :234
	lmgc(topLev, cycleType);           // :238
	if (surfMap.empty()) bk.add(c, *top.sc);      // :244-248
	else for (int64_t i = 0; i < top.sc->n; ++i) for (int t = 0; t < B; ++t) c.data()[(int64_t)surfMap[i] * B + t] += top.sc->data()[i * B + t];
	if (damping != 1.0) bk.scale(c, damping); // :259-260
	return true;
}

// apply_update_defect(): mg_solver_impl.hpp:277-320
bool GMG::apply_update_defect(Vec& c, Vec& d)
{
	if (!apply(c, d)) return false;
	bk.matmul_minus(*surfaceMat, d, c);
	return true;
}

// lmgc(): mg_solver_impl.hpp:2089-2136
void GMG::lmgc(int l, int cycle)
{
	if (l == baseLev) { base_solve(topLev); return; }
	presmooth_and_restriction(l);
	if (l - 1 == baseLev) base_solve(l - 1);
	else if (cycle == F_CYCLE) { lmgc(l - 1, F_CYCLE); lmgc(l - 1, V_CYCLE); }
	else for (int i = 0; i < cycle; ++i) lmgc(l - 1, cycle);
	prolongation_and_postsmooth(l);
}

// presmooth_and_restriction(): mg_solver_impl.hpp:1685-1816
void GMG::presmooth_and_restriction(int l)
{
	LevData& lf = L(l); LevData& lc = L(l - 1);
	for (int nu = 0; nu < numPreSmooth; ++nu) {
		if (!lf.pre->apply(*lf.st, *lf.sd)) throw std::runtime_error("GMG: Smoothing step failed.");
		bk.matmul_minus(*lf.A, *lf.sd, *lf.st);            // :1726
		if (nu < numPreSmooth - 1) bk.add(*lf.sc, *lf.st); // :1729-1730
	}
	bk.set(*lc.sc, 0.0);                                   // :1780
	if (numPreSmooth > 0) bk.add(*lf.sc, *lf.st);          // :1783-1784
	// do_restrict: std_transfer_impl.h:791-792
	bk.apply_ignore_zero_rows(*lf.R, *lc.sd, dampRes, *lf.sd);
}

// prolongation_and_postsmooth(): mg_solver_impl.hpp:1818-1964
void GMG::prolongation_and_postsmooth(int l)
{
	LevData& lf = L(l); LevData& lc = L(l - 1);
	// prolongate: std_transfer_impl.h:738-740 — axpy(uFine, 0.0, uFine, dampProl, uCoarse)
	bk.axpy(*lf.P, *lf.st, 0.0, *lf.st, dampProl, *lc.sc);
	bk.add(*lf.sc, *lf.st);                                // :1905
	for (int nu = 0; nu < numPostSmooth; ++nu) {
		bk.matmul_minus(*lf.A, *lf.sd, *lf.st);            // :1919
		if (!lf.post->apply(*lf.st, *lf.sd)) throw std::runtime_error("GMG: Smoothing step failed.");
		bk.add(*lf.sc, *lf.st);                            // :1943
	}
	// m_LocalFullRefLevel == topLev on a fully refined grid (:1309, :1359)
	if (l >= topLev) bk.matmul_minus(*lf.A, *lf.sd, *lf.st); // :1954-1958
}

// base_solve(): mg_solver_impl.hpp:1967-2086 (non-gathered branch)
void GMG::base_solve(int l)
{
	LevData& ld = L(l);
	if (!baseSolver->apply(*ld.sc, *ld.sd)) throw std::runtime_error("GMG::lmgc: Base solver failed.");
	if (l >= topLev) bk.matmul_minus(*ld.A, *ld.sd, *ld.sc); // :2075-2078
}

} // namespace oracle
