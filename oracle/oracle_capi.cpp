/*
 * oracle/oracle_capi.cpp — C ABI glue of the CPU oracle.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 */
#include "oracle.h"
#include "backend.h"
#include "solvers.h"
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

using namespace oracle;

struct oracle_mat { std::unique_ptr<Mat> m; };

namespace {
std::string g_err;

Backend& BK()
{
#ifdef ORACLE_WITH_UGREF
	static std::unique_ptr<Backend> b(make_ref_backend());
#else
	static std::unique_ptr<Backend> b(make_port_backend());
#endif
	return *b;
}

// wrap a raw array in a backend vector (copy in), copy back on demand
struct VIO {
	VecP v; double* host;
	VIO(int64_t len, int block, const double* src, double* dst = nullptr) : host(dst)
	{
		if (block < 1 || len % block) throw std::runtime_error("vector length not a multiple of the block size");
		v.reset(BK().vector(len / block, block));
		if (src) std::memcpy(v->data(), src, sizeof(double) * len);
	}
	void back() { if (host) std::memcpy(host, v->data(), sizeof(double) * v->len()); }
};

template <class F> int guard(F f)
{
	try { return f(); }
	catch (const std::exception& e) { g_err = e.what(); return -1; }
	catch (...) { g_err = "unknown exception"; return -1; }
}

LinearIterator* make_smoother(int kind, double damp, double ilu_beta = 0.0)
{
	switch (kind) {
		case ORACLE_PRECOND_ILU: return new ILU(BK(), ilu_beta);
		case ORACLE_PRECOND_JACOBI: return new Jacobi(BK(), damp);
		case ORACLE_PRECOND_GS: { GaussSeidel* g = new GaussSeidel(BK(), GaussSeidel::FORWARD); g->relax = damp; return g; }
		case ORACLE_PRECOND_BGS: { GaussSeidel* g = new GaussSeidel(BK(), GaussSeidel::BACKWARD); g->relax = damp; return g; }
		case ORACLE_PRECOND_SGS: { GaussSeidel* g = new GaussSeidel(BK(), GaussSeidel::SYMMETRIC); g->relax = damp; return g; }
	}
	throw std::runtime_error("unknown smoother / preconditioner kind");
}
} // namespace

struct oracle_solver {
	oracle_solver_desc d;
	std::unique_ptr<InverseOperator> inv;
	LinearIterator* precond = nullptr; // owned by inv (or by lone below)
	std::unique_ptr<LinearIterator> lone;
	const oracle_mat* precond_matrix = nullptr; // matrix the (one-level) preconditioner is initialised with instead of A
	GMG* gmg = nullptr;
};

extern "C" {

const char* oracle_backend_name(void) { return BK().name(); }
const char* oracle_last_error(void) { return g_err.c_str(); }

oracle_mat* oracle_mat_create(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals)
{
	try { oracle_mat* A = new oracle_mat; A->m.reset(BK().matrix(block, nrows, ncols, rowptr, cols, vals)); return A; }
	catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void oracle_mat_destroy(oracle_mat* A) { delete A; }
int64_t oracle_mat_nnz(const oracle_mat* A) { return BK().nnz(*A->m); }
int64_t oracle_mat_rows(const oracle_mat* A) { return A->m->nrows; }
int64_t oracle_mat_cols(const oracle_mat* A) { return A->m->ncols; }
int oracle_mat_export(const oracle_mat* A, int64_t* rowptr, int* cols, double* vals)
{ return guard([&] { BK().export_crs(*A->m, rowptr, cols, vals); return 0; }); }
oracle_mat* oracle_mat_transpose(const oracle_mat* A, int keep_zeros)
{
	try { oracle_mat* T = new oracle_mat; T->m.reset(BK().transpose(*A->m, keep_zeros != 0)); return T; }
	catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

oracle_mat* oracle_mat_rap(const oracle_mat* R, const oracle_mat* A, const oracle_mat* P)
{
	try { oracle_mat* T = new oracle_mat; T->m.reset(BK().rap(*R->m, *A->m, *P->m)); return T; }
	catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

int oracle_axpy(const oracle_mat* A, double* dest, double alpha, const double* v, double beta, const double* w, int vb)
{
	return guard([&] {
		VIO D(A->m->nrows * vb, vb, dest, dest), W(A->m->ncols * vb, vb, w);
		if (v == dest || alpha == 0.0) BK().axpy(*A->m, *D.v, alpha, *D.v, beta, *W.v);
		else { VIO V(A->m->nrows * vb, vb, v); BK().axpy(*A->m, *D.v, alpha, *V.v, beta, *W.v); }
		D.back(); return 0;
	});
}
int oracle_vec_set_random(int64_t nblocks, int block, unsigned seed, double from, double to, double* out, double* maxnorm)
{
	return guard([&] {
		VIO A(nblocks * block, block, nullptr, out);
		std::srand(seed);
		BK().set_random(*A.v, from, to);
		*maxnorm = BK().maxnorm(*A.v);
		A.back(); return 0;
	});
}
int oracle_apply_transposed(const oracle_mat* A, double* y, const double* x)
{
	return guard([&] {
		const int vb = A->m->block;
		VIO Y(A->m->ncols * vb, vb, y, y), X(A->m->nrows * vb, vb, x);
		BK().apply_transposed(*A->m, *Y.v, *X.v);
		Y.back(); return 0;
	});
}
int oracle_apply(const oracle_mat* A, double* y, const double* x, int vb)
{
	return guard([&] {
		VIO Y(A->m->nrows * vb, vb, nullptr, y), X(A->m->ncols * vb, vb, x);
		BK().apply(*A->m, *Y.v, *X.v); Y.back(); return 0;
	});
}
int oracle_matmul_minus(const oracle_mat* A, double* y, const double* x, int vb)
{
	return guard([&] {
		VIO Y(A->m->nrows * vb, vb, y, y), X(A->m->ncols * vb, vb, x);
		BK().matmul_minus(*A->m, *Y.v, *X.v); Y.back(); return 0;
	});
}
int oracle_apply_ignore_zero_rows(const oracle_mat* A, double* dest, double beta, const double* w, int vb)
{
	return guard([&] {
		VIO D(A->m->nrows * vb, vb, dest, dest), W(A->m->ncols * vb, vb, w);
		BK().apply_ignore_zero_rows(*A->m, *D.v, beta, *W.v); D.back(); return 0;
	});
}

double oracle_dot(int64_t n, int block, const double* a, const double* b)
{ VIO A(n * block, block, a), B(n * block, block, b); return BK().dot(*A.v, *B.v); }
double oracle_norm(int64_t n, int block, const double* a)
{ VIO A(n * block, block, a); return BK().norm(*A.v); }
int oracle_scale_add2(int64_t len, double* d, double a1, const double* v1, double a2, const double* v2)
{
	return guard([&] {
		VIO D(len, 1, d, d), V1(len, 1, v1), V2(len, 1, v2);
		BK().scale_add2(*D.v, a1, v1 == d ? *D.v : *V1.v, a2, v2 == d ? *D.v : *V2.v); D.back(); return 0;
	});
}
int oracle_scale_add3(int64_t len, double* d, double a1, const double* v1, double a2, const double* v2, double a3, const double* v3)
{
	return guard([&] {
		VIO D(len, 1, d, d), V1(len, 1, v1), V2(len, 1, v2), V3(len, 1, v3);
		BK().scale_add3(*D.v, a1, v1 == d ? *D.v : *V1.v, a2, v2 == d ? *D.v : *V2.v, a3, v3 == d ? *D.v : *V3.v);
		D.back(); return 0;
	});
}

int oracle_jacobi(const oracle_mat* A, double damp, int block_inverse, double* c, const double* d)
{
	return guard([&] {
		const int vb = A->m->block;
		std::unique_ptr<DiagInv> D(BK().jacobi_prepare(*A->m, damp, block_inverse != 0));
		VIO Cv(A->m->nrows * vb, vb, nullptr, c), Dv(A->m->nrows * vb, vb, d);
		BK().jacobi_step(*D, *Cv.v, *Dv.v); Cv.back(); return 0;
	});
}
int oracle_gs(const oracle_mat* A, int kind, double relax, double* c, const double* d)
{
	return guard([&] {
		const int vb = A->m->block;
		VIO Cv(A->m->nrows * vb, vb, c, c), Dv(A->m->nrows * vb, vb, d);
		if (kind == 0) BK().gs_step_LL(*A->m, *Cv.v, *Dv.v, relax);
		else if (kind == 1) BK().gs_step_UR(*A->m, *Cv.v, *Dv.v, relax);
		else BK().sgs_step(*A->m, *Cv.v, *Dv.v, relax);
		Cv.back(); return 0;
	});
}
oracle_mat* oracle_mat_script(int64_t nops, const double* ops, unsigned char* isolated, int64_t isolated_cap)
{
	try {
		std::vector<unsigned char> iso;
		oracle_mat* T = new oracle_mat; T->m.reset(BK().matrix_script(nops, ops, iso));
		for (size_t i = 0; i < iso.size() && (int64_t)i < isolated_cap; ++i) isolated[i] = iso[i];
		return T;
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
oracle_mat* oracle_ilu_factorize(const oracle_mat* A, double beta, double sort_eps)
{
	try { oracle_mat* T = new oracle_mat; T->m.reset(BK().ilu_factorize(*A->m, beta, sort_eps)); return T; }
	catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
int oracle_ilu_apply(const oracle_mat* LU, double inv_eps, double* c, const double* d)
{
	return guard([&] {
		const int vb = LU->m->block;
		// ILU::applyLU without ordering (ilu.h:593-599): h = L^-1 d ; c = U^-1 h
		std::vector<double> h((size_t)LU->m->nrows * vb, 0.0);
		VIO Hv(LU->m->nrows * vb, vb, h.data(), h.data()), Dv(LU->m->nrows * vb, vb, d), Cv(LU->m->nrows * vb, vb, c, c);
		BK().ilu_invert_L(*LU->m, *Hv.v, *Dv.v);
		const bool ok = BK().ilu_invert_U(*LU->m, *Cv.v, *Hv.v, inv_eps);
		Cv.back();
		return ok ? 0 : 1;
	});
}
int oracle_cuthill_mckee(const oracle_mat* A, int reverse, int preserve_consec, int64_t* new_index)
{
	return guard([&] {
		std::vector<size_t> ni;
		BK().cuthill_mckee(*A->m, reverse != 0, preserve_consec != 0, ni);
		for (size_t i = 0; i < ni.size(); ++i) new_index[i] = (int64_t)ni[i];
		return 0;
	});
}
int oracle_lu_solve(const oracle_mat* A, double* x, const double* b)
{
	return guard([&] {
		const int vb = A->m->block;
		std::unique_ptr<DenseLU> lu(BK().lu_init(*A->m));
		if (!lu) throw std::runtime_error("LU: matrix is singular");
		VIO X(A->m->nrows * vb, vb, nullptr, x), B(A->m->nrows * vb, vb, b);
		BK().lu_apply(*lu, *X.v, *B.v); X.back(); return 0;
	});
}

oracle_solver* oracle_solver_create(const oracle_solver_desc* d)
{
	try {
		std::unique_ptr<oracle_solver> s(new oracle_solver);
		s->d = *d;
		std::unique_ptr<LinearIterator> pc;
		if (d->precond == ORACLE_PRECOND_GMG) {
			GMG* g = new GMG(BK());
			pc.reset(g); s->gmg = g;
			g->baseLev = d->base_lev; g->topLev = d->top_lev;
			g->cycleType = d->cycle; g->numPreSmooth = d->nu1; g->numPostSmooth = d->nu2;
			g->smootherProto.reset(make_smoother(d->smoother, d->smoother_damp, d->ilu_beta));
			if (d->base_solver == ORACLE_SOLVER_LU) g->baseSolver.reset(new LU(BK()));
			else if (d->base_solver == ORACLE_SOLVER_CG) {
				CG* c = new CG(BK());
				c->conv.maxSteps = d->base_max_steps; c->conv.minDefect = d->base_min_defect;
				c->conv.relReduction = d->base_rel_reduction;
				g->baseSolver.reset(c);
			} else throw std::runtime_error("unsupported base solver");
			g->lev.resize(d->top_lev - d->base_lev + 1);
		} else if (d->precond != ORACLE_PRECOND_NONE) pc.reset(make_smoother(d->precond, d->damp, d->ilu_beta));
		s->precond = pc.get();
		PrecondInverse* pi = nullptr;
		switch (d->solver) {
			case ORACLE_SOLVER_CG: pi = new CG(BK()); break;
			case ORACLE_SOLVER_BICGSTAB: { BiCGStab* b = new BiCGStab(BK()); b->numRestarts = d->restart > 0 ? d->restart : 0; pi = b; break; }
			case ORACLE_SOLVER_LINEAR: pi = new LinearSolver(BK()); break;
			case ORACLE_SOLVER_GMRES: { GMRES* g = new GMRES(BK()); g->restart = (size_t)(d->restart > 0 ? d->restart : 5); pi = g; break; }
			case ORACLE_SOLVER_LU: s->inv.reset(new LU(BK())); s->lone = std::move(pc); break;
			default: throw std::runtime_error("unknown solver");
		}
		if (pi) { pi->precond = std::move(pc); s->inv.reset(pi); }
		s->inv->conv.maxSteps = d->max_steps; s->inv->conv.minDefect = d->min_defect;
		s->inv->conv.relReduction = d->rel_reduction;
		return s.release();
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void oracle_solver_destroy(oracle_solver* s) { delete s; }
void oracle_set_reduction_mode(int mode) { oracle::set_reduction_mode(mode); }
int oracle_reduction_mode(void) { return oracle::reduction_mode(); }

int oracle_solver_set_level(oracle_solver* s, int lev, const oracle_mat* A, const oracle_mat* P, const oracle_mat* R)
{
	return guard([&] {
		if (!s->gmg) throw std::runtime_error("solver has no GMG preconditioner");
		s->gmg->set_level(lev, A->m.get(), P ? P->m.get() : nullptr, R ? R->m.get() : nullptr);
		return 0;
	});
}
int oracle_solver_set_smoother_matrix(oracle_solver* s, int lev, const oracle_mat* S)
{
	return guard([&] {
		if (!s->gmg) throw std::runtime_error("solver has no GMG preconditioner");
		s->gmg->set_level_smoother_matrix(lev, S->m.get());
		return 0;
	});
}
int oracle_solver_set_surface_map(oracle_solver* s, int64_t n, const int* surf_index_of_level_index)
{
	return guard([&] {
		if (!s->gmg) throw std::runtime_error("solver has no GMG preconditioner");
		s->gmg->surfMap.assign(surf_index_of_level_index, surf_index_of_level_index + n);
		return 0;
	});
}
int oracle_solver_set_precond_matrix(oracle_solver* s, const oracle_mat* M) { s->precond_matrix = M; return 0; }
int oracle_solver_init(oracle_solver* s, const oracle_mat* A)
{
	return guard([&] {
		if (!s->inv->init(*A->m)) throw std::runtime_error("solver init failed");
		if (s->lone && !s->lone->init(*A->m)) throw std::runtime_error("preconditioner init failed");
		if (s->precond_matrix && s->precond && !s->gmg && !s->precond->init(*s->precond_matrix->m))
			throw std::runtime_error("preconditioner init failed");
		return 0;
	});
}
int oracle_solver_apply(oracle_solver* s, double* x, const double* b, int vb)
{
	return guard([&] {
		const Mat& A = *s->inv->A;
		VIO X(A.ncols * vb, vb, x, x), B(A.nrows * vb, vb, b);
		const bool ok = s->inv->apply(*X.v, *B.v);
		X.back();
		return ok ? 0 : 1;
	});
}
int oracle_solver_steps(const oracle_solver* s) { return s->inv->conv.step(); }
int oracle_solver_history(const oracle_solver* s, double* out, int cap)
{
	const std::vector<double>& h = s->inv->conv.history;
	const int n = (int)h.size() < cap ? (int)h.size() : cap;
	for (int i = 0; i < n; ++i) out[i] = h[i];
	return n;
}
int oracle_precond_apply(oracle_solver* s, double* c, const double* d, int vb)
{
	return guard([&] {
		if (!s->precond) throw std::runtime_error("no preconditioner configured");
		const Mat& A = *s->inv->A;
		VIO Cv(A.ncols * vb, vb, nullptr, c), Dv(A.nrows * vb, vb, d);
		const bool ok = s->precond->apply(*Cv.v, *Dv.v);
		Cv.back();
		return ok ? 0 : 1;
	});
}

} // extern "C"
