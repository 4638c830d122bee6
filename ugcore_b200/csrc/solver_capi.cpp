// solver_capi.cpp — descriptor-level C ABI (include/ug4b200_solver.h) over the host-side
// mirror of ugcore's operator API.  Plays the role of util.solver.CreateSolver /
// SolveLinearProblem (scripts/util/solver_util.lua:602, :1182-1210) for non-Lua callers.
#include "../../include/ug4b200_solver.h"
#include "host/multigrid.h"
#include "host/matrix_io.h"
#include <cstring>

using namespace ug;

namespace {
thread_local std::string g_err;

struct SolverBase {
	virtual ~SolverBase() {}
	virtual void set_matrix(int64_t nr, int64_t nc, const int64_t* rp, const int* ci, const double* va) = 0;
	virtual void set_level(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t ncoarse,
	                       const int64_t* prp, const int* pci, const double* pva, const int64_t* rrp, const int* rci,
	                       const double* rva) = 0;
	virtual void set_coloring(int lev, int64_t n, const int* perm, int ncolors, const int64_t* cp) = 0;
	virtual void set_surface_map(int64_t n, const int* map) = 0;
	virtual void set_debug_dir(const char* dir, const double* positions, int64_t npos, int dim, int precision) = 0;
	virtual void set_layouts(int lev, int nneigh, const int* ranks, const int64_t* ptr, const int* idx, int64_t nlocal) = 0;
	virtual void set_smoother_matrix(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va) = 0;
	virtual void set_gathered_base(int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t nlocal, const int* l2g) = 0;
	virtual void set_gathered_level(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t ncoarse,
	                                const int64_t* prp, const int* pci, const double* pva, const int64_t* rrp, const int* rci,
	                                const double* rva) = 0;
	virtual void init() = 0;
	virtual int apply_host(double* x, const double* b, bool zero_guess = false) = 0;
	virtual int apply_device(double* x, const double* b) = 0;
	virtual void precond_apply(double* c, const double* d) = 0;
	virtual int64_t num_dofs() const = 0;
	std::vector<double> history;
	int steps = 0;
	double defect = 0;
};

template <typename TAlgebra>
struct SolverImpl : SolverBase {
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef MatrixOperator<matrix_type, vector_type> matop_t;
	enum { B = TAlgebra::blockSize };

	ug4b200_solver_desc d;
	SmartPtr<matop_t> A;
	SmartPtr<ILinearOperatorInverse<vector_type> > inv;
	SmartPtr<ILinearIterator<vector_type> > precond;
	SmartPtr<AssembledMultiGridCycle<TAlgebra> > gmg;
	SmartPtr<AssembledMultiGridCycle<TAlgebra> > gatheredCycle; // serial cycle below the gathered level (partitioned runs)
	SmartPtr<StdConvCheck<vector_type> > conv;
	std::map<int, std::pair<std::vector<int>, std::vector<int64_t> > > coloring;
	std::map<int, SmartPtr<GPUAlgebraLayouts> > layouts;
	std::map<int, SmartPtr<matrix_type> > smootherMatrix;   // partitioned Gauss-Seidel: consistent level matrices
	vector_type x, b;

	explicit SolverImpl(const ug4b200_solver_desc& desc) : d(desc)
	{
		conv = make_sp<StdConvCheck<vector_type> >(d.max_steps, d.min_defect, d.rel_reduction, false);
		if (d.precond == UG4B200_PRECOND_GMG) {
			const int gatherLev = d.gather_lev > d.base_lev ? d.gather_lev : d.base_lev;
			if (gatherLev > d.base_lev && d.cycle != 1) UG_THROW("a gathered cycle below the partitioned levels needs a V-cycle");
			if (gatherLev >= d.top_lev && gatherLev > d.base_lev) UG_THROW("gather level must lie below the top level");
			gmg = make_cycle(gatherLev, d.top_lev);
			if (gatherLev > d.base_lev) {
				// partitioned levels gatherLev+1..top; levels base..gatherLev replicated on every rank
				gatheredCycle = make_cycle(d.base_lev, gatherLev);
				gatheredCycle->set_base_solver(make_base_solver());
				gmg->set_base_solver(make_sp<CycleAsBaseSolver<TAlgebra> >(gatheredCycle));
			} else gmg->set_base_solver(make_base_solver());
			precond = gmg;
		} else if (d.precond != UG4B200_PRECOND_NONE) precond = make_smoother(d.precond, d.damp);
		switch (d.solver) {
			case UG4B200_SOLVER_CG: {
				SmartPtr<CG<vector_type> > s = make_sp<CG<vector_type> >();
				s->set_device_resident(!(d.flags & UG4B200_FLAG_HOST_SCALARS));
				s->set_use_graph(!(d.flags & UG4B200_FLAG_NO_GRAPH));
				s->set_preconditioner(precond); inv = s; break;
			}
			case UG4B200_SOLVER_BICGSTAB: {
				SmartPtr<BiCGStab<vector_type> > s = make_sp<BiCGStab<vector_type> >();
				s->set_restart(d.restart > 0 ? d.restart : 0);
				s->set_device_resident(!(d.flags & UG4B200_FLAG_HOST_SCALARS));   // device-resident unless the reference-shaped host loop is asked for
				s->set_use_graph(!(d.flags & UG4B200_FLAG_NO_GRAPH));
				s->set_preconditioner(precond); inv = s; break;
			}
			case UG4B200_SOLVER_LINEAR: {
				SmartPtr<LinearSolver<vector_type> > s = make_sp<LinearSolver<vector_type> >();
				s->set_device_resident(!(d.flags & UG4B200_FLAG_HOST_SCALARS));
				s->set_use_graph(!(d.flags & UG4B200_FLAG_NO_GRAPH));
				s->set_preconditioner(precond); inv = s; break;
			}
			case UG4B200_SOLVER_GMRES: { SmartPtr<GMRES<vector_type> > s = make_sp<GMRES<vector_type> >((size_t)(d.restart > 0 ? d.restart : 5)); s->set_preconditioner(precond); inv = s; break; }
			case UG4B200_SOLVER_LU: inv = make_sp<LU<TAlgebra> >(); break;
			case UG4B200_SOLVER_COARSE_CG: inv = make_sp<CoarseCG<TAlgebra> >(); break;
			default: UG_THROW("unknown solver " << d.solver);
		}
		inv->set_convergence_check(conv);
	}

	SmartPtr<AssembledMultiGridCycle<TAlgebra> > make_cycle(int base, int top)
	{
		SmartPtr<AssembledMultiGridCycle<TAlgebra> > g = make_sp<AssembledMultiGridCycle<TAlgebra> >();
		g->set_base_level(base); g->set_surface_level(top);
		g->set_cycle_type(d.cycle); g->set_num_presmooth(d.nu1); g->set_num_postsmooth(d.nu2);
		g->set_smoother(make_smoother(d.smoother, d.smoother_damp));
		g->set_fuse_jacobi(!(d.flags & UG4B200_FLAG_NO_FUSED_JACOBI));
		g->set_compute_final_level_defect((d.flags & UG4B200_FLAG_FINAL_LEVEL_DEFECT) != 0);
		g->set_rap((d.flags & UG4B200_FLAG_RAP) != 0);
		return g;
	}
	SmartPtr<ILinearOperatorInverse<vector_type> > make_base_solver()
	{
		if (d.base_solver == UG4B200_SOLVER_LU) return make_sp<LU<TAlgebra> >();
		if (d.base_solver == UG4B200_SOLVER_COARSE_CG || d.base_solver == UG4B200_SOLVER_CG) {
			SmartPtr<CoarseCG<TAlgebra> > cg = make_sp<CoarseCG<TAlgebra> >();
			cg->set_convergence_check(make_sp<StdConvCheck<vector_type> >(d.base_max_steps, d.base_min_defect, d.base_rel_reduction, false));
			return cg;
		}
		UG_THROW("unsupported base solver " << d.base_solver);
	}
	SmartPtr<ILinearIterator<vector_type> > make_smoother(int kind, double damp)
	{
		switch (kind) {
			case UG4B200_PRECOND_JACOBI: return make_sp<Jacobi<TAlgebra> >(damp);
			case UG4B200_PRECOND_GS: { SmartPtr<GaussSeidel<TAlgebra> > g = make_sp<GaussSeidel<TAlgebra> >(); g->set_sor_relax(damp); return g; }
			case UG4B200_PRECOND_BGS: { SmartPtr<BackwardGaussSeidel<TAlgebra> > g = make_sp<BackwardGaussSeidel<TAlgebra> >(); g->set_sor_relax(damp); return g; }
			case UG4B200_PRECOND_SGS: { SmartPtr<SymmetricGaussSeidel<TAlgebra> > g = make_sp<SymmetricGaussSeidel<TAlgebra> >(); g->set_sor_relax(damp); return g; }
			case UG4B200_PRECOND_ILU: {
				SmartPtr<ILU<TAlgebra> > g = make_sp<ILU<TAlgebra> >(d.ilu_beta);
				if (d.ilu_order == UG4B200_ILU_ORDER_CMK) g->set_sort(true);
				else if (d.ilu_order == UG4B200_ILU_ORDER_MULTICOLOR) g->set_multicolor_ordering(true);
				else if (d.ilu_order != UG4B200_ILU_ORDER_NATURAL) UG_THROW("unknown ILU ordering " << d.ilu_order);
				return g;
			}
		}
		UG_THROW("unknown smoother / preconditioner kind " << kind);
	}

	void set_matrix(int64_t nr, int64_t nc, const int64_t* rp, const int* ci, const double* va) override
	{
		A = make_sp<matop_t>();
		A->set_from_crs((size_t)nr, (size_t)nc, rp, ci, va);
	}
	void set_level(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t ncoarse,
	               const int64_t* prp, const int* pci, const double* pva, const int64_t* rrp, const int* rci,
	               const double* rva) override
	{
		if (!gmg) UG_THROW("solver has no GMG preconditioner");
		if (rp) {
			SmartPtr<matop_t> Al = make_sp<matop_t>();
			Al->set_from_crs((size_t)nrows, (size_t)nrows, rp, ci, va);
			gmg->set_level_operator(lev, Al);
		}
		if (prp) {
			SmartPtr<GPUTransferMatrix> P = make_sp<GPUTransferMatrix>(), R;
			P->set_from_crs((size_t)nrows, (size_t)ncoarse, prp, pci, pva);
			if (rrp) { R = make_sp<GPUTransferMatrix>(); R->set_from_crs((size_t)ncoarse, (size_t)nrows, rrp, rci, rva); }
			gmg->set_level_transfer(lev, P, R);
		}
	}
	void set_debug_dir(const char* dir, const double* positions, int64_t npos, int dim, int precision) override
	{
		if (!dir) { inv->set_debug(SmartPtr<IVectorDebugWriter<vector_type> >()); return; }
		IOPositions p; p.dim = dim; p.resize((size_t)(npos > 0 ? npos : 0));
		if (positions && npos > 0) std::memcpy(p.xyz.data(), positions, sizeof(double) * 3 * (size_t)npos);
		if (!positions) p.resize((size_t)num_dofs());
		inv->set_debug(make_sp<ConnectionViewerVectorWriter<vector_type> >(std::string(dir), p, dim, precision));
	}
	void set_surface_map(int64_t n, const int* map) override
	{
		if (!gmg) UG_THROW("solver has no GMG preconditioner");
		std::vector<char> seen((size_t)n, 0);
		for (int64_t i = 0; i < n; ++i) { if (map[i] < 0 || map[i] >= n || seen[(size_t)map[i]]) UG_THROW("surface map is not a permutation"); seen[(size_t)map[i]] = 1; }
		gmg->set_surface_to_level_map(std::vector<int>(map, map + n));
	}
	void set_coloring(int lev, int64_t n, const int* perm, int ncolors, const int64_t* cp) override
	{
		coloring[lev] = std::make_pair(std::vector<int>(perm, perm + n), std::vector<int64_t>(cp, cp + ncolors + 1));
		if (lev >= 0) UG_THROW("per-level colourings: greedy colouring is used inside GMG this round");
		GaussSeidelBase<TAlgebra>* g = dynamic_cast<GaussSeidelBase<TAlgebra>*>(precond.get());
		if (!g) UG_THROW("set_coloring: preconditioner is not a Gauss-Seidel sweep");
		g->set_coloring(coloring[lev].first, coloring[lev].second);
	}
	void set_layouts(int lev, int nneigh, const int* ranks, const int64_t* ptr, const int* idx, int64_t nlocal) override
	{
		SmartPtr<GPUAlgebraLayouts> l = make_sp<GPUAlgebraLayouts>(nneigh, ranks, ptr, idx, nlocal, GPUManager::proc_rank());
		layouts[lev] = l;
		if (gmg) gmg->set_level_layouts(lev, l);
	}
	void set_smoother_matrix(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va) override
	{
		SmartPtr<matrix_type> Ac = make_sp<matrix_type>();
		Ac->set_from_crs((size_t)nrows, (size_t)nrows, rp, ci, va);
		smootherMatrix[lev] = Ac;
		if (gmg) gmg->set_level_smoother_matrix(lev, Ac);
	}
	void set_gathered_base(int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t nlocal, const int* l2g) override
	{
		if (!gmg) UG_THROW("solver has no GMG preconditioner");
		SmartPtr<matop_t> G = make_sp<matop_t>();
		G->set_from_crs((size_t)nrows, (size_t)nrows, rp, ci, va);
		gmg->set_gathered_base(G, std::vector<int>(l2g, l2g + nlocal));
	}
	void set_gathered_level(int lev, int64_t nrows, const int64_t* rp, const int* ci, const double* va, int64_t ncoarse,
	                        const int64_t* prp, const int* pci, const double* pva, const int64_t* rrp, const int* rci,
	                        const double* rva) override
	{
		if (!gatheredCycle) UG_THROW("solver has no gathered cycle (desc.gather_lev <= base_lev)");
		if (rp) {
			SmartPtr<matop_t> Al = make_sp<matop_t>();
			Al->set_from_crs((size_t)nrows, (size_t)nrows, rp, ci, va);
			gatheredCycle->set_level_operator(lev, Al);
		}
		if (prp) {
			SmartPtr<GPUTransferMatrix> P = make_sp<GPUTransferMatrix>(), R;
			P->set_from_crs((size_t)nrows, (size_t)ncoarse, prp, pci, pva);
			if (rrp) { R = make_sp<GPUTransferMatrix>(); R->set_from_crs((size_t)ncoarse, (size_t)nrows, rrp, rci, rva); }
			gatheredCycle->set_level_transfer(lev, P, R);
		}
	}
	SmartPtr<GPUAlgebraLayouts> top_layouts()
	{
		if (layouts.empty()) return SmartPtr<GPUAlgebraLayouts>();
		return layouts.rbegin()->second;
	}
	void init() override
	{
		if (!A) UG_THROW("solver: matrix not set");
		Jacobi<TAlgebra>* j = dynamic_cast<Jacobi<TAlgebra>*>(precond.get());
		if (j) j->set_layouts(top_layouts());
		GaussSeidelBase<TAlgebra>* g = dynamic_cast<GaussSeidelBase<TAlgebra>*>(precond.get());
		if (g && top_layouts()) {
			g->set_layouts(top_layouts());
			if (!smootherMatrix.empty()) g->set_consistent_matrix(smootherMatrix.rbegin()->second);
		}
		ILU<TAlgebra>* ilu = dynamic_cast<ILU<TAlgebra>*>(precond.get());
		if (ilu && top_layouts()) {
			ilu->set_layouts(top_layouts());
			if (!smootherMatrix.empty()) ilu->set_consistent_matrix(smootherMatrix.rbegin()->second);
		}
		x.create(A->num_cols()); b.create(A->num_rows());
		x.set_layouts(top_layouts()); b.set_layouts(top_layouts());
		if (!inv->init(A, x)) UG_THROW("solver init failed");
		if (d.solver == UG4B200_SOLVER_LU || d.solver == UG4B200_SOLVER_COARSE_CG) {
			if (precond && !precond->init(A, x)) UG_THROW("preconditioner init failed");
		}
	}
	int finish(bool ok)
	{
		steps = conv->step(); defect = conv->defect(); history = conv->get_defects();
		return ok ? 0 : 1;
	}
	int apply_host(double* xh, const double* bh, bool zero_guess) override
	{
		ug4b200_ctx* c = GPUManager::ctx();
		if (zero_guess) x.set(0.0);
		else UG_GPU_CHECK(ug4b200_h2d(c, x.dev(), xh, x.len() * sizeof(double)));
		UG_GPU_CHECK(ug4b200_h2d(c, b.dev(), bh, b.len() * sizeof(double)));
		x.set_storage_type(PST_CONSISTENT); b.set_storage_type(PST_ADDITIVE);
		const bool ok = inv->apply(x, b);
		UG_GPU_CHECK(ug4b200_d2h(c, xh, x.dev(), x.len() * sizeof(double)));
		return finish(ok);
	}
	int apply_device(double* xd, const double* bd) override
	{
		ug4b200_ctx* c = GPUManager::ctx();
		UG_GPU_CHECK(ug4b200_vec_copy(c, x.len(), x.dev(), xd));
		UG_GPU_CHECK(ug4b200_vec_copy(c, b.len(), b.dev(), bd));
		x.set_storage_type(PST_CONSISTENT); b.set_storage_type(PST_ADDITIVE);
		const bool ok = inv->apply(x, b);
		UG_GPU_CHECK(ug4b200_vec_copy(c, x.len(), xd, x.dev()));
		UG_GPU_CHECK(ug4b200_batch_flush(c));   // the caller owns xd and the stream from here on
		return finish(ok);
	}
	void precond_apply(double* ch, const double* dh) override
	{
		if (!precond) UG_THROW("no preconditioner configured");
		ug4b200_ctx* c = GPUManager::ctx();
		UG_GPU_CHECK(ug4b200_h2d(c, b.dev(), dh, b.len() * sizeof(double)));
		b.set_storage_type(PST_ADDITIVE);
		if (!precond->apply(x, b)) UG_THROW("preconditioner apply failed");
		UG_GPU_CHECK(ug4b200_d2h(c, ch, x.dev(), x.len() * sizeof(double)));
	}
	int64_t num_dofs() const override { return A ? (int64_t)A->num_rows() * B : 0; }
};

template <class F> int guard(F f)
{
	try { return f(); }
	catch (const std::exception& e) { g_err = e.what(); return -1; }
	catch (...) { g_err = "unknown exception"; return -1; }
}
} // namespace

struct ug4b200_solver { std::unique_ptr<SolverBase> p; };

extern "C" {

int ug4b200_host_init(int device, void* stream) { return guard([&] { GPUManager::init(device, stream); return 0; }); }
int ug4b200_host_finalize(void) { return guard([&] { GPUManager::finalize(); return 0; }); }
ug4b200_ctx* ug4b200_host_ctx(void) { return GPUManager::ctx_or_null(); }
const char* ug4b200_host_last_error(void) { return g_err.c_str(); }
int ug4b200_host_comm_init(int nranks, int rank, const unsigned char id[UG4B200_NCCL_ID_BYTES])
{ return guard([&] { UG_GPU_CHECK(ug4b200_comm_init(GPUManager::ctx(), nranks, rank, id)); GPUManager::set_procs(nranks, rank); return 0; }); }

int ug4b200_solver_create(const ug4b200_solver_desc* d, ug4b200_solver** out)
{
	*out = nullptr;
	return guard([&] {
		std::unique_ptr<ug4b200_solver> s(new ug4b200_solver);
		if (d->block == 1) s->p.reset(new SolverImpl<GPUAlgebra>(*d));
		else if (d->block == 2) s->p.reset(new SolverImpl<GPUBlockAlgebra<2> >(*d));
		else if (d->block == 3) s->p.reset(new SolverImpl<GPUBlockAlgebra<3> >(*d));
		else UG_THROW("block size must be 1, 2 or 3");
		*out = s.release();
		return 0;
	});
}
int ug4b200_solver_destroy(ug4b200_solver* s) { return guard([&] { delete s; return 0; }); }
int ug4b200_solver_set_matrix(ug4b200_solver* s, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals)
{ return guard([&] { s->p->set_matrix(nrows, ncols, rowptr, cols, vals); return 0; }); }
int ug4b200_solver_set_level(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols, const double* vals,
                             int64_t ncoarse, const int64_t* p_rowptr, const int* p_cols, const double* p_vals,
                             const int64_t* r_rowptr, const int* r_cols, const double* r_vals)
{ return guard([&] { s->p->set_level(lev, nrows, rowptr, cols, vals, ncoarse, p_rowptr, p_cols, p_vals, r_rowptr, r_cols, r_vals); return 0; }); }
int ug4b200_solver_set_debug_dir(ug4b200_solver* s, const char* dir, const double* positions, int64_t npos, int dim, int precision)
{ return guard([&] { s->p->set_debug_dir(dir, positions, npos, dim, precision); return 0; }); }
int ug4b200_solver_set_surface_map(ug4b200_solver* s, int64_t n, const int* surf_index_of_level_index)
{ return guard([&] { s->p->set_surface_map(n, surf_index_of_level_index); return 0; }); }
int ug4b200_solver_set_coloring(ug4b200_solver* s, int lev, int64_t n, const int* perm, int ncolors, const int64_t* color_ptr)
{ return guard([&] { s->p->set_coloring(lev, n, perm, ncolors, color_ptr); return 0; }); }
int ug4b200_solver_set_layouts(ug4b200_solver* s, int lev, int nneigh, const int* neigh_rank, const int64_t* neigh_ptr, const int* indices, int64_t nlocal)
{ return guard([&] { s->p->set_layouts(lev, nneigh, neigh_rank, neigh_ptr, indices, nlocal); return 0; }); }
int ug4b200_solver_set_smoother_matrix(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols, const double* vals)
{ return guard([&] { s->p->set_smoother_matrix(lev, nrows, rowptr, cols, vals); return 0; }); }
int ug4b200_solver_set_gathered_base(ug4b200_solver* s, int64_t nrows, const int64_t* rowptr, const int* cols, const double* vals,
                                     int64_t nlocal, const int* local_to_global)
{ return guard([&] { s->p->set_gathered_base(nrows, rowptr, cols, vals, nlocal, local_to_global); return 0; }); }
int ug4b200_solver_set_gathered_level(ug4b200_solver* s, int lev, int64_t nrows, const int64_t* rowptr, const int* cols, const double* vals,
                                      int64_t ncoarse, const int64_t* p_rowptr, const int* p_cols, const double* p_vals,
                                      const int64_t* r_rowptr, const int* r_cols, const double* r_vals)
{ return guard([&] { s->p->set_gathered_level(lev, nrows, rowptr, cols, vals, ncoarse, p_rowptr, p_cols, p_vals, r_rowptr, r_cols, r_vals); return 0; }); }
int ug4b200_solver_init(ug4b200_solver* s) { return guard([&] { s->p->init(); return 0; }); }
int ug4b200_solver_apply(ug4b200_solver* s, double* x_host, const double* b_host) { return guard([&] { return s->p->apply_host(x_host, b_host); }); }
int ug4b200_solver_apply_zero_guess(ug4b200_solver* s, double* x_host, const double* b_host) { return guard([&] { return s->p->apply_host(x_host, b_host, true); }); }
int ug4b200_solver_apply_device(ug4b200_solver* s, double* x_dev, const double* b_dev) { return guard([&] { return s->p->apply_device(x_dev, b_dev); }); }
int ug4b200_solver_steps(const ug4b200_solver* s) { return s->p->steps; }
double ug4b200_solver_defect(const ug4b200_solver* s) { return s->p->defect; }
int ug4b200_solver_history(const ug4b200_solver* s, double* out, int cap)
{
	const int n = (int)s->p->history.size() < cap ? (int)s->p->history.size() : cap;
	for (int i = 0; i < n; ++i) out[i] = s->p->history[i];
	return n;
}
int ug4b200_solver_precond_apply(ug4b200_solver* s, double* c_host, const double* d_host) { return guard([&] { s->p->precond_apply(c_host, d_host); return 0; }); }
int64_t ug4b200_solver_num_dofs(const ug4b200_solver* s) { return s->p->num_dofs(); }

/* ---- import / export (host/matrix_io.h) ---- */
} // extern "C"
struct ug4b200_host_matrix { GPUSparseMatrix<double> A; IOPositions pos; int dim = 3; };
namespace {
int io_format(const char* filename, int format)
{
	if (format == 1 || format == 2) return format;
	const std::string f(filename);
	return (f.size() >= 4 && f.compare(f.size() - 4, 4, ".mtx") == 0) ? 2 : 1;
}
IOPositions io_positions(const double* positions, size_t n, int dim)
{
	IOPositions p; p.dim = dim; p.resize(n);
	if (positions) std::memcpy(p.xyz.data(), positions, sizeof(double) * 3 * n);
	return p;
}
} // namespace
extern "C" {

int ug4b200_io_read_matrix(const char* filename, int format, int keep_zeros, int64_t n_to, ug4b200_host_matrix** out)
{
	*out = nullptr;
	return guard([&] {
		std::unique_ptr<ug4b200_host_matrix> m(new ug4b200_host_matrix);
		if (io_format(filename, format) == 2) {
			MatrixIOMtx io(filename);
			io.read_into(m->A);
			m->pos.resize(0);
		} else if (!ConnectionViewer::ReadMatrix(filename, m->A, m->pos, m->dim, keep_zeros != 0, (size_t)(n_to > 0 ? n_to : 0)))
			UG_THROW("cannot open " << filename);
		*out = m.release();
		return 0;
	});
}
int ug4b200_io_matrix_info(const ug4b200_host_matrix* m, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* dim, int64_t* npos)
{
	return guard([&] {
		if (nrows) *nrows = (int64_t)m->A.num_rows();
		if (ncols) *ncols = (int64_t)m->A.num_cols();
		if (nnz) *nnz = (int64_t)m->A.total_num_connections();
		if (dim) *dim = m->dim;
		if (npos) *npos = (int64_t)m->pos.size();
		return 0;
	});
}
int ug4b200_io_matrix_export(const ug4b200_host_matrix* m, int64_t* rowptr, int* cols, double* vals, double* positions)
{
	return guard([&] {
		const std::vector<int64_t>& rp = m->A.crs_rowptr(); const std::vector<int>& ci = m->A.crs_cols(); const std::vector<double>& va = m->A.crs_vals();
		std::memcpy(rowptr, rp.data(), sizeof(int64_t) * rp.size());
		if (!ci.empty()) { std::memcpy(cols, ci.data(), sizeof(int) * ci.size()); std::memcpy(vals, va.data(), sizeof(double) * va.size()); }
		if (positions && m->pos.size()) std::memcpy(positions, m->pos.xyz.data(), sizeof(double) * m->pos.xyz.size());
		return 0;
	});
}
void ug4b200_io_matrix_free(ug4b200_host_matrix* m) { delete m; }
int ug4b200_io_write_matrix(const char* filename, int format, int64_t nrows, int64_t ncols, const int64_t* rowptr,
                            const int* cols, const double* vals, const double* positions, int dim, int from_to,
                            int precision)
{
	return guard([&] {
		GPUSparseMatrix<double> A;
		A.set_from_crs((size_t)nrows, (size_t)ncols, rowptr, cols, vals);
		if (io_format(filename, format) == 2) {
			MatrixIOMtx io(filename);
			if (precision > 0) io.set_precision(precision);
			io.write_from(A);
		} else if (from_to) {
			const IOPositions to = io_positions(positions, (size_t)nrows, dim);
			const IOPositions from = io_positions(positions ? positions + 3 * nrows : nullptr, (size_t)ncols, dim);
			if (!ConnectionViewer::WriteMatrix(std::string(filename), A, from, to, (size_t)dim, precision)) UG_THROW("WriteMatrix: positions do not match the matrix");
		} else {
			if (nrows != ncols) UG_THROW("ConnectionViewer::WriteMatrix: a rectangular matrix needs the from / to form");
			ConnectionViewer::WriteMatrix(std::string(filename), A, io_positions(positions, (size_t)nrows, dim), dim, precision);
		}
		return 0;
	});
}
int ug4b200_host_rap(int64_t nc, int64_t nf, const int64_t* r_rowptr, const int* r_cols, const double* r_vals,
                     const int64_t* a_rowptr, const int* a_cols, const double* a_vals,
                     const int64_t* p_rowptr, const int* p_cols, const double* p_vals, ug4b200_host_matrix** out)
{
	*out = nullptr;
	return guard([&] {
		GPUSparseMatrix<double> R, A, P;
		R.set_from_crs((size_t)nc, (size_t)nf, r_rowptr, r_cols, r_vals);
		A.set_from_crs((size_t)nf, (size_t)nf, a_rowptr, a_cols, a_vals);
		P.set_from_crs((size_t)nf, (size_t)nc, p_rowptr, p_cols, p_vals);
		std::unique_ptr<ug4b200_host_matrix> m(new ug4b200_host_matrix);
		m->A.resize_and_clear((size_t)nc, (size_t)nc);
		AddMultiplyOf(m->A, R, A, P);
		m->pos.resize(0);
		*out = m.release();
		return 0;
	});
}
} // extern "C"
namespace {
template <typename TAlgebra>
void apply_transposed_impl(int64_t nrows, int64_t ncols, const int64_t* rp, const int* ci, const double* va, double* y, const double* x)
{
	typename TAlgebra::matrix_type A;
	A.set_from_crs((size_t)nrows, (size_t)ncols, rp, ci, va);
	typename TAlgebra::vector_type vx((size_t)nrows), vy((size_t)ncols);
	vx.assign_from_host(x);
	if (!A.apply_transposed(vy, vx)) UG_THROW("apply_transposed failed");
	vy.copy_to_host(y);
}
} // namespace
namespace {
template <typename TAlgebra>
void vector_selftest_impl(int64_t n, unsigned seed, double from, double to, int64_t nidx, const int64_t* idx, const double* addv,
                          double* values, double* got, double* maxnorm)
{
	typedef typename TAlgebra::vector_type V;
	typedef typename V::value_type T;
	enum { B = TAlgebra::blockSize };
	V v((size_t)n);
	std::srand(seed);
	v.set_random(from, to);
	v.copy_to_host(values);
	std::vector<size_t> ind(idx, idx + nidx);
	std::vector<T> u((size_t)nidx);
	std::memcpy((void*)u.data(), addv, sizeof(double) * B * (size_t)nidx);
	v.add(u.data(), ind.data(), (size_t)nidx);          // host mirror
	v *= 2.0;                                            // device (uploads the mirror first)
	v.get(u.data(), ind.data(), (size_t)nidx);          // host mirror again (downloads)
	std::memcpy(got, (const void*)u.data(), sizeof(double) * B * (size_t)nidx);
	*maxnorm = v.maxnorm();
}
} // namespace
extern "C" {
int ug4b200_host_vector_selftest(int block, int64_t n, unsigned seed, double from, double to, int64_t nidx, const int64_t* idx,
                                 const double* add_vals, double* values_out, double* got_out, double* maxnorm_out)
{
	return guard([&] {
		if (block == 1) vector_selftest_impl<GPUAlgebra>(n, seed, from, to, nidx, idx, add_vals, values_out, got_out, maxnorm_out);
		else if (block == 3) vector_selftest_impl<GPUBlockAlgebra<3> >(n, seed, from, to, nidx, idx, add_vals, values_out, got_out, maxnorm_out);
		else UG_THROW("block size must be 1 or 3");
		return 0;
	});
}
int ug4b200_host_apply_transposed(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
                                  const double* vals, double* y_host, const double* x_host)
{
	return guard([&] {
		if (block == 1) apply_transposed_impl<GPUAlgebra>(nrows, ncols, rowptr, cols, vals, y_host, x_host);
		else if (block == 2) apply_transposed_impl<GPUBlockAlgebra<2> >(nrows, ncols, rowptr, cols, vals, y_host, x_host);
		else if (block == 3) apply_transposed_impl<GPUBlockAlgebra<3> >(nrows, ncols, rowptr, cols, vals, y_host, x_host);
		else UG_THROW("block size must be 1, 2 or 3");
		return 0;
	});
}
int ug4b200_host_matrix_script(int64_t nops, const double* ops, ug4b200_host_matrix** out)
{
	*out = nullptr;
	return guard([&] {
		typedef GPUSparseMatrix<double> M;
		std::unique_ptr<ug4b200_host_matrix> m(new ug4b200_host_matrix);
		M* A = &m->A;
		std::unique_ptr<ug4b200_host_matrix> tmp;
		for (int64_t k = 0; k < nops; ++k) {
			const int code = (int)ops[4 * k]; const size_t r = (size_t)ops[4 * k + 1], c = (size_t)ops[4 * k + 2]; const double v = ops[4 * k + 3];
			switch (code) {
				case 0: A->resize_and_clear(r, c); break;
				case 1: (*A)(r, c) = v; break;
				case 2: (*A)(r, c) += v; break;
				case 3: A->scale(v); break;
				case 4: A->clear_retain_structure(); break;
				case 5: A->resize_and_keep_values(r, c); break;
				case 6: A->defragment(); break;
				case 7: A->set(v); break;
				case 8: case 9: {
					tmp.reset(new ug4b200_host_matrix);
					if (code == 8) tmp->A.set_as_transpose_of(*A, v); else tmp->A.set_as_copy_of(*A, v);
					m.swap(tmp); A = &m->A;
					break;
				}
				case 10: { const M& cA = *A; volatile double sink = cA(r, c); (void)sink; break; }
				case 11: case 13: {
					std::vector<M::connection> row(c);
					for (size_t t = 0; t < c; ++t) {
						if (k + 1 + (int64_t)t >= nops || (int)ops[4 * (k + 1 + t)] != 12) UG_THROW("matrix script: row entries (code 12) missing");
						row[t].iIndex = (size_t)ops[4 * (k + 1 + t) + 2]; row[t].dValue = ops[4 * (k + 1 + t) + 3];
					}
					if (code == 11) A->set_matrix_row(r, row.data(), c); else A->add_matrix_row(r, row.data(), c);
					k += (int64_t)c;
					break;
				}
				default: UG_THROW("matrix script: unknown operation " << code);
			}
		}
		m->pos.resize(0);
		*out = m.release();
		return 0;
	});
}
int ug4b200_host_matrix_isolated(const ug4b200_host_matrix* m, unsigned char* isolated)
{
	return guard([&] { for (size_t i = 0; i < m->A.num_rows(); ++i) isolated[i] = m->A.is_isolated(i) ? 1 : 0; return 0; });
}
int ug4b200_host_ilu_factorize_block(int block, int64_t n, const int64_t* rowptr, const int* cols, double* vals, double beta, double sort_eps)
{
	return guard([&] {
		if (block < 1 || block > 3) UG_THROW("ILU: block size must be 1, 2 or 3");
		std::vector<int64_t> rp(rowptr, rowptr + n + 1);
		std::vector<int> ci(cols, cols + rp[(size_t)n]);
		std::vector<double> va(vals, vals + rp[(size_t)n] * block * block);
		for (int64_t i = 0; i < n; ++i)
			for (int64_t p = rp[(size_t)i] + 1; p < rp[(size_t)i + 1]; ++p)
				if (ci[(size_t)p - 1] >= ci[(size_t)p]) UG_THROW("ILU: columns of row " << i << " are not sorted");
		FactorizeILU(block, n, rp, ci, va, beta, sort_eps);
		std::memcpy(vals, va.data(), sizeof(double) * va.size());
		return 0;
	});
}
int ug4b200_host_ilu_factorize(int64_t n, const int64_t* rowptr, const int* cols, double* vals, double beta, double sort_eps)
{ return ug4b200_host_ilu_factorize_block(1, n, rowptr, cols, vals, beta, sort_eps); }
int ug4b200_host_level_sets(int64_t n, const int64_t* rowptr, const int* cols, int lower, int* level, int* nlevels)
{
	return guard([&] {
		std::vector<int64_t> rp(rowptr, rowptr + n + 1);
		std::vector<int> ci(cols, cols + rp[(size_t)n]);
		std::vector<int> lev;
		*nlevels = level_sets(n, rp, ci, lower != 0, lev);
		for (int64_t i = 0; i < n; ++i) level[i] = lev[(size_t)i];
		return 0;
	});
}
int ug4b200_host_cuthill_mckee(int64_t n, const int64_t* rowptr, const int* cols, int reverse, int preserve_consec,
                               int64_t* new_index)
{
	return guard([&] {
		std::vector<size_t> ni;
		GetCuthillMcKeeOrder(n, rowptr, cols, ni, reverse != 0, preserve_consec != 0);
		for (size_t i = 0; i < ni.size(); ++i) new_index[i] = (int64_t)ni[i];
		return 0;
	});
}
int ug4b200_io_vector_size(const char* filename, int64_t* n, int* dim)
{
	return guard([&] {
		std::fstream f(filename, std::ios::in);
		if (!f.is_open()) UG_THROW("cannot open " << filename);
		int version = -1, d = -1; long long g = -1;
		f >> version >> d >> g;
		if (!f || version != 1 || g < 0) UG_THROW(filename << " is not a version-1 ConnectionViewer file");
		*n = g; if (dim) *dim = d;
		return 0;
	});
}
int ug4b200_io_read_vector(const char* filename, int64_t n, double* values, double* positions)
{
	return guard([&] {
		std::vector<double> v; IOPositions pos; int dim = 0;
		if (!ConnectionViewer::ReadVector(filename, v, pos, dim)) UG_THROW("cannot open " << filename);
		if ((int64_t)v.size() != n) UG_THROW("ReadVector: " << filename << " holds " << v.size() << " entries, caller expects " << n);
		if (n) std::memcpy(values, v.data(), sizeof(double) * v.size());
		if (positions && n) std::memcpy(positions, pos.xyz.data(), sizeof(double) * pos.xyz.size());
		return 0;
	});
}
int ug4b200_io_write_vector(const char* filename, int64_t n, const double* values, const double* positions, int dim,
                            int precision)
{
	return guard([&] {
		ConnectionViewer::WriteVector(filename, values, (size_t)n, io_positions(positions, (size_t)n, dim), dim, precision);
		return 0;
	});
}

} // extern "C"
