// multigrid.h — StdTransfer and AssembledMultiGridCycle for the GPU algebra.
//
//   StdTransfer              ugbase/lib_disc/operator/linear_operator/std_transfer.h:56-59,
//                            std_transfer_impl.h:719-772 (prolongate), :774-806 (do_restrict)
//   AssembledMultiGridCycle  ugbase/lib_disc/operator/linear_operator/multi_grid_solver/
//                            mg_solver.h:81-84, mg_solver_impl.hpp:174-275 (apply),
//                            :1685-1816 (presmooth_and_restriction), :1818-1964
//                            (prolongation_and_postsmooth), :1967-2086 (base_solve),
//                            :2089-2136 (lmgc)
//
// In ugcore these classes are templated on <TDomain, TAlgebra> and pull level operators
// and P from the ApproximationSpace; assembly stays on the CPU, so here
// the assembled level matrices and transfer matrices are handed in after assembly
// (set_level_operator / set_level_transfer — the hook points named in SURVEY.md §3.2).
// The cycle itself never touches host memory: every level vector is device resident and
// the surface<->level copies (mg_solver_impl.hpp:211-217, 244-248) are device gathers.
#pragma once
#include "solvers.h"
#include "sparse_util.h"

namespace ug {

/// transfer matrices carry scalars; for block algebras ugcore stores the scalar on the block
/// diagonal (DoFRef(P, ..)), which acts component-wise — done by the kernel instead
typedef GPUSparseMatrix<double> GPUTransferMatrix;

template <typename TAlgebra>
class StdTransfer {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	StdTransfer() : m_dampProl(1.0), m_dampRes(1.0), m_bUseTransposed(true) {}
	void set_prolongation_damping(number damp) { m_dampProl = damp; }
	void set_restriction_damping(number damp) { m_dampRes = damp; }
	void set_use_transposed(bool b) { m_bUseTransposed = b; }
	/// std_transfer.h: the P1 fast path of the CPU-side assembly of P (std_transfer_impl.h:46-168) — assembly is upstream
	void enable_p1_lagrange_optimization(bool) {}
	template <typename TWriter> void set_debug(SmartPtr<TWriter>) {}
	/// cached matrices of one level pair (std_transfer.h:195-211); R may be null -> R = P^T
	void set_matrices(SmartPtr<GPUTransferMatrix> P, SmartPtr<GPUTransferMatrix> R) { m_P = P; m_R = R; }
	void init()
	{
		if (!m_P) UG_THROW("StdTransfer: prolongation matrix not set");
		if (!m_R) {
			if (!m_bUseTransposed) UG_THROW("StdTransfer: restriction matrix not set");
			m_R = make_sp<GPUTransferMatrix>();
			m_R->set_as_transpose_of(*m_P); // std_transfer_impl.h:694-695
		}
		m_P->device(); m_R->device();
	}
	number restriction_damping() const { return m_dampRes; }
	SmartPtr<GPUTransferMatrix> prolongation() { return m_P; }
	SmartPtr<GPUTransferMatrix> restriction() { return m_R; }
	/// uFine = dampProl * P * uCoarse   (std_transfer_impl.h:738-740)
	void prolongate(vector_type& uFine, const vector_type& uCoarse)
	{
		m_P->axpy(uFine, 0.0, uFine, m_dampProl, uCoarse);
		uFine.set_storage_type(uCoarse.get_storage_mask());
	}
	/// uCoarse = dampRes * R * uFine, rows without connections untouched (std_transfer_impl.h:791-792)
	void do_restrict(vector_type& uCoarse, const vector_type& uFine)
	{
		m_R->apply_ignore_zero_rows(uCoarse, m_dampRes, uFine);
		uCoarse.set_storage_type(uFine.get_storage_mask());
	}
	SmartPtr<StdTransfer> clone() { SmartPtr<StdTransfer> t(new StdTransfer()); t->m_dampProl = m_dampProl; t->m_dampRes = m_dampRes; t->m_bUseTransposed = m_bUseTransposed; return t; }
  protected:
	number m_dampProl, m_dampRes;
	bool m_bUseTransposed;
	SmartPtr<GPUTransferMatrix> m_P, m_R;
};

enum { _V_ = 1, _W_ = 2, _F_ = -1 };

template <typename TAlgebra> class AssembledMultiGridCycle;

/// One multigrid cycle as "base solver" of a partitioned hierarchy: below the gathered level every
/// rank holds the whole (small) grid and runs the serial cycle redundantly instead of exchanging
/// interface values of a few hundred DoFs per smoothing step.  This is ugcore's gathered base
/// solve (mg_solver_impl.hpp:2003-2070) with the coarse levels kept inside the same V-cycle:
/// lmgc(l) on the gathered level l with sc_l = 0 is exactly apply() of a cycle whose top level is l,
/// so the result equals the fully partitioned cycle up to summation order (V-cycles only).
template <typename TAlgebra>
class CycleAsBaseSolver : public ILinearOperatorInverse<typename TAlgebra::vector_type> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	explicit CycleAsBaseSolver(SmartPtr<AssembledMultiGridCycle<TAlgebra> > cycle) : m_spCycle(cycle) {}
	virtual const char* name() const { return "GatheredCycle"; }
	virtual bool supports_parallel() const { return false; }
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > L);
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > J, const vector_type&) { return init(J); }
	virtual bool apply(vector_type& u, const vector_type& f);
	virtual bool apply_return_defect(vector_type& u, vector_type& f);
	SmartPtr<AssembledMultiGridCycle<TAlgebra> > cycle() { return m_spCycle; }
  protected:
	SmartPtr<AssembledMultiGridCycle<TAlgebra> > m_spCycle;
	SmartPtr<ILinearOperator<vector_type> > m_spOp;
};

template <typename TAlgebra>
class AssembledMultiGridCycle : public ILinearIterator<typename TAlgebra::vector_type> {
  public:
	typedef TAlgebra algebra_type;
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef MatrixOperator<matrix_type, vector_type> matrix_operator_type;
	typedef ILinearIterator<vector_type> smoother_type;
	enum { B = TAlgebra::blockSize };

	AssembledMultiGridCycle()
	    : m_baseLev(0), m_topLev(0), m_cycleType(_V_), m_numPreSmooth(2), m_numPostSmooth(2), m_bFinalDefect(false),
	      m_bFuseJacobi(true) {}
	virtual const char* name() const { return "Geometric MultiGrid"; }
	virtual bool supports_parallel() const { return true; }

	// ---- configuration (mg_solver.h:113-205) ----
	void set_base_level(int baseLevel) { m_baseLev = baseLevel; }
	void set_surface_level(int topLevel) { m_topLev = topLevel; }
	void set_cycle_type(int type) { m_cycleType = type; }
	void set_cycle_type(const std::string& type)
	{
		if (type == "V") m_cycleType = _V_; else if (type == "W") m_cycleType = _W_; else if (type == "F") m_cycleType = _F_;
		else UG_THROW("GMG::set_cycle_type: option '" << type << "' not supported.");
	}
	void set_num_presmooth(int num) { m_numPreSmooth = num; }
	void set_num_postsmooth(int num) { m_numPostSmooth = num; }
	void set_smoother(SmartPtr<smoother_type> smoother) { m_spPreSmootherPrototype = smoother; m_spPostSmootherPrototype = smoother; }
	void set_presmoother(SmartPtr<smoother_type> smoother) { m_spPreSmootherPrototype = smoother; }
	void set_postsmoother(SmartPtr<smoother_type> smoother) { m_spPostSmootherPrototype = smoother; }
	void set_base_solver(SmartPtr<ILinearOperatorInverse<vector_type> > baseSolver) { m_spBaseSolver = baseSolver; }
	void set_transfer(SmartPtr<StdTransfer<TAlgebra> > P) { m_spTransferPrototype = P; }
	/// the reference recomputes the top-level defect after the last post-smoothing step
	/// (mg_solver_impl.hpp:1954-1958) although no caller reads it; off by default
	void set_compute_final_level_defect(bool b) { m_bFinalDefect = b; }
	void set_fuse_jacobi(bool b) { m_bFuseJacobi = b; }
	/// Galerkin coarse operators (mg_solver.h: set_rap; solver_util.lua:478 default false): the level
	/// operators below the top level are not assembled but computed as A_{l-1} = R_l A_l P_l at init
	/// (init_rap_operator, mg_solver_impl.hpp:828-1013) — only the transfers have to be handed over
	void set_rap(bool b) { m_bRAP = b; }
	// ---- the remaining calls util.solver.CreatePreconditioner makes on every GMG (solver_util.lua:846-904) ----
	/// mg_solver.h: set_gathered_base_solver_if_ambiguous — a partitioned hierarchy always gathers its base solve here
	/// (mg_solver_impl.hpp:2003-2070), so both settings give the same cycle
	void set_gathered_base_solver_if_ambiguous(bool b) { m_bGatheredBaseIfAmbiguous = b; }
	/// adaptive-grid options (rim smoothing :1830-1850, emulation of a fully refined grid): only fully refined
	/// hierarchies are handed to this class (adaptive rim handling is outside the path)
	void set_smooth_on_surface_rim(bool b) { if (b) UG_THROW("GMG: adaptive hierarchies (surface rim) are not supported by the GPU algebra"); }
	void set_emulate_full_refined_grid(bool b) { if (b) UG_THROW("GMG: adaptive hierarchies are not supported by the GPU algebra"); }
	/// inside ugcore the discretisation assembles the level operators on the CPU (assemble_level_operator :526-752) and
	/// the results arrive through set_level_operator; debug writer and statistics objects are CPU-side tooling.
	/// The stand-alone mirror accepts and ignores them so that the factory code compiles unchanged.
	template <typename TAss> void set_discretization(SmartPtr<TAss>) {}
	template <typename TWriter> void set_debug(SmartPtr<TWriter>) {}
	template <typename TStats> void set_mg_stats(SmartPtr<TStats>) {}
	/// projection of the solution to coarser levels (nonlinear / time-dependent problems): not used by the linear solve
	template <typename TTransfer> void set_projection(SmartPtr<TTransfer>) {}

	// ---- what assembly hands over (replaces assemble_level_operator :526-752 and the cached
	//      StdTransfer::prolongation()/restriction() :602-717) ----
	void set_level_operator(int lev, SmartPtr<matrix_operator_type> A) { level(lev).A = A; level(lev).AisSurface = false; }
	void set_level_transfer(int lev, SmartPtr<GPUTransferMatrix> P, SmartPtr<GPUTransferMatrix> R) { level(lev).P = P; level(lev).R = R; }
	/// surface index of every top-level index (vSurfLevelMap); empty = identity (full refinement)
	void set_surface_to_level_map(const std::vector<int>& surfIndexOfLevelIndex) { m_surfMap = surfIndexOfLevelIndex; }
	// partitioned runs
	void set_level_layouts(int lev, SmartPtr<GPUAlgebraLayouts> l) { level(lev).layouts = l; }
	/// partitioned Gauss-Seidel smoothing: the level matrix made consistent on the interface rows
	/// (what GaussSeidelBase::preprocess obtains from MakeConsistent, gauss_seidel.h:137)
	void set_level_smoother_matrix(int lev, SmartPtr<matrix_type> Acons) { level(lev).Aconsistent = Acons; }
	/// gathered base solve (mg_solver_impl.hpp:2003-2070): every rank holds the assembled global
	/// base matrix; local additive defects are summed into it with one all-reduce
	void set_gathered_base(SmartPtr<matrix_operator_type> globalA, const std::vector<int>& localToGlobal)
	{ m_spGatheredA = globalA; m_baseLocalToGlobal = localToGlobal; }

	virtual SmartPtr<ILinearIterator<vector_type> > clone() { UG_THROW("GMG::clone: not supported for the GPU algebra"); }

	virtual bool init(SmartPtr<ILinearOperator<vector_type> > J, const vector_type&) { return init(J); }
	virtual bool init(SmartPtr<ILinearOperator<vector_type> > L)
	{
		m_spSurfaceMat = sp_cast_dynamic<matrix_operator_type>(L);
		if (!m_spSurfaceMat) UG_THROW("GMG:init: Can not cast Operator to Matrix.");
		if (m_baseLev > m_topLev) UG_THROW("GMG::init: Base Level greater than Surface level.");
		if (!m_spBaseSolver) UG_THROW("GMG::init: Base Solver not set.");
		if (!m_spPreSmootherPrototype) UG_THROW("GMG::init: PreSmoother not set.");
		if (!m_spPostSmootherPrototype) UG_THROW("GMG::init: PostSmoother not set.");
		if (!m_spTransferPrototype) m_spTransferPrototype = make_sp<StdTransfer<TAlgebra> >();
		ug4b200_ctx* ctx = GPUManager::ctx();
		GPUManager::bump_generation();   // level vectors / matrices are rebuilt: captured solver graphs are stale
		UG_GPU_ZONE(GMG_Init);
		if (m_bRAP) { UG_GPU_ZONE(GMG_BuildRAP_AllLevelMat); init_rap_operator(); }
		for (int lev = m_baseLev; lev <= m_topLev; ++lev) {
			LevData& ld = level(lev);
			if (lev == m_topLev && (!ld.A || ld.AisSurface)) {
				// copy of the surface matrix (:623-656), identity map — taken anew at EVERY init: solver:init(J, u) with a
				// re-assembled J must not leave the top level smoothing with the previous operator
				ld.A = m_spSurfaceMat; ld.AisSurface = true;
			}
			if (!ld.A) UG_THROW("GMG::init: level operator of level " << lev << " missing");
			const size_t n = ld.A->num_rows();
			ld.sc.create(n); ld.sd.create(n); ld.st.create(n); ld.st2.create(n);
			for (vector_type* v : {&ld.sc, &ld.sd, &ld.st, &ld.st2}) v->set_layouts(ld.layouts);
			ld.A->device();
			if (lev > m_baseLev) {
				if (!ld.P) UG_THROW("GMG::init: prolongation of level " << lev << " missing");
				ld.transfer = m_spTransferPrototype->clone();
				ld.transfer->set_matrices(ld.P, ld.R);
				ld.transfer->init();
				ld.PreSmoother = m_spPreSmootherPrototype->clone();
				if (m_spPreSmootherPrototype == m_spPostSmootherPrototype) ld.PostSmoother = ld.PreSmoother;
				else ld.PostSmoother = m_spPostSmootherPrototype->clone();
				for (SmartPtr<smoother_type> s : {ld.PreSmoother, ld.PostSmoother}) {
					Jacobi<TAlgebra>* j = dynamic_cast<Jacobi<TAlgebra>*>(s.get());
					if (j) j->set_layouts(ld.layouts);
					GaussSeidelBase<TAlgebra>* g = dynamic_cast<GaussSeidelBase<TAlgebra>*>(s.get());
					if (g) { g->set_layouts(ld.layouts); g->set_consistent_matrix(ld.layouts ? ld.Aconsistent : SmartPtr<matrix_type>()); }
					ILU<TAlgebra>* ilu = dynamic_cast<ILU<TAlgebra>*>(s.get());
					if (ilu) { ilu->set_layouts(ld.layouts); ilu->set_consistent_matrix(ld.layouts ? ld.Aconsistent : SmartPtr<matrix_type>()); }
				}
				if (!ld.PreSmoother->init(ld.A)) UG_THROW("GMG::init: Cannot init pre-smoother for level " << lev);
				if (ld.PostSmoother != ld.PreSmoother && !ld.PostSmoother->init(ld.A))
					UG_THROW("GMG::init: Cannot init post-smoother for level " << lev);
			}
		}
		// base solver (:1171-1229)
		LevData& lb = level(m_baseLev);
		if (lb.layouts) {
			if (!m_spGatheredA) UG_THROW("GMG::init: partitioned hierarchy needs a gathered base matrix");
			m_gatheredD.create(m_spGatheredA->num_rows()); m_gatheredC.create(m_spGatheredA->num_rows());
			THROW_IF_NOT_EQUAL(m_baseLocalToGlobal.size(), lb.A->num_rows());
			GPUManager::free_bytes(m_dBaseMap);
			m_dBaseMap = (int*)GPUManager::alloc_bytes(sizeof(int) * m_baseLocalToGlobal.size());
			UG_GPU_CHECK(ug4b200_h2d(ctx, m_dBaseMap, m_baseLocalToGlobal.data(), sizeof(int) * m_baseLocalToGlobal.size()));
			if (m_gather) { ug4b200_gather_destroy(ctx, m_gather); m_gather = nullptr; }
			UG_GPU_CHECK(ug4b200_gather_create(ctx, (int64_t)m_spGatheredA->num_rows(), (int64_t)m_baseLocalToGlobal.size(),
			                                   m_baseLocalToGlobal.data(), B, &m_gather));
			UG_GPU_CHECK(ug4b200_gather_commit(ctx, m_gather));
			if (!m_spBaseSolver->init(m_spGatheredA)) UG_THROW("GMG::init: Cannot init base solver");
		} else if (!m_spBaseSolver->init(lb.A)) UG_THROW("GMG::init: Cannot init base solver on baselevel " << m_baseLev);
		// surface <-> level map
		GPUManager::free_bytes(m_dSurfMap); m_dSurfMap = nullptr;
		if (!m_surfMap.empty()) {
			THROW_IF_NOT_EQUAL(m_surfMap.size(), level(m_topLev).A->num_rows());
			m_dSurfMap = (int*)GPUManager::alloc_bytes(sizeof(int) * m_surfMap.size());
			UG_GPU_CHECK(ug4b200_h2d(ctx, m_dSurfMap, m_surfMap.data(), sizeof(int) * m_surfMap.size()));
		}
		UG_GPU_CHECK(ug4b200_sync(ctx));
		return true;
	}
	~AssembledMultiGridCycle()
	{
		GPUManager::free_bytes(m_dSurfMap); GPUManager::free_bytes(m_dBaseMap);
		if (m_gather && GPUManager::ctx_or_null()) ug4b200_gather_destroy(GPUManager::ctx_or_null(), m_gather);
	}

	/// mg_solver_impl.hpp:174-275
	virtual bool apply(vector_type& c, const vector_type& d)
	{
		UG_GPU_ZONE(GMG_Apply);
		ug4b200_ctx* ctx = GPUManager::ctx();
		LevData& top = level(m_topLev);
		THROW_IF_NOT_EQUAL(d.size(), top.sd.size());
		m_pTopC = nullptr; m_pTopD = nullptr;
		// project defect from surface to level (:211-217)
		if (m_dSurfMap) {
			UG_GPU_CHECK(ug4b200_vec_gather(ctx, (int64_t)top.sd.size(), B, top.sd.dev(), d.dev(), m_dSurfMap));
			top.sd.set_storage_type(d.get_storage_mask());
		} else {
			// identity map (full refinement): the top level accumulates its correction straight into c,
			// and the fused Jacobi pre-smoother reads the caller's defect in its first step — neither
			// the surface->level copy of d nor the level->surface copy of the correction is needed
			if (m_topLev > m_baseLev && &c != &d) {
				THROW_IF_NOT_EQUAL(c.size(), top.sc.size());
				m_pTopC = &c;
				if (m_numPreSmooth > 0 && fused_jacobi(top.PreSmoother)) m_pTopD = &d;
			}
			if (!m_pTopD) {
				UG_GPU_CHECK(ug4b200_vec_copy(ctx, top.sd.len(), top.sd.dev(), d.dev()));
				top.sd.set_storage_type(d.get_storage_mask());
			}
		}
		top.scZero = true;                     // sc = 0 (:234), carried out by the first accumulation
		if (m_topLev == m_baseLev) materialize_sc(m_topLev);
		{ UG_GPU_ZONE(GMG_Apply_lmgc); lmgc(m_topLev, m_cycleType); }   // :238
		// c = 0 (:231) ; c[surf] += sc[lev] (:244-248)
		if (m_dSurfMap) {
			c.set(0.0);
			UG_GPU_CHECK(ug4b200_vec_scatter_add(ctx, (int64_t)top.sc.size(), B, c.dev(), m_dSurfMap, top.sc.dev()));
		} else if (!m_pTopC) {
			// 0.0 + sc == sc bit for bit (up to the sign of zero)
			UG_GPU_CHECK(ug4b200_vec_copy(ctx, c.len(), c.dev(), top.sc.dev()));
		}
		m_pTopC = nullptr; m_pTopD = nullptr;
		c.set_storage_type(PST_CONSISTENT);
		const number kappa = this->damping()->damping(c, d, m_spSurfaceMat);
		if (kappa != 1.0) c *= kappa;          // :259-260
		return true;
	}
	/// mg_solver_impl.hpp:277-320
	virtual bool apply_update_defect(vector_type& c, vector_type& rD)
	{
		if (!apply(c, rD)) return false;
		m_spSurfaceMat->matmul_minus(rD, c);
		return true;
	}

  protected:
	struct LevData {
		SmartPtr<matrix_operator_type> A;
		SmartPtr<GPUTransferMatrix> P, R;
		SmartPtr<StdTransfer<TAlgebra> > transfer;
		SmartPtr<smoother_type> PreSmoother, PostSmoother;
		vector_type sc, sd, st, st2;
		SmartPtr<GPUAlgebraLayouts> layouts;
		SmartPtr<matrix_type> Aconsistent;   // partitioned Gauss-Seidel only
		bool AisSurface = false;             // A was not handed over but taken from the surface operator at init
		bool scZero = false;   // sc is logically 0: the next accumulation assigns (UG4B200_SMOOTH_SC_ZERO)
		bool stReady = false;  // st already holds S*sd (produced by the fused restriction of the finer level)
	};
	LevData& level(int lev)
	{
		if (lev < 0) UG_THROW("GMG: negative level");
		if ((int)m_vLevData.size() <= lev) { const size_t o = m_vLevData.size(); m_vLevData.resize(lev + 1); for (size_t i = o; i < m_vLevData.size(); ++i) m_vLevData[i] = make_sp<LevData>(); }
		return *m_vLevData[lev];
	}
	/// mg_solver_impl.hpp:828-1013 on a fully refined hierarchy: the top level is the surface matrix,
	/// A_{l-1} += R_l A_l P_l (AddMultiplyOf, :959) downwards; host side, results are uploaded like
	/// assembled level matrices.  Partitioned runs: the additive local matrices give additive coarse
	/// matrices (R and P of a box partition are local), which is ugcore's parallel RAP without
	/// vertical interfaces.
	void init_rap_operator()
	{
		if (!level(m_topLev).A || level(m_topLev).AisSurface) { level(m_topLev).A = m_spSurfaceMat; level(m_topLev).AisSurface = true; }
		for (int lev = m_topLev; lev > m_baseLev; --lev) {
			LevData& lf = level(lev); LevData& lc = level(lev - 1);
			if (!lf.P) UG_THROW("GMG::init_rap_operator: prolongation of level " << lev << " missing");
			if (!lf.R) { lf.R = make_sp<GPUTransferMatrix>(); lf.R->set_as_transpose_of(*lf.P); }   // std_transfer_impl.h:694-695
			SmartPtr<matrix_operator_type> Ac = make_sp<matrix_operator_type>();
			Ac->resize_and_clear(lf.P->num_cols(), lf.P->num_cols());
			AddMultiplyOf(static_cast<matrix_type&>(*Ac), *lf.R, static_cast<const matrix_type&>(*lf.A), *lf.P);
			lc.A = Ac;
		}
	}

	/// correction of a level; on the top level of a fully refined hierarchy this is the caller's c
	vector_type& SC(int lev) { return (lev == m_topLev && m_pTopC) ? *m_pTopC : level(lev).sc; }
	void materialize_sc(int lev)
	{
		LevData& ld = level(lev);
		if (ld.scZero) { SC(lev).set(0.0); ld.scZero = false; }
	}
	Jacobi<TAlgebra>* fused_jacobi(SmartPtr<smoother_type> s)
	{
		if (!m_bFuseJacobi || !s) return nullptr;
		Jacobi<TAlgebra>* j = dynamic_cast<Jacobi<TAlgebra>*>(s.get());
		return (j && j->damping()->constant_damping()) ? j : nullptr;
	}
	bool base_solver_overwrites() const
	{
		const ILinearOperatorInverse<vector_type>* bs = m_spBaseSolver.get();
		return dynamic_cast<const LU<TAlgebra>*>(bs) != nullptr || dynamic_cast<const CoarseCG<TAlgebra>*>(bs) != nullptr ||
		       dynamic_cast<const CycleAsBaseSolver<TAlgebra>*>(bs) != nullptr;
	}
	void make_consistent(vector_type& v)
	{
		if (!v.layouts()) return;
		v.set_storage_type(PST_ADDITIVE);
		if (!v.change_storage_type(PST_CONSISTENT)) UG_THROW("GMG: cannot make correction consistent");
	}

	/// mg_solver_impl.hpp:2089-2136
	void lmgc(int lev, int cycleType)
	{
		if (lev == m_baseLev) { base_solve(m_topLev); return; }
		else if (lev < m_baseLev) UG_THROW("GMG::lmgc: call lmgc only for lev > baseLev.");
		presmooth_and_restriction(lev);
		if (lev - 1 == m_baseLev) base_solve(lev - 1);
		else if (cycleType == _F_) { lmgc(lev - 1, _F_); lmgc(lev - 1, _V_); }
		else for (int i = 0; i < cycleType; ++i) lmgc(lev - 1, cycleType);
		prolongation_and_postsmooth(lev);
	}

	/// mg_solver_impl.hpp:1685-1816
	void presmooth_and_restriction(int lev)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		LevData& lf = level(lev); LevData& lc = level(lev - 1);
		Jacobi<TAlgebra>* jac = fused_jacobi(lf.PreSmoother);
		{
		UG_GPU_ZONE(GMG_PreSmooth);                                   // mg_solver_impl.hpp:1698
		if (jac && m_numPreSmooth > 0) {
			// fused: st0 = D sd ; then per step one kernel  { sc += st ; sd -= A st ; [st' = D sd] }
			const int64_t n = (int64_t)lf.sd.size();
			vector_type* cur = &lf.st; vector_type* alt = &lf.st2;
			const vector_type* dsrc = (lev == m_topLev) ? m_pTopD : nullptr; // defect still in the caller's vector
			if (!lf.stReady) {
				UG_GPU_CHECK(ug4b200_jacobi_step(ctx, n, B, jac->diag_inv_dev(), cur->dev(), dsrc ? dsrc->dev() : lf.sd.dev()));
				make_consistent(*cur);
			}
			lf.stReady = false;
			for (int nu = 0; nu < m_numPreSmooth; ++nu) {
				const bool last = (nu == m_numPreSmooth - 1);
				const int flags = UG4B200_SMOOTH_ADD_IN | (last ? 0 : UG4B200_SMOOTH_JACOBI) | (lf.scZero ? UG4B200_SMOOTH_SC_ZERO : 0);
				if (!last) alt->arm_consistent_push();   // the kernel stores the interface rows of its output at the neighbours
				UG_GPU_CHECK(ug4b200_jacobi_smooth_fused_src(ctx, lf.A->device(), jac->diag_inv_dev(), lf.sd.dev(),
				                                             dsrc ? dsrc->dev() : nullptr, cur->dev(), last ? nullptr : alt->dev(),
				                                             SC(lev).dev(), flags));
				if (dsrc) { lf.sd.set_storage_type(dsrc->get_storage_mask()); dsrc = nullptr; }
				lf.scZero = false;
				if (!last) { make_consistent(*alt); std::swap(cur, alt); }
			}
		} else {
			if (m_numPreSmooth > 0) materialize_sc(lev);
			for (int nu = 0; nu < m_numPreSmooth; ++nu) {
				if (!lf.PreSmoother->apply(lf.st, lf.sd)) UG_THROW("GMG: Smoothing step " << nu + 1 << " on level " << lev << " failed.");
				lf.A->apply_sub(lf.sd, lf.st);                      // :1726
				if (nu < m_numPreSmooth - 1) SC(lev) += lf.st;      // :1729-1730
			}
			if (m_numPreSmooth > 0) SC(lev) += lf.st;               // :1783-1784
		}
		}
		UG_GPU_ZONE(GMG_Restrict_Transfer);                           // mg_solver_impl.hpp:1800
		// sc_{l-1} = 0 (:1780): deferred to the first accumulation; the base solvers overwrite it
		if (lev - 1 == m_baseLev && base_solver_overwrites()) lc.scZero = false;
		else lc.scZero = true;
		// restriction (:1802); with a fused Jacobi pre-smoother on the coarse level its first step
		// st = D sd is computed by the same kernel
		Jacobi<TAlgebra>* jc = (lev - 1 > m_baseLev && m_numPreSmooth > 0 && B == 1) ? fused_jacobi(lc.PreSmoother) : nullptr;
		if (jc) {
			lc.st.arm_consistent_push();
			UG_GPU_CHECK(ug4b200_restrict_jacobi_fused(ctx, lf.transfer->restriction()->device(), jc->diag_inv_dev(), lc.sd.dev(),
			                                           lf.transfer->restriction_damping(), lf.sd.dev(), lc.st.dev()));
			lc.sd.set_storage_type(lf.sd.get_storage_mask());
			make_consistent(lc.st);
			lc.stReady = true;
		} else {
			lf.transfer->do_restrict(lc.sd, lf.sd);
			lc.stReady = false;
		}
	}

	/// mg_solver_impl.hpp:1818-1964
	void prolongation_and_postsmooth(int lev)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		LevData& lf = level(lev); LevData& lc = level(lev - 1);
		lc.stReady = false;
		materialize_sc(lev - 1);
		{ UG_GPU_ZONE(GMG_Prolongate_Transfer); lf.transfer->prolongate(lf.st, SC(lev - 1)); }   // :1863-1865
		UG_GPU_ZONE(GMG_PostSmooth);                                  // mg_solver_impl.hpp:1913 (incl. GMG_AddCoarseGridCorrection :1904)
		Jacobi<TAlgebra>* jac = fused_jacobi(lf.PostSmoother);
		if (jac) {
			vector_type* cur = &lf.st; vector_type* alt = &lf.st2;
			// without interfaces the last step's output needs no exchange and is accumulated in the kernel
			const bool addOut = !lf.layouts && m_numPostSmooth > 0;
			for (int nu = 0; nu < m_numPostSmooth; ++nu) {
				const bool last = (nu == m_numPostSmooth - 1);
				const int flags = UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_JACOBI | (lf.scZero ? UG4B200_SMOOTH_SC_ZERO : 0) |
				                  ((last && addOut) ? UG4B200_SMOOTH_ADD_OUT : 0);
				alt->arm_consistent_push();
				UG_GPU_CHECK(ug4b200_jacobi_smooth_fused(ctx, lf.A->device(), jac->diag_inv_dev(), lf.sd.dev(), cur->dev(), alt->dev(),
				                                         SC(lev).dev(), flags));
				lf.scZero = false;
				make_consistent(*alt);
				std::swap(cur, alt);
			}
			if (!addOut) { materialize_sc(lev); SC(lev) += *cur; }   // last :1905 / :1943
			if (m_bFinalDefect && lev >= m_topLev) lf.A->apply_sub(lf.sd, *cur);
		} else {
			materialize_sc(lev);
			SC(lev) += lf.st;                                       // :1905
			for (int nu = 0; nu < m_numPostSmooth; ++nu) {
				lf.A->apply_sub(lf.sd, lf.st);                      // :1919
				if (!lf.PostSmoother->apply(lf.st, lf.sd)) UG_THROW("GMG: Smoothing step " << nu + 1 << " on level " << lev << " failed.");
				SC(lev) += lf.st;                                   // :1943
			}
			if (m_bFinalDefect && lev >= m_topLev) lf.A->apply_sub(lf.sd, lf.st); // :1954-1958
		}
	}

	/// mg_solver_impl.hpp:1967-2086
	void base_solve(int lev)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		LevData& ld = level(lev);
		if (!ld.layouts) {
			UG_GPU_ZONE(GMG_BaseSolver_Apply);                        // mg_solver_impl.hpp:1990
			if (!m_spBaseSolver->apply(ld.sc, ld.sd)) UG_THROW("GMG::lmgc: Base solver on base level " << lev << " failed.");
		} else {
			// gathered: additive local defects -> global consistent defect on every rank
			UG_GPU_ZONE(GMG_GatheredBaseSolver_Apply);                // mg_solver_impl.hpp:2018-2045
			UG_GPU_CHECK(ug4b200_gather_sum(ctx, m_gather, m_gatheredD.dev(), ld.sd.dev()));
			if (!m_spBaseSolver->apply(m_gatheredC, m_gatheredD)) UG_THROW("GMG::lmgc: Base solver on base level " << lev << " failed.");
			UG_GPU_CHECK(ug4b200_vec_gather(ctx, (int64_t)ld.sc.size(), B, ld.sc.dev(), m_gatheredC.dev(), m_dBaseMap));
			ld.sc.set_storage_type(PST_CONSISTENT);
		}
		ld.scZero = false;
		if (lev >= m_topLev) ld.A->apply_sub(ld.sd, ld.sc);         // :2075-2078
	}

	int m_baseLev, m_topLev, m_cycleType, m_numPreSmooth, m_numPostSmooth;
	bool m_bFinalDefect, m_bFuseJacobi;
	bool m_bRAP = false, m_bGatheredBaseIfAmbiguous = false;
	SmartPtr<smoother_type> m_spPreSmootherPrototype, m_spPostSmootherPrototype;
	SmartPtr<ILinearOperatorInverse<vector_type> > m_spBaseSolver;
	SmartPtr<StdTransfer<TAlgebra> > m_spTransferPrototype;
	SmartPtr<matrix_operator_type> m_spSurfaceMat;
	std::vector<SmartPtr<LevData> > m_vLevData;
	std::vector<int> m_surfMap;
	int* m_dSurfMap = nullptr;
	vector_type* m_pTopC = nullptr;        // set during apply(): the caller's c doubles as the top level's sc
	const vector_type* m_pTopD = nullptr;  // set during apply(): the caller's d is read by the first smoothing step
	// gathered base
	SmartPtr<matrix_operator_type> m_spGatheredA;
	std::vector<int> m_baseLocalToGlobal;
	int* m_dBaseMap = nullptr;
	ug4b200_gather* m_gather = nullptr;
	vector_type m_gatheredD, m_gatheredC;
};

template <typename TAlgebra>
bool CycleAsBaseSolver<TAlgebra>::init(SmartPtr<ILinearOperator<vector_type> > L) { m_spOp = L; return m_spCycle->init(L); }
template <typename TAlgebra>
bool CycleAsBaseSolver<TAlgebra>::apply(vector_type& u, const vector_type& f) { return m_spCycle->apply(u, f); }
template <typename TAlgebra>
bool CycleAsBaseSolver<TAlgebra>::apply_return_defect(vector_type& u, vector_type& f)
{
	if (!m_spCycle->apply(u, f)) return false;
	m_spOp->apply_sub(f, u);
	return true;
}

} // namespace ug
