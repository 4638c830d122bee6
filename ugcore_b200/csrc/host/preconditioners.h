// preconditioners.h — Jacobi and (multicolour) Gauss-Seidel for the GPU algebra.
//
//   Jacobi            ugbase/lib_algebra/operator/preconditioner/jacobi.h:155-300
//                     (legacy shape: GPUJacobi, gpujacobi.h:58-196)
//   GaussSeidel / BackwardGaussSeidel / SymmetricGaussSeidel
//                     ugbase/lib_algebra/operator/preconditioner/gauss_seidel.h:50-380
//                     kernels: ugbase/lib_algebra/algebra_common/core_smoothers.h:105-206
// Multicolour GS = the reference's lexicographic sweep applied in a colour-sorted DoF
// order; the reordering follows the pattern ILU uses for its ordering algorithms
// (ilu.h:434-461 SetMatrixAsPermutation, :593-610 SetVectorAsPermutation).
//   ILU               ugbase/lib_algebra/operator/preconditioner/ilu.h:324-700 (factorisation on the host at
//                     init: ilu_factor.h; triangular solves = the Gauss-Seidel sweep kernels over the two factors)
#pragma once
#include "operators.h"
#include "ilu_factor.h"

namespace ug {

template <typename TAlgebra>
class Jacobi : public IPreconditioner<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef IPreconditioner<TAlgebra> base_type;
	typedef typename base_type::matrix_operator_type matrix_operator_type;
	using base_type::damping;
	enum { B = TAlgebra::blockSize };

	Jacobi() : m_bBlock(true) { this->set_damp(1.0); }
	explicit Jacobi(number damp) : m_bBlock(true) { this->set_damp(damp); }
	~Jacobi() { free_diag(); }
	virtual SmartPtr<ILinearIterator<vector_type> > clone()
	{
		SmartPtr<Jacobi<TAlgebra> > j(new Jacobi<TAlgebra>());
		j->set_damp(this->damping());
		j->m_bBlock = m_bBlock;
		return j;
	}
	virtual bool supports_parallel() const { return true; }
	void set_block(bool b) { m_bBlock = b; }
	virtual const char* name() const { return "Jacobi"; }

	/// device array of inverted (damped) diagonal blocks, column-major, for fused V-cycle kernels
	const double* diag_inv_dev() const { return m_diagInv; }

	virtual bool apply(vector_type& c, const vector_type& d)
	{
		if (!this->m_bInit) return false;
		if (d.layouts() && !d.has_storage_type(PST_ADDITIVE))
			UG_THROW(name() << "::apply: Wrong parallel storage format. Defect must be additive.");
		THROW_IF_NOT_EQUAL(c.size(), d.size());
		if (!step(this->m_spApproxOperator, c, d)) return false;
		// constant damping is folded into the inverse diagonal (jacobi.h:283-290)
		if (!damping()->constant_damping()) {
			const number kappa = damping()->damping(c, d, this->m_spApproxOperator);
			if (kappa != 1.0) c *= kappa;
		}
		if (c.layouts() && !c.change_storage_type(PST_CONSISTENT))
			UG_THROW(name() << "::apply': Cannot change parallel storage type of correction to consistent.");
		return true;
	}

  protected:
	virtual bool preprocess(SmartPtr<matrix_operator_type> pOp)
	{
		matrix_type& mat = *pOp;
		if (mat.num_rows() != mat.num_cols()) return false;
		ug4b200_ctx* ctx = GPUManager::ctx();
		GPUManager::bump_generation();   // the device buffers a captured solver graph points at are rebuilt
		free_diag();
		m_n = mat.num_rows();
		m_diagInv = GPUManager::alloc(m_n * B * B);
		number damp = 1.0;
		if (damping()->constant_damping()) damp = damping()->damping();
		if (!m_layouts) {
			UG_GPU_CHECK(ug4b200_jacobi_prepare(ctx, mat.device(), damp, m_bBlock ? 1 : 0, m_diagInv));
		} else {
			// parallel: the additive diagonal is made consistent first (jacobi.h:171-187)
			double* diag = GPUManager::alloc(m_n * B * B);
			UG_GPU_CHECK(ug4b200_matrix_get_diag(ctx, mat.device(), diag));
			UG_GPU_CHECK(ug4b200_additive_to_consistent(ctx, m_layouts->iface(), diag, B * B));
			UG_GPU_CHECK(ug4b200_jacobi_invert_diag(ctx, (int64_t)m_n, B, damp, m_bBlock ? 1 : 0, diag, m_diagInv));
			UG_GPU_CHECK(ug4b200_sync(ctx));
			GPUManager::release(diag, m_n * B * B);
		}
		return true;
	}
	virtual bool step(SmartPtr<matrix_operator_type>, vector_type& c, const vector_type& d)
	{
		UG_GPU_CHECK(ug4b200_jacobi_step(GPUManager::ctx(), (int64_t)m_n, B, m_diagInv, c.dev(), d.dev()));
		if (c.layouts()) {
			c.set_storage_type(PST_ADDITIVE);
			if (!c.change_storage_type(PST_CONSISTENT)) return false;
		}
		return true;
	}
	virtual bool postprocess() { return true; }

  public:
	/// partitioned runs: layouts of the level this smoother lives on (mat.layouts() in ugcore)
	void set_layouts(SmartPtr<GPUAlgebraLayouts> l) { m_layouts = l; }

  protected:
	void free_diag() { if (m_diagInv) GPUManager::release(m_diagInv, m_n * B * B); m_diagInv = nullptr; }
	bool m_bBlock;
	double* m_diagInv = nullptr;
	size_t m_n = 0;
	SmartPtr<GPUAlgebraLayouts> m_layouts;
};

/// Common base of the three sweeps (GaussSeidelBase, gauss_seidel.h:50-255)
template <typename TAlgebra>
class GaussSeidelBase : public IPreconditioner<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef IPreconditioner<TAlgebra> base_type;
	typedef typename base_type::matrix_operator_type matrix_operator_type;
	enum { B = TAlgebra::blockSize };

	GaussSeidelBase() : m_relax(1.0) {}
	void set_sor_relax(number relaxFactor) { m_relax = relaxFactor; }
	virtual bool supports_parallel() const { return true; }
	/// partitioned runs: layouts of the level this smoother lives on (mat.layouts() in ugcore) and the
	/// level matrix made consistent on the interface rows (MakeConsistent(*pOp, m_A), gauss_seidel.h:137,
	/// parallel_matrix_overlap_impl.h:438-459).  ugcore builds m_A inside preprocess by exchanging matrix
	/// rows over MPI; here the caller does that exchange on the host at init (ugcore_b200/dist.py:
	/// make_consistent) and hands the result over — the device is not involved.
	void set_layouts(SmartPtr<GPUAlgebraLayouts> l) { m_layouts = l; }
	void set_consistent_matrix(SmartPtr<matrix_type> A) { m_spConsistent = A; }
	/// gauss_seidel.h:96-100: the two other parallel modes of ugcore's Gauss-Seidel (consistent interfaces, overlap);
	/// the default mode (unique defect, Dirichlet rows on the h-slaves) is the one implemented
	void enable_consistent_interfaces(bool enable) { if (enable) UG_THROW(this->name() << ": the consistent-interfaces mode is not available for the GPU algebra"); }
	void enable_overlap(bool enable) { if (enable) UG_THROW(this->name() << ": the overlap mode is not available for the GPU algebra"); }
	/// ordering: old index i -> new index perm[i]; the new order must be colour-sorted with the
	/// given colour offsets.  Without it a greedy colouring of the stored pattern is used.
	void set_coloring(const std::vector<int>& perm, const std::vector<int64_t>& colorPtr) { m_perm = perm; m_colorPtr = colorPtr; m_userPerm = true; }
	const std::vector<int>& ordering() const { return m_perm; }
	const std::vector<int64_t>& color_ptr() const { return m_colorPtr; }
	~GaussSeidelBase() { free_dev(); }

  protected:
	virtual int kind() const = 0;
	void copy_settings(GaussSeidelBase& o) const { o.m_relax = m_relax; o.m_userPerm = m_userPerm; if (m_userPerm) { o.m_perm = m_perm; o.m_colorPtr = m_colorPtr; } o.set_damp(const_cast<GaussSeidelBase*>(this)->damping()); }

	virtual bool preprocess(SmartPtr<matrix_operator_type> pOp)
	{
		matrix_type& A = m_layouts ? parallel_matrix(*pOp) : static_cast<matrix_type&>(*pOp);
		THROW_IF_NOT_EQUAL(A.num_rows(), A.num_cols());
		const size_t n = A.num_rows();
		const std::vector<int64_t>& rp = A.crs_rowptr();
		const std::vector<int>& ci = A.crs_cols();
		const int BB = B * B;
		// parallel (gauss_seidel.h:134-142): consistent matrix with the rows of the h-slaves set to
		// Dirichlet rows (SetDirichletRow, sparsematrix_util.h:878-897: all blocks 0, diagonal block = 1)
		std::vector<double> vaPar;
		if (m_layouts) {
			vaPar = A.crs_vals();
			const std::vector<int> slaves = m_layouts->slave_indices();
			for (size_t k = 0; k < slaves.size(); ++k) {
				const int r = slaves[k];
				bool haveDiag = false;
				for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
					for (int t = 0; t < BB; ++t) vaPar[p * BB + t] = 0.0;
					if (ci[p] == r) { haveDiag = true; for (int t = 0; t < B; ++t) vaPar[p * BB + t + B * t] = 1.0; }
				}
				if (!haveDiag) UG_THROW(this->name() << ": interface row " << r << " has no diagonal connection");
			}
		}
		const std::vector<double>& va = m_layouts ? vaPar : A.crs_vals();
		// the sweeps divide by a_ii (core_smoothers.h:105-206): a row without a stored diagonal connection has no
		// meaningful sweep (the reference divides by the 0 its const access returns) — refuse it here
		for (size_t r = 0; r < n; ++r) {
			bool haveDiag = false;
			for (int64_t p = rp[r]; p < rp[r + 1] && !haveDiag; ++p) haveDiag = ((size_t)ci[p] == r);
			if (!haveDiag) UG_THROW(this->name() << ": row " << r << " has no diagonal connection");
		}
		if (m_userPerm) { THROW_IF_NOT_EQUAL(m_perm.size(), n); }
		else {
			// greedy colouring in row order, colours sorted, stable inside a colour; recomputed on every
			// preprocess (the pattern may have changed at the same n)
			std::vector<int> color(n); int nc = 0;
			ug4b200_color_greedy((int64_t)n, rp.data(), ci.data(), color.data(), &nc);
			m_colorPtr.assign(nc + 1, 0);
			for (size_t i = 0; i < n; ++i) m_colorPtr[color[i] + 1]++;
			for (int k = 0; k < nc; ++k) m_colorPtr[k + 1] += m_colorPtr[k];
			std::vector<int64_t> fill(m_colorPtr.begin(), m_colorPtr.end() - 1);
			m_perm.resize(n);
			for (size_t i = 0; i < n; ++i) m_perm[i] = (int)fill[color[i]]++;
		}
		// PA(perm[r], perm[c]) = A(r, c)   (SetMatrixAsPermutation, permutation_util.h:50-64)
		std::vector<int> inv(n);
		for (size_t i = 0; i < n; ++i) inv[m_perm[i]] = (int)i;
		std::vector<int64_t> prp(n + 1, 0);
		for (size_t nr = 0; nr < n; ++nr) prp[nr + 1] = prp[nr] + (rp[inv[nr] + 1] - rp[inv[nr]]);
		std::vector<int> pci(ci.size()); std::vector<double> pva(va.size());
		std::vector<std::pair<int, int64_t> > row;
		for (size_t nr = 0; nr < n; ++nr) {
			const int r = inv[nr];
			row.clear();
			for (int64_t p = rp[r]; p < rp[r + 1]; ++p) row.push_back(std::make_pair(m_perm[ci[p]], p));
			std::sort(row.begin(), row.end());
			int64_t q = prp[nr];
			for (size_t k = 0; k < row.size(); ++k, ++q) {
				pci[q] = row[k].first;
				for (int t = 0; t < BB; ++t) pva[q * BB + t] = va[row[k].second * BB + t];
			}
		}
		if (ug4b200_color_check((int64_t)n, prp.data(), pci.data(), (int)m_colorPtr.size() - 1, m_colorPtr.data()) != 0)
			UG_THROW(this->name() << ": the given ordering is not a valid multicolour ordering of the matrix pattern");
		free_dev();
		GPUManager::bump_generation();   // the device buffers a captured solver graph points at are rebuilt
		ug4b200_ctx* ctx = GPUManager::ctx();
		UG_GPU_CHECK(ug4b200_matrix_upload_crs(ctx, B, (int64_t)n, (int64_t)n, prp.data(), pci.data(), pva.data(), 0, &m_PA));
		ug4b200_matrix_info info; ug4b200_matrix_get_info(m_PA, &info);
		m_n = n;
		m_dPerm = (int*)GPUManager::alloc_bytes(sizeof(int) * (n ? n : 1));
		UG_GPU_CHECK(ug4b200_h2d(ctx, m_dPerm, m_perm.data(), sizeof(int) * n));
		UG_GPU_CHECK(ug4b200_sync(ctx));
		m_pd = GPUManager::alloc(n * B); m_pc = GPUManager::alloc(n * B);
		if (m_layouts) { m_dUnique.create(n); m_dUnique.set_layouts(m_layouts); }
		return true;
	}
	virtual bool step(SmartPtr<matrix_operator_type>, vector_type& c, const vector_type& d)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		const double* dsrc = d.dev();
		if (m_layouts) {
			// make defect unique (gauss_seidel.h:204-207): the sum of all copies on the h-master, 0 on the
			// slaves.  The reference clones d in every step ("todo: do not clone every time"); the copy
			// lives in a member here so that a captured CUDA graph always sees the same buffer.
			m_dUnique = d;
			if (!m_dUnique.change_storage_type(PST_UNIQUE)) UG_THROW(this->name() << ": cannot make the defect unique");
			dsrc = m_dUnique.dev();
		}
		// Pd[perm[i]] = d[i]; sweep; c[i] = Pc[perm[i]]   (ilu.h:605-610)
		UG_GPU_CHECK(ug4b200_vec_scatter(ctx, (int64_t)m_n, B, m_pd, m_dPerm, dsrc));
		UG_GPU_CHECK(ug4b200_gs_step(ctx, m_PA, (int)m_colorPtr.size() - 1, m_colorPtr.data(), kind(), m_relax, m_pc, m_pd));
		UG_GPU_CHECK(ug4b200_vec_gather(ctx, (int64_t)m_n, B, c.dev(), m_pc, m_dPerm));
		if (m_layouts) {
			// the slave rows are Dirichlet rows and their defect is 0: the correction is unique; make it
			// consistent (gauss_seidel.h:211-215)
			c.set_storage_type(PST_UNIQUE);
			if (!c.change_storage_type(PST_CONSISTENT)) return false;
		} else c.set_storage_type(PST_CONSISTENT);
		return true;
	}
	matrix_type& parallel_matrix(matrix_type& A)
	{
		if (!m_spConsistent)
			UG_THROW(this->name() << ": a partitioned level needs the consistent level matrix (set_consistent_matrix; "
			         "ug4b200_solver_set_smoother_matrix)");
		THROW_IF_NOT_EQUAL(m_spConsistent->num_rows(), A.num_rows());
		THROW_IF_NOT_EQUAL((size_t)m_layouts->num_local(), A.num_rows());
		return *m_spConsistent;
	}
	virtual bool postprocess() { return true; }
	void free_dev()
	{
		if (m_PA && GPUManager::ctx_or_null()) ug4b200_matrix_destroy(GPUManager::ctx_or_null(), m_PA);
		m_PA = nullptr;
		GPUManager::free_bytes(m_dPerm); m_dPerm = nullptr;
		if (m_pd) GPUManager::release(m_pd, m_n * B);
		if (m_pc) GPUManager::release(m_pc, m_n * B);
		m_pd = m_pc = nullptr;
	}

	number m_relax;
	std::vector<int> m_perm;
	std::vector<int64_t> m_colorPtr;
	bool m_userPerm = false;
	ug4b200_matrix* m_PA = nullptr;
	int* m_dPerm = nullptr;
	double *m_pd = nullptr, *m_pc = nullptr;
	size_t m_n = 0;
	SmartPtr<GPUAlgebraLayouts> m_layouts;
	SmartPtr<matrix_type> m_spConsistent;
	vector_type m_dUnique;
};

/// ILU(0) / ILU(beta) (ilu.h:324-700).  preprocess keeps ugcore's sequence — copy of the matrix, parallel:
/// slave rows added to the master rows + Dirichlet rows on the slaves (:536-543), ordering (:434-461),
/// FactorizeILUSorted / FactorizeILUBeta (:573-575) — on the host; the factors are then split into
/// L (strictly lower part + unit diagonal) and U (diagonal + upper part) and uploaded.  invert_L (:233-252) is
/// a forward Gauss-Seidel sweep over L with relaxation 1 (c_i = 1.0 * (d_i - sum_j<i l_ij c_j) / 1.0), invert_U
/// (:257-322) a backward sweep over U (c_i = 1.0 * (d_i - sum_j>i u_ij c_j) / u_ii): the same expressions as
/// the reference's, evaluated by the sweep kernels group by group, where a group is a set of rows without
/// mutual dependencies:
///   * colour-sorted ordering (set_coloring / set_multicolor_ordering, the ordering "algorithm" made for the
///     device): groups = colours, rows keep their order -> bit-identical to the reference's ILU in that ordering;
///   * any other ordering (natural, Cuthill-McKee, user): groups = level sets of the factor's dependency graph
///     (level scheduling); rows are renumbered by level, so a row accumulates its terms in another order than
///     the reference does (round-off level differences), and a sweep costs one launch per level.
/// invert_U's "near-zero last diagonal entry" guard (ilu.h:285-297, set_inversion_eps) is not evaluated on the
/// device.  Scalar and 2x2 / 3x3 block algebras (block entries: DenseMatrix arithmetic restated in ilu_factor.h).
template <typename TAlgebra>
class ILU : public IPreconditioner<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	typedef typename TAlgebra::matrix_type matrix_type;
	typedef IPreconditioner<TAlgebra> base_type;
	typedef typename base_type::matrix_operator_type matrix_operator_type;
	typedef std::vector<size_t> ordering_container_type;
	enum { B = TAlgebra::blockSize };

	explicit ILU(double beta = 0.0) : m_beta(beta) {}
	~ILU() { free_dev(); }
	virtual const char* name() const { return "ILU"; }
	virtual bool supports_parallel() const { return true; }
	virtual SmartPtr<ILinearIterator<vector_type> > clone()
	{
		SmartPtr<ILU> c(new ILU(m_beta));
		c->m_sortEps = m_sortEps; c->m_invEps = m_invEps; c->m_bDisablePreprocessing = m_bDisablePreprocessing;
		c->m_mode = m_mode; c->m_userOrdering = m_userOrdering; c->m_userColorPtr = m_userColorPtr;
		c->set_damp(this->damping());
		return c;
	}
	void set_beta(double beta) { m_beta = beta; }
	void set_sort_eps(number eps) { m_sortEps = eps; }
	void set_inversion_eps(number eps) { m_invEps = eps; }
	void set_disable_preprocessing(bool b) { m_bDisablePreprocessing = b; }
	/// ILU::set_sort(true) (ilu.h:403-414): NativeCuthillMcKeeOrdering, i.e. ComputeCuthillMcKeeOrder(o, neighbours,
	/// reverse = false, preserveConsec = true) (native_cuthill_mckee.h:101, 124)
	void set_sort(bool b) { m_mode = b ? ORDER_CMK : ORDER_NONE; }
	/// what set_ordering_algorithm(..) ends in (ilu.h:434-461): ordering[old] = new
	void set_ordering(const ordering_container_type& o) { m_userOrdering = o; m_userColorPtr.clear(); m_mode = ORDER_USER; }
	/// colour-sorted ordering with its colour offsets (see class comment)
	void set_coloring(const std::vector<int>& perm, const std::vector<int64_t>& colorPtr)
	{ m_userOrdering.assign(perm.begin(), perm.end()); m_userColorPtr = colorPtr; m_mode = ORDER_USER; }
	/// greedy multicolour ordering of the stored pattern (the one the Gauss-Seidel smoothers use)
	void set_multicolor_ordering(bool b) { m_mode = b ? ORDER_COLOR : ORDER_NONE; }
	/// partitioned runs: see GaussSeidelBase
	void set_layouts(SmartPtr<GPUAlgebraLayouts> l) { m_layouts = l; }
	void set_consistent_matrix(SmartPtr<matrix_type> A) { m_spConsistent = A; }
	/// ilu.h:423-429: ugcore's two other parallel modes; the default one (ilu.h:536-543) is implemented
	void enable_consistent_interfaces(bool enable) { if (enable) UG_THROW("ILU: the consistent-interfaces mode is not available for the GPU algebra"); }
	void enable_overlap(bool enable) { if (enable) UG_THROW("ILU: the overlap mode is not available for the GPU algebra"); }
	/// ilu.h:397-400: any object with ugcore's IOrderingAlgorithm interface (init(&A), compute(), ordering())
	/// is evaluated on the host copy at preprocess time inside ugcore; here its result is what counts
	template <typename TOrderingAlgo> void set_ordering_algorithm(SmartPtr<TOrderingAlgo> algo)
	{ if (algo) set_ordering(algo->ordering()); else { m_mode = ORDER_NONE; m_userOrdering.clear(); m_userColorPtr.clear(); } }
	int num_groups_L() const { return (int)m_ptrL.size() - 1; }
	int num_groups_U() const { return (int)m_ptrU.size() - 1; }

  protected:
	enum { ORDER_NONE = 0, ORDER_CMK, ORDER_COLOR, ORDER_USER };

	virtual bool preprocess(SmartPtr<matrix_operator_type> pOp)
	{
		if (m_bDisablePreprocessing && m_L) return true;
		matrix_type* pA = &static_cast<matrix_type&>(*pOp);
		if (m_layouts) {
			if (!m_spConsistent) UG_THROW("ILU: a partitioned matrix needs its consistent counterpart (set_consistent_matrix)");
			pA = m_spConsistent.get();
		}
		matrix_type& A = *pA;
		THROW_IF_NOT_EQUAL(A.num_rows(), A.num_cols());
		const int64_t n = (int64_t)A.num_rows();
		const std::vector<int64_t>& rp = A.crs_rowptr();
		const std::vector<int>& ci = A.crs_cols();
		std::vector<double> va = A.crs_vals();                 // m_ILU = mat (ilu.h:533)
		const int BB = B * B;
		if (m_layouts) {
			// rows of the h-slaves become Dirichlet rows (ilu.h:539-543)
			const std::vector<int> slaves = m_layouts->slave_indices();
			for (size_t k = 0; k < slaves.size(); ++k) {
				const int r = slaves[k];
				bool haveDiag = false;
				for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
					for (int t = 0; t < BB; ++t) va[p * BB + t] = 0.0;
					if (ci[p] == r) { for (int t = 0; t < B; ++t) va[p * BB + t + B * t] = 1.0; haveDiag = true; }
				}
				if (!haveDiag) UG_THROW("ILU: interface row " << r << " has no diagonal connection");
			}
		}
		// ---- ordering (apply_ordering, ilu.h:434-461) ----
		std::vector<size_t> ord((size_t)n);
		std::vector<int64_t> colorPtr;
		for (int64_t i = 0; i < n; ++i) ord[(size_t)i] = (size_t)i;
		if (m_mode == ORDER_CMK) GetCuthillMcKeeOrder(n, rp.data(), ci.data(), ord, false, true);
		else if (m_mode == ORDER_COLOR) {
			std::vector<int> color((size_t)(n > 0 ? n : 1)); int nc = 0;
			ug4b200_color_greedy(n, rp.data(), ci.data(), color.data(), &nc);
			colorPtr.assign((size_t)nc + 1, 0);
			for (int64_t i = 0; i < n; ++i) colorPtr[(size_t)color[(size_t)i] + 1]++;
			for (int k = 0; k < nc; ++k) colorPtr[(size_t)k + 1] += colorPtr[(size_t)k];
			std::vector<int64_t> fill(colorPtr.begin(), colorPtr.end() - 1);
			for (int64_t i = 0; i < n; ++i) ord[(size_t)i] = (size_t)fill[(size_t)color[(size_t)i]]++;
		} else if (m_mode == ORDER_USER) {
			THROW_IF_NOT_EQUAL(m_userOrdering.size(), (size_t)n);
			ord = m_userOrdering; colorPtr = m_userColorPtr;
		}
		{
			std::vector<char> seen((size_t)n, 0);
			for (int64_t i = 0; i < n; ++i) {
				if (ord[(size_t)i] >= (size_t)n || seen[ord[(size_t)i]]) UG_THROW("ILU: the ordering is not a permutation");
				seen[ord[(size_t)i]] = 1;
			}
		}
		// PA(ord[r], ord[c]) = A(r, c)  (SetMatrixAsPermutation, permutation_util.h:50-64)
		std::vector<int64_t> prp; std::vector<int> pci; std::vector<double> pva;
		permute(n, rp, ci, va, ord, ord, prp, pci, pva);
		// ---- factorisation (ilu.h:573-575; SparseMatrix::rows_sorted == true) ----
		FactorizeILU(B, n, prp, pci, pva, m_beta, m_sortEps);
		// ---- split into L (unit diagonal) and U ----
		std::vector<int64_t> lrp((size_t)n + 1, 0), urp((size_t)n + 1, 0);
		std::vector<int> lci, uci; std::vector<double> lva, uva;
		for (int64_t i = 0; i < n; ++i) {
			bool haveDiag = false;
			for (int64_t p = prp[(size_t)i]; p < prp[(size_t)i + 1]; ++p) {
				if (pci[(size_t)p] < i) { lci.push_back(pci[(size_t)p]); lva.insert(lva.end(), pva.begin() + p * BB, pva.begin() + (p + 1) * BB); }
				else { if (pci[(size_t)p] == i) haveDiag = true; uci.push_back(pci[(size_t)p]); uva.insert(uva.end(), pva.begin() + p * BB, pva.begin() + (p + 1) * BB); }
			}
			if (!haveDiag) UG_THROW("ILU: row " << i << " has no diagonal entry");
			lci.push_back((int)i);
			for (int t = 0; t < BB; ++t) lva.push_back((t % B) == (t / B) ? 1.0 : 0.0);   // unit diagonal block
			lrp[(size_t)i + 1] = (int64_t)lci.size(); urp[(size_t)i + 1] = (int64_t)uci.size();
		}
		// ---- groups of independent rows ----
		std::vector<size_t> permL((size_t)n), permU((size_t)n);
		bool identity = true;
		if (!colorPtr.empty()) {
			if (ug4b200_color_check(n, prp.data(), pci.data(), (int)colorPtr.size() - 1, colorPtr.data()) != 0)
				UG_THROW("ILU: the given ordering is not a valid multicolour ordering of the matrix pattern");
			for (int64_t i = 0; i < n; ++i) permL[(size_t)i] = permU[(size_t)i] = (size_t)i;
			m_ptrL = colorPtr; m_ptrU = colorPtr;
		} else {
			std::vector<int> lev;
			const int nl = level_sets(n, lrp, lci, true, lev);
			sort_by_level(n, lev, nl, false, permL, m_ptrL);
			const int nu = level_sets(n, urp, uci, false, lev);
			sort_by_level(n, lev, nu, true, permU, m_ptrU);
			for (int64_t i = 0; i < n && identity; ++i) identity = permL[(size_t)i] == (size_t)i && permU[(size_t)i] == (size_t)i;
			if (!identity) {
				std::vector<int64_t> rp2; std::vector<int> ci2; std::vector<double> va2;
				permute(n, lrp, lci, lva, permL, permL, rp2, ci2, va2); lrp.swap(rp2); lci.swap(ci2); lva.swap(va2);
				permute(n, urp, uci, uva, permU, permU, rp2, ci2, va2); urp.swap(rp2); uci.swap(ci2); uva.swap(va2);
			}
		}
		// ---- upload ----
		free_dev();
		GPUManager::bump_generation();   // the device buffers a captured solver graph points at are rebuilt
		ug4b200_ctx* ctx = GPUManager::ctx();
		UG_GPU_CHECK(ug4b200_matrix_upload_crs(ctx, B, n, n, lrp.data(), lci.data(), lva.data(), UG4B200_MAT_DEFAULT, &m_L));
		UG_GPU_CHECK(ug4b200_matrix_upload_crs(ctx, B, n, n, urp.data(), uci.data(), uva.data(), UG4B200_MAT_DEFAULT, &m_U));
		m_n = (size_t)n;
		std::vector<int> mapIn((size_t)n), mapLU((size_t)n), mapOut((size_t)n);
		for (int64_t i = 0; i < n; ++i) {
			mapIn[(size_t)i] = (int)permL[ord[(size_t)i]];             // dL[mapIn[i]] = d[i]
			mapOut[(size_t)i] = (int)permU[ord[(size_t)i]];            // c[i] = cU[mapOut[i]]
			mapLU[permU[(size_t)i]] = (int)permL[(size_t)i];           // yU[j] = yL[mapLU[j]]
		}
		m_bSameOrder = identity;
		m_dIn = upload_ints(mapIn); m_dOut = upload_ints(mapOut); m_dLU = m_bSameOrder ? nullptr : upload_ints(mapLU);
		UG_GPU_CHECK(ug4b200_sync(ctx));
		m_t0 = GPUManager::alloc(m_n * B); m_t1 = GPUManager::alloc(m_n * B);
		if (m_layouts) { m_dUnique.create(m_n); m_dUnique.set_layouts(m_layouts); }
		return true;
	}

	/// ILU::step (ilu.h:614-655) with applyLU (:591-611)
	virtual bool step(SmartPtr<matrix_operator_type>, vector_type& c, const vector_type& d)
	{
		ug4b200_ctx* ctx = GPUManager::ctx();
		if (m_n == 0) return true;
		const double* dsrc = d.dev();
		if (m_layouts) {
			m_dUnique = d;                                                  // make defect unique (:640-644)
			if (!m_dUnique.change_storage_type(PST_UNIQUE)) UG_THROW("ILU: cannot make the defect unique");
			dsrc = m_dUnique.dev();
		}
		const int64_t n = (int64_t)m_n;
		UG_GPU_CHECK(ug4b200_vec_scatter(ctx, n, B, m_t0, m_dIn, dsrc));                                            // SetVectorAsPermutation(tmp, d, ordering)
		UG_GPU_CHECK(ug4b200_gs_step(ctx, m_L, (int)m_ptrL.size() - 1, m_ptrL.data(), 0, 1.0, m_t1, m_t0));           // invert_L
		if (m_bSameOrder) {
			UG_GPU_CHECK(ug4b200_gs_step(ctx, m_U, (int)m_ptrU.size() - 1, m_ptrU.data(), 1, 1.0, m_t0, m_t1));       // invert_U
		} else {
			UG_GPU_CHECK(ug4b200_vec_gather(ctx, n, B, m_t0, m_t1, m_dLU));
			UG_GPU_CHECK(ug4b200_gs_step(ctx, m_U, (int)m_ptrU.size() - 1, m_ptrU.data(), 1, 1.0, m_t1, m_t0));
			std::swap(m_t0, m_t1);
		}
		UG_GPU_CHECK(ug4b200_vec_gather(ctx, n, B, c.dev(), m_t0, m_dOut));                                          // SetVectorAsPermutation(c, tmp, old_ordering)
		if (m_layouts) {
			c.set_storage_type(PST_ADDITIVE);                               // :646
			if (!c.change_storage_type(PST_CONSISTENT)) return false;       // :652
		} else c.set_storage_type(PST_CONSISTENT);
		return true;
	}
	virtual bool postprocess() { return true; }

	/// B(pr[r], pc[c]) = A(r, c), rows sorted
	static void permute(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, const std::vector<double>& va,
	                    const std::vector<size_t>& pr, const std::vector<size_t>& pc, std::vector<int64_t>& orp, std::vector<int>& oci,
	                    std::vector<double>& ova)
	{
		const int BB = B * B;
		std::vector<size_t> inv((size_t)n);
		for (int64_t i = 0; i < n; ++i) inv[pr[(size_t)i]] = (size_t)i;
		orp.assign((size_t)n + 1, 0);
		for (int64_t nr = 0; nr < n; ++nr) orp[(size_t)nr + 1] = orp[(size_t)nr] + (rp[inv[(size_t)nr] + 1] - rp[inv[(size_t)nr]]);
		oci.resize(ci.size()); ova.resize(va.size());
		std::vector<std::pair<int, int64_t> > row;
		for (int64_t nr = 0; nr < n; ++nr) {
			const size_t r = inv[(size_t)nr];
			row.clear();
			for (int64_t p = rp[r]; p < rp[r + 1]; ++p) row.push_back(std::make_pair((int)pc[(size_t)ci[(size_t)p]], p));
			std::sort(row.begin(), row.end());
			int64_t q = orp[(size_t)nr];
			for (size_t k = 0; k < row.size(); ++k, ++q) {
				oci[(size_t)q] = row[k].first;
				for (int t = 0; t < BB; ++t) ova[(size_t)q * BB + t] = va[(size_t)row[k].second * BB + t];
			}
		}
	}
	/// rows grouped by level, stable inside a level; descending = true puts the highest level first (U: the
	/// backward sweep walks the groups from the last to the first)
	static void sort_by_level(int64_t n, const std::vector<int>& lev, int nlev, bool descending, std::vector<size_t>& perm,
	                          std::vector<int64_t>& ptr)
	{
		ptr.assign((size_t)nlev + 1, 0);
		for (int64_t i = 0; i < n; ++i) { const int g = descending ? nlev - 1 - lev[(size_t)i] : lev[(size_t)i]; ptr[(size_t)g + 1]++; }
		for (int g = 0; g < nlev; ++g) ptr[(size_t)g + 1] += ptr[(size_t)g];
		std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
		for (int64_t i = 0; i < n; ++i) { const int g = descending ? nlev - 1 - lev[(size_t)i] : lev[(size_t)i]; perm[(size_t)i] = (size_t)fill[(size_t)g]++; }
	}
	static int* upload_ints(const std::vector<int>& v)
	{
		int* d = (int*)GPUManager::alloc_bytes(sizeof(int) * (v.empty() ? 1 : v.size()));
		UG_GPU_CHECK(ug4b200_h2d(GPUManager::ctx(), d, v.data(), sizeof(int) * v.size()));
		return d;
	}
	void free_dev()
	{
		ug4b200_ctx* c = GPUManager::ctx_or_null();
		if (m_L && c) ug4b200_matrix_destroy(c, m_L);
		if (m_U && c) ug4b200_matrix_destroy(c, m_U);
		m_L = m_U = nullptr;
		GPUManager::free_bytes(m_dIn); GPUManager::free_bytes(m_dOut); GPUManager::free_bytes(m_dLU);
		m_dIn = m_dOut = m_dLU = nullptr;
		if (m_t0) GPUManager::release(m_t0, m_n * B);
		if (m_t1) GPUManager::release(m_t1, m_n * B);
		m_t0 = m_t1 = nullptr;
	}

	double m_beta;
	number m_sortEps = 1e-50, m_invEps = 1e-8;
	bool m_bDisablePreprocessing = false;
	int m_mode = ORDER_NONE;
	ordering_container_type m_userOrdering;
	std::vector<int64_t> m_userColorPtr;
	ug4b200_matrix *m_L = nullptr, *m_U = nullptr;
	std::vector<int64_t> m_ptrL, m_ptrU;
	int *m_dIn = nullptr, *m_dOut = nullptr, *m_dLU = nullptr;
	bool m_bSameOrder = true;
	double *m_t0 = nullptr, *m_t1 = nullptr;
	size_t m_n = 0;
	SmartPtr<GPUAlgebraLayouts> m_layouts;
	SmartPtr<matrix_type> m_spConsistent;
	vector_type m_dUnique;
};

template <typename TAlgebra>
class GaussSeidel : public GaussSeidelBase<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	virtual const char* name() const { return "Gauss-Seidel"; }
	virtual SmartPtr<ILinearIterator<vector_type> > clone() { SmartPtr<GaussSeidel> g(new GaussSeidel()); this->copy_settings(*g); return g; }
  protected:
	virtual int kind() const { return 0; }
};
template <typename TAlgebra>
class BackwardGaussSeidel : public GaussSeidelBase<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	virtual const char* name() const { return "Backward Gauss-Seidel"; }
	virtual SmartPtr<ILinearIterator<vector_type> > clone() { SmartPtr<BackwardGaussSeidel> g(new BackwardGaussSeidel()); this->copy_settings(*g); return g; }
  protected:
	virtual int kind() const { return 1; }
};
template <typename TAlgebra>
class SymmetricGaussSeidel : public GaussSeidelBase<TAlgebra> {
  public:
	typedef typename TAlgebra::vector_type vector_type;
	virtual const char* name() const { return "Symmetric Gauss-Seidel"; }
	virtual SmartPtr<ILinearIterator<vector_type> > clone() { SmartPtr<SymmetricGaussSeidel> g(new SymmetricGaussSeidel()); this->copy_settings(*g); return g; }
  protected:
	virtual int kind() const { return 2; }
};

} // namespace ug
