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
#pragma once
#include "operators.h"

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
	/// ordering: old index i -> new index perm[i]; the new order must be colour-sorted with the
	/// given colour offsets.  Without it a greedy colouring of the stored pattern is used.
	void set_coloring(const std::vector<int>& perm, const std::vector<int64_t>& colorPtr) { m_perm = perm; m_colorPtr = colorPtr; }
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
		if (m_perm.size() != n) {
			// greedy colouring in row order, colours sorted, stable inside a colour
			std::vector<int> color(n); int nc = 0;
			ug4b200_color_greedy((int64_t)n, rp.data(), ci.data(), color.data(), &nc);
			m_colorPtr.assign(nc + 1, 0);
			for (size_t i = 0; i < n; ++i) m_colorPtr[color[i] + 1]++;
			for (int k = 0; k < nc; ++k) m_colorPtr[k + 1] += m_colorPtr[k];
			std::vector<int64_t> fill(m_colorPtr.begin(), m_colorPtr.end() - 1);
			m_perm.resize(n);
			for (size_t i = 0; i < n; ++i) m_perm[i] = (int)fill[color[i]]++;
			m_userPerm = false;
		} else m_userPerm = true;
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
