// gpu_sparsematrix.h — GPUSparseMatrix<T>: host-assembled sparse matrix with an immutable
// device mirror (SELL-32) that is uploaded on first device use.
//
// Mirrors the part of SparseMatrix<T> (ugbase/lib_algebra/cpu_algebra/sparsematrix.h:116-343)
// that assembly and the solve path touch; storage on the host is sorted-row CRS (what
// SparseMatrix holds after defragment(), sparsematrix.h:588-592).  Shape of the drop-in:
// legacy GPUSparseMatrix::copy_to_device()/check_device()
// (ugbase/lib_algebra/gpu_algebra/gpusparsematrix.h:614-651).  In partitioned runs the
// matrix is stored additive (ParallelMatrix, parallel_matrix_impl.h:88-114): SpMV needs
// no communication and maps consistent -> additive.
#pragma once
#include "gpu_vector.h"
#include <algorithm>

namespace ug {

template <typename T> struct gpu_value_access;
template <> struct gpu_value_access<double> {
	enum { B = 1 };
	static double get(const double& v, int, int) { return v; }
	static void set(double& v, int, int, double x) { v = x; }
};
template <size_t N> struct gpu_value_access<DenseMatrix<FixedArray2<double, N, N> > > {
	enum { B = N };
	static double get(const DenseMatrix<FixedArray2<double, N, N> >& v, int r, int c) { return v(r, c); }
	static void set(DenseMatrix<FixedArray2<double, N, N> >& v, int r, int c, double x) { v(r, c) = x; }
};

template <typename TValueType>
class GPUSparseMatrix {
  public:
	typedef TValueType value_type;
	typedef GPUSparseMatrix<TValueType> this_type;
	enum { blockSize = gpu_value_access<TValueType>::B };
	struct connection { size_t iIndex; value_type dValue; };

	GPUSparseMatrix() {}
	virtual ~GPUSparseMatrix() { drop_device(); }
	GPUSparseMatrix(const GPUSparseMatrix&) = delete;
	GPUSparseMatrix& operator=(const GPUSparseMatrix&) = delete;

	// ---- assembly-side API (host) ----
	void resize_and_clear(size_t rows, size_t cols)
	{
		m_rows.assign(rows, std::vector<connection>()); m_numCols = cols; m_fragmented = true; m_crsRows = 0;
		m_rowptr.clear(); m_cols.clear(); m_vals.clear(); touch();
	}
	size_t num_rows() const { return m_fragmented ? m_rows.size() : m_crsRows; }
	size_t num_cols() const { return m_numCols; }
	size_t total_num_connections() const { const_cast<this_type*>(this)->defragment(); return m_cols.size(); }
	size_t num_connections(size_t r) const { const_cast<this_type*>(this)->defragment(); return (size_t)(m_rowptr[r + 1] - m_rowptr[r]); }

	/// inserting access (sparsematrix_impl.h:596-700): keeps rows sorted
	value_type& operator()(size_t r, size_t c)
	{
		fragment(); touch();
		std::vector<connection>& row = m_rows[r];
		auto it = std::lower_bound(row.begin(), row.end(), c, [](const connection& a, size_t cc) { return a.iIndex < cc; });
		if (it == row.end() || it->iIndex != c) { connection n; n.iIndex = c; n.dValue = value_type(); zero(n.dValue); it = row.insert(it, n); }
		return it->dValue;
	}
	/// sparsematrix_impl.h:429-460: the given connections are set, other connections of the row stay
	void set_matrix_row(size_t r, connection* c, size_t nr)
	{ for (size_t i = 0; i < nr; ++i) (*this)(r, c[i].iIndex) = c[i].dValue; }
	/// sparsematrix_impl.h:463-468
	void add_matrix_row(size_t row, connection* c, size_t nr)
	{ for (size_t i = 0; i < nr; ++i) add_value((*this)(row, c[i].iIndex), c[i].dValue); }
	/// const access: 0 if the connection is not stored (sparsematrix.h:272-283)
	const value_type& operator()(size_t r, size_t c) const
	{
		static const value_type zeroValue = make_zero();
		const this_type* self = this;
		const_cast<this_type*>(self)->fragment();
		const std::vector<connection>& row = m_rows[r];
		auto it = std::lower_bound(row.begin(), row.end(), c, [](const connection& a, size_t cc) { return a.iIndex < cc; });
		return (it == row.end() || it->iIndex != c) ? zeroValue : it->dValue;
	}
	bool has_connection(size_t r, size_t c) const
	{
		const_cast<this_type*>(this)->fragment();
		const std::vector<connection>& row = m_rows[r];
		auto it = std::lower_bound(row.begin(), row.end(), c, [](const connection& a, size_t cc) { return a.iIndex < cc; });
		return it != row.end() && it->iIndex == c;
	}
	void clear_and_free() { resize_and_clear(0, 0); }
	/// sparsematrix_impl.h:112-139: new rows are empty, connections to columns >= newCols disappear
	void resize_and_keep_values(size_t newRows, size_t newCols)
	{
		if (newRows == 0 && newCols == 0) return resize_and_clear(0, 0);
		fragment(); touch();
		m_rows.resize(newRows);
		if (newCols < m_numCols)
			for (std::vector<connection>& row : m_rows)
				while (!row.empty() && row.back().iIndex >= newCols) row.pop_back();
		m_numCols = newCols;
	}
	/// all values 0, pattern kept (sparsematrix_impl.h:142-145)
	void clear_retain_structure() { fragment(); touch(); for (std::vector<connection>& row : m_rows) for (connection& c : row) zero(c.dValue); }
	/// sparsematrix_impl.h:472-482
	void set_as_copy_of(const this_type& Bm, double scaleFactor = 1.0)
	{
		const_cast<this_type&>(Bm).fragment();
		resize_and_clear(Bm.num_rows(), Bm.num_cols());
		for (size_t i = 0; i < Bm.m_rows.size(); ++i)
			for (const connection& c : Bm.m_rows[i]) { value_type v = c.dValue; scale_value(v, scaleFactor); (*this)(i, c.iIndex) = v; }
	}
	/// sparsematrix_impl.h:487-496
	void scale(double d) { fragment(); touch(); for (std::vector<connection>& row : m_rows) for (connection& c : row) scale_value(c.dValue, d); }
	/// diagonal entries := a, all others := 0, pattern kept (sparsematrix_impl.h:397-412)
	void set(double a)
	{
		fragment(); touch();
		for (size_t r = 0; r < m_rows.size(); ++r) for (connection& c : m_rows[r]) { if (c.iIndex == r) c.dValue = a; else c.dValue = 0.0; }
	}
	/// no non-zero connection to another index (sparsematrix_impl.h:416-425)
	bool is_isolated(size_t i) const
	{
		const_cast<this_type*>(this)->fragment();
		for (const connection& c : m_rows[i]) if (c.iIndex != i && !is_zero(c.dValue)) return false;
		return true;
	}
	/// local (element) matrices: M offers num_rows/num_cols, row_index(i), col_index(j), operator()(i,j)
	/// (sparsematrix_impl.h:505-548)
	template <typename M> void add(const M& mat)
	{ for (size_t i = 0; i < mat.num_rows(); ++i) for (size_t j = 0; j < mat.num_cols(); ++j) add_value((*this)(mat.row_index(i), mat.col_index(j)), mat(i, j)); }
	template <typename M> void set(const M& mat)
	{ for (size_t i = 0; i < mat.num_rows(); ++i) for (size_t j = 0; j < mat.num_cols(); ++j) (*this)(mat.row_index(i), mat.col_index(j)) = mat(i, j); }
	template <typename M> void get(M& mat) const
	{ for (size_t i = 0; i < mat.num_rows(); ++i) for (size_t j = 0; j < mat.num_cols(); ++j) mat(i, j) = (*this)(mat.row_index(i), mat.col_index(j)); }
	/// row access in ascending column order (begin_row / end_row, sparsematrix.h:484-487); the iterators offer
	/// index() and value() like ugcore's
	struct row_iterator {
		typename std::vector<connection>::iterator it;
		size_t index() const { return it->iIndex; }
		value_type& value() { return it->dValue; }
		bool operator!=(const row_iterator& o) const { return it != o.it; }
		bool operator==(const row_iterator& o) const { return it == o.it; }
		void operator++() { ++it; }
	};
	struct const_row_iterator {
		typename std::vector<connection>::const_iterator it;
		size_t index() const { return it->iIndex; }
		const value_type& value() const { return it->dValue; }
		bool operator!=(const const_row_iterator& o) const { return it != o.it; }
		bool operator==(const const_row_iterator& o) const { return it == o.it; }
		void operator++() { ++it; }
	};
	row_iterator begin_row(size_t r) { fragment(); touch(); return row_iterator{m_rows[r].begin()}; }
	row_iterator end_row(size_t r) { fragment(); return row_iterator{m_rows[r].end()}; }
	const_row_iterator begin_row(size_t r) const { const_cast<this_type*>(this)->fragment(); return const_row_iterator{m_rows[r].begin()}; }
	const_row_iterator end_row(size_t r) const { const_cast<this_type*>(this)->fragment(); return const_row_iterator{m_rows[r].end()}; }
	/// bulk load of a defragmented CRS (what copy_crs exports, sparsematrix.h:607-617);
	/// vals: block*block doubles per entry, column-major inside a block
	void set_from_crs(size_t rows, size_t cols, const int64_t* rowptr, const int* colidx, const double* vals)
	{
		m_rows.clear(); m_fragmented = false; m_crsRows = rows; m_numCols = cols;
		m_rowptr.assign(rowptr, rowptr + rows + 1);
		m_cols.assign(colidx, colidx + rowptr[rows]);
		m_vals.assign(vals, vals + (size_t)rowptr[rows] * blockSize * blockSize);
		touch();
	}
	void defragment()
	{
		if (!m_fragmented) return;
		const size_t n = m_rows.size(); const int B = blockSize, BB = B * B;
		m_rowptr.assign(n + 1, 0);
		for (size_t r = 0; r < n; ++r) m_rowptr[r + 1] = m_rowptr[r] + (int64_t)m_rows[r].size();
		m_cols.resize(m_rowptr[n]); m_vals.resize((size_t)m_rowptr[n] * BB);
		for (size_t r = 0; r < n; ++r) {
			int64_t p = m_rowptr[r];
			for (const connection& c : m_rows[r]) {
				m_cols[p] = (int)c.iIndex;
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) m_vals[p * BB + i + B * j] = gpu_value_access<value_type>::get(c.dValue, i, j);
				++p;
			}
		}
		m_crsRows = n; m_rows.clear(); m_fragmented = false;
	}
	/// set_as_transpose_of (sparsematrix_impl.h:148-183): explicit zeros are kept
	void set_as_transpose_of(const this_type& Bm, double scale = 1.0)
	{
		const_cast<this_type&>(Bm).defragment();
		const int B = blockSize, BB = B * B;
		const size_t nr = Bm.num_cols(), nc = Bm.num_rows();
		std::vector<int64_t> rp(nr + 1, 0);
		for (int c : Bm.m_cols) rp[c + 1]++;
		for (size_t r = 0; r < nr; ++r) rp[r + 1] += rp[r];
		std::vector<int> ci(Bm.m_cols.size()); std::vector<double> va(Bm.m_vals.size());
		std::vector<int64_t> fill(rp.begin(), rp.end() - 1);
		for (size_t r = 0; r < nc; ++r)
			for (int64_t p = Bm.m_rowptr[r]; p < Bm.m_rowptr[r + 1]; ++p) {
				const int64_t q = fill[Bm.m_cols[p]]++;
				ci[q] = (int)r;
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) va[q * BB + i + B * j] = scale * Bm.m_vals[p * BB + j + B * i];
			}
		set_from_crs(nr, nc, rp.data(), ci.data(), va.data());
	}
	const std::vector<int64_t>& crs_rowptr() const { const_cast<this_type*>(this)->defragment(); return m_rowptr; }
	const std::vector<int>& crs_cols() const { const_cast<this_type*>(this)->defragment(); return m_cols; }
	const std::vector<double>& crs_vals() const { const_cast<this_type*>(this)->defragment(); return m_vals; }

	// ---- device mirror ----
	const ug4b200_matrix* device() const
	{
		this_type* s = const_cast<this_type*>(this);
		if (!s->m_dev) {
			s->defragment();
			UG_GPU_CHECK(ug4b200_matrix_upload_crs(GPUManager::ctx(), blockSize, (int64_t)m_crsRows, (int64_t)m_numCols,
			                                       m_rowptr.data(), m_cols.data(), m_vals.data(), UG4B200_MAT_DEFAULT, &s->m_dev));
		}
		return m_dev;
	}
	/// free the host copy once uploaded (large level matrices: "uploaded once after assembly")
	void release_host() { device(); std::vector<int>().swap(m_cols); std::vector<double>().swap(m_vals); m_hostReleased = true; }

	// ---- solve-path API: all on the device (sparsematrix.h:184-230) ----
	template <typename V> bool apply(V& res, const V& x) const
	{
		UG_GPU_CHECK(ug4b200_matrix_apply(GPUManager::ctx(), device(), res.dev(), x.dev(), V::blockSize));
		res.set_storage_type(PST_ADDITIVE); // parallel_matrix_impl.h:105-114
		return true;
	}
	template <typename V> bool matmul_minus(V& res, const V& x) const
	{
		UG_GPU_CHECK(ug4b200_matrix_matmul_minus(GPUManager::ctx(), device(), res.dev(), x.dev(), V::blockSize));
		res.set_storage_type(PST_ADDITIVE); // additive - A_additive * consistent: no longer unique
		return true;
	}
	template <typename V> bool axpy(V& dest, const number& alpha1, const V& v1, const number& beta1, const V& w1) const
	{
		UG_GPU_ZONE(SparseMatrix_axpy);                               // sparsematrix_impl.h:298
		UG_GPU_CHECK(ug4b200_matrix_axpy(GPUManager::ctx(), device(), dest.dev(), alpha1, alpha1 == 0.0 ? nullptr : v1.dev(), beta1,
		                                 w1.dev(), V::blockSize));
		return true;
	}
	/// dest = alpha1 * v1 + beta1 * A^T * w1 (sparsematrix_impl.h:341-370; sparsematrix.h:194-197 apply_transposed).
	/// The reference scatters row by row; the explicit transpose (built and uploaded on first use, dropped when the
	/// matrix changes) accumulates every entry of dest over the same terms in the same (ascending row) order.
	template <typename V> bool axpy_transposed(V& dest, const number& alpha1, const V& v1, const number& beta1, const V& w1) const
	{
		if (!m_spTransposed) {
			UG_COND_THROW(m_hostReleased, "GPUSparseMatrix: host copy was released before the transpose was built");
			m_spTransposed.reset(new this_type()); m_spTransposed->set_as_transpose_of(*this);
		}
		return m_spTransposed->axpy(dest, alpha1, v1, beta1, w1);
	}
	template <typename V> bool apply_transposed(V& res, const V& x) const
	{
		if (!axpy_transposed(res, 0.0, res, 1.0, x)) return false;
		res.set_storage_type(PST_ADDITIVE);
		return true;
	}
	template <typename V> bool apply_ignore_zero_rows(V& dest, const number& beta1, const V& w1) const
	{
		UG_GPU_CHECK(ug4b200_matrix_apply_ignore_zero_rows(GPUManager::ctx(), device(), dest.dev(), beta1, w1.dev(), V::blockSize));
		return true;
	}

  private:
	static void zero(double& v) { v = 0.0; }
	template <class X> static void zero(X& v) { v = 0.0; }
	static value_type make_zero() { value_type v = value_type(); zero(v); return v; }
	static void add_value(value_type& a, const value_type& b)
	{ for (int i = 0; i < blockSize; ++i) for (int j = 0; j < blockSize; ++j) gpu_value_access<value_type>::set(a, i, j, gpu_value_access<value_type>::get(a, i, j) + gpu_value_access<value_type>::get(b, i, j)); }
	static void scale_value(value_type& a, double d)
	{ for (int i = 0; i < blockSize; ++i) for (int j = 0; j < blockSize; ++j) gpu_value_access<value_type>::set(a, i, j, gpu_value_access<value_type>::get(a, i, j) * d); }
	static bool is_zero(const value_type& a)
	{ for (int i = 0; i < blockSize; ++i) for (int j = 0; j < blockSize; ++j) if (gpu_value_access<value_type>::get(a, i, j) != 0.0) return false; return true; }
	void touch() { drop_device(); UG_COND_THROW(m_hostReleased, "GPUSparseMatrix: host copy was released, matrix is immutable"); }
	void drop_device()
	{
		if (m_dev) GPUManager::bump_generation();   // graphs that captured this mirror must not be replayed
		if (m_dev && GPUManager::ctx_or_null()) ug4b200_matrix_destroy(GPUManager::ctx_or_null(), m_dev);
		m_dev = nullptr;
		m_spTransposed.reset();
	}
	void fragment()
	{
		if (m_fragmented) return;
		const int B = blockSize, BB = B * B;
		m_rows.assign(m_crsRows, std::vector<connection>());
		for (size_t r = 0; r < m_crsRows; ++r)
			for (int64_t p = m_rowptr[r]; p < m_rowptr[r + 1]; ++p) {
				connection c; c.iIndex = (size_t)m_cols[p];
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) gpu_value_access<value_type>::set(c.dValue, i, j, m_vals[p * BB + i + B * j]);
				m_rows[r].push_back(c);
			}
		m_fragmented = true;
	}

	// fragmented (assembly) form
	std::vector<std::vector<connection> > m_rows;
	bool m_fragmented = false;
	// defragmented CRS
	size_t m_crsRows = 0, m_numCols = 0;
	std::vector<int64_t> m_rowptr;
	std::vector<int> m_cols;
	std::vector<double> m_vals;
	bool m_hostReleased = false;
	ug4b200_matrix* m_dev = nullptr;
	mutable std::unique_ptr<this_type> m_spTransposed;   // explicit transpose for axpy_transposed / apply_transposed
};

/// CPUAlgebra / CPUBlockAlgebra<N> counterparts (ugbase/lib_algebra/cpu_algebra_types.h:76-142;
/// the commented-out GPUAlgebra stub :101-120)
struct GPUAlgebra {
	typedef GPUSparseMatrix<double> matrix_type;
	typedef GPUVector<double> vector_type;
	static const int blockSize = 1;
	static AlgebraType get_type() { return AlgebraType(AlgebraType::GPU, 1); }
};
template <int TBlockSize> struct GPUBlockAlgebra {
	typedef GPUSparseMatrix<DenseMatrix<FixedArray2<double, TBlockSize, TBlockSize> > > matrix_type;
	typedef GPUVector<DenseVector<FixedArray1<double, TBlockSize> > > vector_type;
	static const int blockSize = TBlockSize;
	static AlgebraType get_type() { return AlgebraType(AlgebraType::GPU, TBlockSize); }
};

} // namespace ug
