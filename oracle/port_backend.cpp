/*
 * oracle/port_backend.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain C++ restatement ("port") of the arithmetic of ugcore's CPU algebra on
 * the GMG/Krylov hot path.  No reference headers are included; every routine
 * cites the reference lines it restates (paths relative to
 * /root/reference/ugbase).  Compile WITHOUT -ffast-math / -march so that no FMA
 * contraction happens (reference release flags, cmake/ug/debug.cmake:76-88).
 *
 * Pinned against: tests/ref/sm_transpose.out (transpose nnz), the FV1
 * known-answer stencils, and — when /root/reference is present — bit-for-bit
 * against the compiled reference templates (oracle/_ref, RefBackend).
 */
#include "backend.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace oracle {
namespace {

struct PMat : Mat {
	std::vector<int64_t> rp;
	std::vector<int> ci;
	std::vector<double> va; // block*block per entry, column-major (fixed_array_impl.h:182-203)
};
struct PVec : Vec {
	std::vector<double> v;
	double* data() override { return v.data(); }
	const double* data() const override { return v.data(); }
};
struct PDiag : DiagInv { int B; std::vector<double> inv; };
struct PLU : DenseLU { int64_t n; std::vector<double> a; std::vector<size_t> piv; };

inline const PMat& M(const Mat& m) { return static_cast<const PMat&>(m); }

// dest = beta*A*w   (small_matrix/densematrix_operations.h:56-65; scalar: matrix_use_operators.h:43-47)
inline void blk_mult(int B, double* dest, double beta, const double* A, const double* w)
{
	for (int r = 0; r < B; ++r) {
		dest[r] = beta * A[r] * w[0];
		for (int c = 1; c < B; ++c) dest[r] = 1.0 * dest[r] + beta * A[r + B * c] * w[c];
	}
}
// dest = 1.0*dest + beta*A*w  (densematrix_operations.h:69-79 with alpha1=1, v1=dest)
inline void blk_mult_add(int B, double* dest, double beta, const double* A, const double* w)
{
	for (int r = 0; r < B; ++r) {
		dest[r] = 1.0 * dest[r];
		for (int c = 0; c < B; ++c) dest[r] = 1.0 * dest[r] + beta * A[r + B * c] * w[c];
	}
}

// dest = beta * mat^{-1} vec  (double.h:157-161; densematrix_inverse.h:78-86,137-147,229-245)
inline bool blk_inverse_mult(int B, double* dest, double beta, const double* m, const double* v)
{
#define MM(r, c) m[(r) + B * (c)]
	if (B == 1) { if (m[0] == 0.0) return false; dest[0] = beta * v[0] / m[0]; return true; }
	if (B == 2) {
		const double det = MM(0,0) * MM(1,1) - MM(1,0) * MM(0,1);
		if (det == 0.0) return false;
		dest[0] = beta * (MM(1,1) * v[0] - MM(0,1) * v[1]) / det;
		dest[1] = beta * (-MM(1,0) * v[0] + MM(0,0) * v[1]) / det;
		return true;
	}
	const double det = MM(0,0)*MM(1,1)*MM(2,2) + MM(0,1)*MM(1,2)*MM(2,0) + MM(0,2)*MM(1,0)*MM(2,1)
	                 - MM(0,0)*MM(1,2)*MM(2,1) - MM(0,1)*MM(1,0)*MM(2,2) - MM(0,2)*MM(1,1)*MM(2,0);
	if (det == 0.0) return false;
	dest[0] = (( MM(1,1)*MM(2,2) - MM(1,2)*MM(2,1)) * v[0] +
	           (-MM(0,1)*MM(2,2) + MM(0,2)*MM(2,1)) * v[1] +
	           ( MM(0,1)*MM(1,2) - MM(0,2)*MM(1,1)) * v[2]) * beta / det;
	dest[1] = ((-MM(1,0)*MM(2,2) + MM(1,2)*MM(2,0)) * v[0] +
	           ( MM(0,0)*MM(2,2) - MM(0,2)*MM(2,0)) * v[1] +
	           (-MM(0,0)*MM(1,2) + MM(0,2)*MM(1,0)) * v[2]) * beta / det;
	dest[2] = (( MM(1,0)*MM(2,1) - MM(1,1)*MM(2,0)) * v[0] +
	           (-MM(0,0)*MM(2,1) + MM(0,1)*MM(2,0)) * v[1] +
	           ( MM(0,0)*MM(1,1) - MM(0,1)*MM(1,0)) * v[2]) * beta / det;
	return true;
#undef MM
}

// inv = m^{-1}  (double.h:197-201; densematrix_inverse.h:46-51,96-108,163-181)
inline bool blk_get_inverse(int B, double* inv, const double* m)
{
#define MM(r, c) m[(r) + B * (c)]
#define II(r, c) inv[(r) + B * (c)]
	if (B == 1) { inv[0] = 1.0 / m[0]; return m[0] != 0.0; }
	if (B == 2) {
		double invdet = MM(0,0) * MM(1,1) - MM(1,0) * MM(0,1);
		if (invdet == 0.0) return false;
		invdet = 1.0 / invdet;
		II(0,0) = MM(1,1) * invdet; II(1,1) = MM(0,0) * invdet;
		II(0,1) = MM(0,1) * -invdet; II(1,0) = MM(1,0) * -invdet;
		return true;
	}
	double invdet = MM(0,0)*MM(1,1)*MM(2,2) + MM(0,1)*MM(1,2)*MM(2,0) + MM(0,2)*MM(1,0)*MM(2,1)
	              - MM(0,0)*MM(1,2)*MM(2,1) - MM(0,1)*MM(1,0)*MM(2,2) - MM(0,2)*MM(1,1)*MM(2,0);
	if (invdet == 0.0) return false;
	invdet = 1.0 / invdet;
	II(0,0) = ( MM(1,1)*MM(2,2) - MM(1,2)*MM(2,1)) * invdet;
	II(0,1) = (-MM(0,1)*MM(2,2) + MM(0,2)*MM(2,1)) * invdet;
	II(0,2) = ( MM(0,1)*MM(1,2) - MM(0,2)*MM(1,1)) * invdet;
	II(1,0) = (-MM(1,0)*MM(2,2) + MM(1,2)*MM(2,0)) * invdet;
	II(1,1) = ( MM(0,0)*MM(2,2) - MM(0,2)*MM(2,0)) * invdet;
	II(1,2) = (-MM(0,0)*MM(1,2) + MM(0,2)*MM(1,0)) * invdet;
	II(2,0) = ( MM(1,0)*MM(2,1) - MM(1,1)*MM(2,0)) * invdet;
	II(2,1) = (-MM(0,0)*MM(2,1) + MM(0,1)*MM(2,0)) * invdet;
	II(2,2) = ( MM(0,0)*MM(1,1) - MM(0,1)*MM(1,0)) * invdet;
	return true;
#undef MM
#undef II
}

struct PortBackend : Backend {
	const char* name() const override { return "port"; }

	Mat* matrix(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
	            const double* vals) override
	{
		PMat* m = new PMat;
		m->nrows = nrows; m->ncols = ncols; m->block = block;
		m->rp.assign(rowptr, rowptr + nrows + 1);
		const int64_t nnz = rowptr[nrows];
		m->ci.assign(cols, cols + nnz);
		m->va.assign(vals, vals + nnz * block * block);
		return m;
	}
	Vec* vector(int64_t nblocks, int block) override
	{
		PVec* v = new PVec; v->n = nblocks; v->block = block; v->v.assign((size_t)nblocks * block, 0.0);
		return v;
	}
	int64_t nnz(const Mat& A) override { return (int64_t)M(A).ci.size(); }
	void export_crs(const Mat& A_, int64_t* rowptr, int* cols, double* vals) override
	{
		const PMat& A = M(A_);
		std::copy(A.rp.begin(), A.rp.end(), rowptr);
		std::copy(A.ci.begin(), A.ci.end(), cols);
		std::copy(A.va.begin(), A.va.end(), vals);
	}

	// cpu_algebra/sparsematrix_impl.h:148-183 (set_as_transpose_of: every stored
	// connection is re-inserted, explicit zeros included) and set_as_transpose_of2
	// (zero-valued connections are dropped); block entries are transposed.
	Mat* transpose(const Mat& A_, bool keep_zeros) override
	{
		const PMat& A = M(A_);
		const int B = A.block, BB = B * B;
		PMat* T = new PMat;
		T->nrows = A.ncols; T->ncols = A.nrows; T->block = B;
		T->rp.assign(T->nrows + 1, 0);
		auto is_zero = [&](int64_t p) {
			for (int t = 0; t < BB; ++t) if (A.va[p * BB + t] != 0.0) return false;
			return true;
		};
		for (int64_t r = 0; r < A.nrows; ++r)
			for (int64_t p = A.rp[r]; p < A.rp[r + 1]; ++p)
				if (keep_zeros || !is_zero(p)) T->rp[A.ci[p] + 1]++;
		for (int64_t r = 0; r < T->nrows; ++r) T->rp[r + 1] += T->rp[r];
		T->ci.resize(T->rp[T->nrows]); T->va.resize((size_t)T->rp[T->nrows] * BB);
		std::vector<int64_t> fill(T->rp.begin(), T->rp.end() - 1);
		for (int64_t r = 0; r < A.nrows; ++r)
			for (int64_t p = A.rp[r]; p < A.rp[r + 1]; ++p) {
				if (!keep_zeros && is_zero(p)) continue;
				const int64_t q = fill[A.ci[p]]++;
				T->ci[q] = (int)r;
				for (int i = 0; i < B; ++i)
					for (int j = 0; j < B; ++j) T->va[q * BB + i + B * j] = A.va[p * BB + j + B * i];
			}
		return T;
	}

	// algebra_common/sparsematrix_util.h:152-230 (AddMultiplyOf into an empty M): for row i, for
	// R_ik != 0, for A_kl != 0: ab = R_ik * A_kl; for P_lj != 0: M_ij += ab * P_lj — accumulated in
	// that order into an unsorted sparse row (UnsortedSparseVector), then stored sorted by column.
	// R, P scalar: on the block diagonal they scale every component (zero off-diagonal terms add
	// exact zeros in the reference's block products).
	Mat* rap(const Mat& R_, const Mat& A_, const Mat& P_) override
	{
		const PMat& R = M(R_); const PMat& A = M(A_); const PMat& P = M(P_);
		if (R.block != 1 || P.block != 1) throw std::runtime_error("port backend: rap needs scalar transfers");
		if (R.ncols != A.nrows || A.ncols != P.nrows) throw std::runtime_error("port backend: rap size mismatch");
		const int B = A.block, BB = B * B;
		PMat* T = new PMat;
		T->nrows = R.nrows; T->ncols = P.ncols; T->block = B;
		T->rp.assign(T->nrows + 1, 0);
		std::vector<int> slot(P.ncols, -1), touched;
		std::vector<double> acc, ab(BB);
		for (int64_t i = 0; i < R.nrows; ++i) {
			touched.clear(); acc.clear();
			for (int64_t pr = R.rp[i]; pr < R.rp[i + 1]; ++pr) {
				const double r = R.va[pr];
				if (r == 0.0) continue;
				const int64_t k = R.ci[pr];
				for (int64_t pa = A.rp[k]; pa < A.rp[k + 1]; ++pa) {
					bool zero = true;
					for (int t = 0; t < BB; ++t) if (A.va[pa * BB + t] != 0.0) zero = false;
					if (zero) continue;
					for (int t = 0; t < BB; ++t) ab[t] = r * A.va[pa * BB + t];
					const int64_t l = A.ci[pa];
					for (int64_t pp = P.rp[l]; pp < P.rp[l + 1]; ++pp) {
						const double c = P.va[pp];
						if (c == 0.0) continue;
						const int j = P.ci[pp];
						if (slot[j] < 0) { slot[j] = (int)touched.size(); touched.push_back(j); acc.insert(acc.end(), BB, 0.0); }
						double* d = &acc[(size_t)slot[j] * BB];
						for (int t = 0; t < BB; ++t) d[t] = d[t] + ab[t] * c;
					}
				}
			}
			std::vector<std::pair<int, int> > order;
			for (size_t t = 0; t < touched.size(); ++t) order.push_back(std::make_pair(touched[t], (int)t));
			std::sort(order.begin(), order.end());
			for (size_t t = 0; t < order.size(); ++t) {
				T->ci.push_back(order[t].first);
				for (int q = 0; q < BB; ++q) T->va.push_back(acc[(size_t)order[t].second * BB + q]);
			}
			T->rp[i + 1] = (int64_t)T->ci.size();
			for (int j : touched) slot[j] = -1;
		}
		return T;
	}

	// sparsematrix_impl.h:300-316 (alpha1 == 0 branch of axpy): first connection
	// assigned, the others accumulated in ascending column order; empty row -> 0.
	void apply(const Mat& A_, Vec& y, const Vec& x) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* yd = y.data(); const double* xd = x.data();
		for (int64_t i = 0; i < A.nrows; ++i) {
			int64_t p = A.rp[i]; const int64_t e = A.rp[i + 1];
			double* d = yd + i * B;
			if (p == e) { for (int t = 0; t < B; ++t) d[t] = 0.0; continue; }
			blk_mult(B, d, 1.0, &A.va[p * BB], xd + (int64_t)A.ci[p] * B);
			for (++p; p != e; ++p) blk_mult_add(B, d, 1.0, &A.va[p * BB], xd + (int64_t)A.ci[p] * B);
		}
	}
	// sparsematrix.h:199-207 -> axpy(dest,1,dest,-1,w): sparsematrix_impl.h:318-329,
	// mat_mult_add_row :257-268 — term-by-term accumulation INTO dest[i].
	void matmul_minus(const Mat& A_, Vec& y, const Vec& x) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* yd = y.data(); const double* xd = x.data();
		for (int64_t i = 0; i < A.nrows; ++i)
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p)
				blk_mult_add(B, yd + i * B, -1.0, &A.va[p * BB], xd + (int64_t)A.ci[p] * B);
	}
	// general axpy, sparsematrix_impl.h:291-339
	void axpy(const Mat& A_, Vec& dest, double alpha, const Vec& v, double beta, const Vec& w) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* dd = dest.data(); const double* vd = v.data(); const double* wd = w.data();
		if (A.block == 1 && dest.block > 1) {
			// scalar transfer matrix acting on block vectors (prolongation): only the
			// alpha1 == 0 branch is used on the path (std_transfer_impl.h:738-740)
			if (alpha != 0.0) throw std::runtime_error("port: scalar-on-block axpy needs alpha == 0");
			const int VB = dest.block;
			for (int64_t i = 0; i < A.nrows; ++i) {
				const int64_t p = A.rp[i], e = A.rp[i + 1];
				for (int t = 0; t < VB; ++t) {
					if (p == e) { dd[i * VB + t] = 0.0; continue; }
					double s = beta * A.va[p] * wd[(int64_t)A.ci[p] * VB + t];
					for (int64_t q = p + 1; q != e; ++q) s = 1.0 * s + beta * A.va[q] * wd[(int64_t)A.ci[q] * VB + t];
					dd[i * VB + t] = s;
				}
			}
			return;
		}
		if (alpha == 0.0) {
			for (int64_t i = 0; i < A.nrows; ++i) {
				int64_t p = A.rp[i]; const int64_t e = A.rp[i + 1];
				double* d = dd + i * B;
				if (p == e) { for (int t = 0; t < B; ++t) d[t] = 0.0; continue; }
				blk_mult(B, d, beta, &A.va[p * BB], wd + (int64_t)A.ci[p] * B);
				for (++p; p != e; ++p) blk_mult_add(B, d, beta, &A.va[p * BB], wd + (int64_t)A.ci[p] * B);
			}
		} else {
			const bool inplace = (&dest == &v);
			for (int64_t i = 0; i < A.nrows; ++i) {
				double* d = dd + i * B;
				if (inplace) { if (alpha != 1.0) for (int t = 0; t < B; ++t) d[t] *= alpha; }
				else for (int t = 0; t < B; ++t) d[t] = alpha * vd[i * B + t];
				for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p)
					blk_mult_add(B, d, beta, &A.va[p * BB], wd + (int64_t)A.ci[p] * B);
			}
		}
	}
	// sparsematrix_impl.h:271-288: rows without connections leave dest untouched.
	// A.block == 1 with a block vector = scalar transfer acting per component
	// (P/R of a block algebra carry the scalar on the block diagonal).
	void apply_ignore_zero_rows(const Mat& A_, Vec& dest, double beta, const Vec& w) override
	{
		const PMat& A = M(A_);
		const int VB = dest.block;
		double* dd = dest.data(); const double* wd = w.data();
		if (A.block == 1 && VB > 1) {
			for (int64_t i = 0; i < A.nrows; ++i) {
				int64_t p = A.rp[i]; const int64_t e = A.rp[i + 1];
				if (p == e) continue;
				for (int t = 0; t < VB; ++t) {
					double s = beta * A.va[p] * wd[(int64_t)A.ci[p] * VB + t];
					for (int64_t q = p + 1; q != e; ++q) s = 1.0 * s + beta * A.va[q] * wd[(int64_t)A.ci[q] * VB + t];
					dd[i * VB + t] = s;
				}
			}
			return;
		}
		const int B = A.block, BB = B * B;
		for (int64_t i = 0; i < A.nrows; ++i) {
			int64_t p = A.rp[i]; const int64_t e = A.rp[i + 1];
			if (p == e) continue;
			double* d = dd + i * B;
			blk_mult(B, d, beta, &A.va[p * BB], wd + (int64_t)A.ci[p] * B);
			for (++p; p != e; ++p) blk_mult_add(B, d, beta, &A.va[p * BB], wd + (int64_t)A.ci[p] * B);
		}
	}

	// cpu_algebra/vector_impl.h:72-79: strictly sequential left-to-right sum of
	// VecProd(values[i], w[i]); for a block entry VecProd is itself a sequential sum
	// started at 0 (common/operations_vec.h:187-194)
	double dot(const Vec& a, const Vec& b) override
	{
		const double* x = a.data(); const double* y = b.data();
		const int B = a.block;
		double s = 0;
		for (int64_t i = 0; i < a.n; ++i) {
			double l = 0;
			for (int t = 0; t < B; ++t) l += x[i * B + t] * y[i * B + t];
			s += l;
		}
		return s;
	}
	// vector_impl.h:323-329: sqrt(sum BlockNorm2); BlockNorm2 of a DenseVector is the
	// sequential sum of squares of its components (small_algebra/blocks.h)
	double norm(const Vec& a) override
	{
		const double* x = a.data(); const int B = a.block;
		double d = 0;
		for (int64_t i = 0; i < a.n; ++i) {
			double s = 0;
			for (int t = 0; t < B; ++t) s += x[i * B + t] * x[i * B + t];
			d += s;
		}
		return std::sqrt(d);
	}
	void set(Vec& a, double v) override { std::fill(a.data(), a.data() + a.len(), v); }
	void assign(Vec& dst, const Vec& src) override { std::memcpy(dst.data(), src.data(), sizeof(double) * dst.len()); }
	void add(Vec& dst, const Vec& src) override
	{ double* d = dst.data(); const double* s = src.data(); for (int64_t i = 0, n = dst.len(); i < n; ++i) d[i] += s[i]; }
	void sub(Vec& dst, const Vec& src) override
	{ double* d = dst.data(); const double* s = src.data(); for (int64_t i = 0, n = dst.len(); i < n; ++i) d[i] -= s[i]; }
	void scale(Vec& dst, double f) override
	{ double* d = dst.data(); for (int64_t i = 0, n = dst.len(); i < n; ++i) d[i] *= f; }
	// common/operations_vec.h:55-66, 162-175
	void scale_add2(Vec& d_, double a1, const Vec& v1_, double a2, const Vec& v2_) override
	{
		double* d = d_.data(); const double* v1 = v1_.data(); const double* v2 = v2_.data();
		for (int64_t i = 0, n = d_.len(); i < n; ++i) d[i] = a1 * v1[i] + a2 * v2[i];
	}
	void scale_add3(Vec& d_, double a1, const Vec& v1_, double a2, const Vec& v2_, double a3, const Vec& v3_) override
	{
		double* d = d_.data(); const double* v1 = v1_.data(); const double* v2 = v2_.data(); const double* v3 = v3_.data();
		for (int64_t i = 0, n = d_.len(); i < n; ++i) d[i] = a1 * v1[i] + a2 * v2[i] + a3 * v3[i];
	}

	// operator/preconditioner/jacobi.h:196-220 (serial branch):
	//   m = A(i,i) (or only its diagonal when !block);  m *= 1./damp;  GetInverse(diagInv, m)
	DiagInv* jacobi_prepare(const Mat& A_, double damp, bool block) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		PDiag* D = new PDiag; D->B = B; D->inv.assign((size_t)A.nrows * BB, 0.0);
		const double s = 1. / damp;
		for (int64_t i = 0; i < A.nrows; ++i) {
			double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p)
				if (A.ci[p] == i) { std::memcpy(m, &A.va[p * BB], sizeof(double) * BB); break; }
			if (!block && B > 1)
				for (int r = 0; r < B; ++r) for (int c = 0; c < B; ++c) if (r != c) m[r + B * c] = 0.0;
			for (int t = 0; t < BB; ++t) m[t] *= s;
			blk_get_inverse(B, &D->inv[i * BB], m);
		}
		return D;
	}
	// jacobi.h:228-232: c[i] = 1.0 * diagInv[i] * d[i]
	void jacobi_step(const DiagInv& D_, Vec& c, const Vec& d) override
	{
		const PDiag& D = static_cast<const PDiag&>(D_); const int B = D.B, BB = B * B;
		double* cd = c.data(); const double* dd = d.data();
		for (int64_t i = 0; i < c.n; ++i) blk_mult(B, cd + i * B, 1.0, &D.inv[i * BB], dd + i * B);
	}

	// algebra_common/core_smoothers.h:105-130
	void gs_step_LL(const Mat& A_, Vec& c, const Vec& d, double relax) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* cd = c.data(); const double* dd = d.data();
		for (int64_t i = 0; i < A.nrows; ++i) {
			double s[3];
			for (int t = 0; t < B; ++t) s[t] = dd[i * B + t];
			int64_t p = A.rp[i]; const int64_t e = A.rp[i + 1];
			for (; p != e && A.ci[p] < i; ++p) blk_mult_add(B, s, -1.0, &A.va[p * BB], cd + (int64_t)A.ci[p] * B);
			double zero[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
			const double* Aii = (p != e && A.ci[p] == i) ? &A.va[p * BB] : zero;
			blk_inverse_mult(B, cd + i * B, relax, Aii, s);
		}
	}
	// core_smoothers.h:147-169
	void gs_step_UR(const Mat& A_, Vec& c, const Vec& d, double relax) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* cd = c.data(); const double* dd = d.data();
		if (A.nrows == 0) return;
		int64_t i = A.nrows - 1;
		do {
			double s[3];
			for (int t = 0; t < B; ++t) s[t] = dd[i * B + t];
			int64_t diag = A.rp[i]; const int64_t e = A.rp[i + 1];
			while (diag != e && A.ci[diag] != i) ++diag;
			for (int64_t p = diag + 1; p < e; ++p) blk_mult_add(B, s, -1.0, &A.va[p * BB], cd + (int64_t)A.ci[p] * B);
			blk_inverse_mult(B, cd + i * B, relax, &A.va[diag * BB], s);
		} while (i-- != 0);
	}
	// core_smoothers.h:187-206
	void sgs_step(const Mat& A_, Vec& c, const Vec& d, double relax) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		gs_step_LL(A_, c, d, relax);
		double* cd = c.data();
		for (int64_t i = 0; i < A.nrows; ++i) {
			double s[3];
			for (int t = 0; t < B; ++t) s[t] = cd[i * B + t];
			const double* Aii = nullptr;
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) if (A.ci[p] == i) { Aii = &A.va[p * BB]; break; }
			blk_mult(B, cd + i * B, 1.0, Aii, s);
		}
		gs_step_UR(A_, c, c, relax);
	}

	// operator/preconditioner/ilu.h:174-228 FactorizeILUSorted (beta == 0), :110-171 FactorizeILUBeta.
	// Scalar matrices only (the block version needs DenseMatrix::operator/= = multiplication by the inverse
	// block; the compiled reference covers blocks).
	Mat* ilu_factorize(const Mat& A_, double beta, double sortEps) override
	{
		const PMat& A0 = M(A_);
		if (A0.block != 1) throw std::runtime_error("port oracle: ILU of block matrices not restated (use the ref backend)");
		PMat* F = new PMat(A0);
		std::vector<int64_t>& rp = F->rp; std::vector<int>& ci = F->ci; std::vector<double>& va = F->va;
		const int64_t n = F->nrows;
		auto find = [&](int64_t r, int c) -> int64_t {   // get_connection(r, c): binary search in the sorted row
			const int* b = ci.data() + rp[r]; const int* e = ci.data() + rp[r + 1];
			const int* p = std::lower_bound(b, e, c);
			return (p != e && *p == c) ? (int64_t)(p - ci.data()) : -1;
		};
		if (beta != 0.0) {
			for (int64_t i = 1; i < n; ++i) {
				const int64_t dii = find(i, (int)i);
				if (dii < 0) throw std::runtime_error("ILU: row without diagonal entry");   // A(i,i) would insert
				double Nii = va[dii]; Nii *= 0.0;                                           // ilu.h:121-122
				for (int64_t pik = rp[i]; pik != rp[i + 1] && ci[pik] < i; ++pik) {
					const int k = ci[pik];
					va[pik] /= va[find(k, k)];                                              // :138
					const double a_ik = va[pik];
					for (int64_t pkj = rp[k]; pkj != rp[k + 1]; ++pkj) {                    // :142-166
						const int j = ci[pkj];
						if (j <= k) continue;
						const double a_kj = va[pkj];
						const int64_t pij = find(i, j);
						if (pij >= 0) va[pij] -= a_ik * a_kj;
						else Nii -= a_ik * a_kj;
					}
				}
				va[dii] += beta * Nii;                                                      // AddMult(Aii, beta, Nii) :170
			}
			return F;
		}
		for (int64_t i = 1; i < n; ++i) {
			for (int64_t pik = rp[i]; pik != rp[i + 1] && ci[pik] < i; ++pik) {             // :185
				const int k = ci[pik];
				const double a_kk = va[find(k, k)];
				if (std::fabs(std::fabs(a_kk)) < sortEps * std::fabs(va[pik]))                 // :194 (BlockNorm(double) = |.|)
					{ delete F; throw std::runtime_error("ILU: Blocknorm of diagonal is near-zero"); }
				va[pik] /= a_kk;                                                            // :200
				const double a_ik = va[pik];
				int64_t pij = pik + 1, pkj = rp[k];                                         // :205-222: merge of the two sorted rows
				while (pij != rp[i + 1] && pkj != rp[k + 1]) {
					if (ci[pij] > ci[pkj]) ++pkj;
					else if (ci[pij] < ci[pkj]) ++pij;
					else { va[pij] -= a_ik * va[pkj]; ++pkj; ++pij; }
				}
			}
		}
		return F;
	}
	// ilu.h:233-252
	void ilu_invert_L(const Mat& LU_, Vec& x, const Vec& b) override
	{
		const PMat& A = M(LU_);
		if (A.block != 1) throw std::runtime_error("port oracle: ILU of block matrices not restated");
		double* xd = x.data(); const double* bd = b.data();
		for (int64_t i = 0; i < A.nrows; ++i) {
			double s = bd[i];
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) {
				if (A.ci[p] >= i) continue;
				s = 1.0 * s + (-1.0 * A.va[p]) * xd[A.ci[p]];        // MatMultAdd(s, 1.0, s, -1.0, a, x[j])
			}
			xd[i] = s;
		}
	}
	// ilu.h:257-322
	bool ilu_invert_U(const Mat& LU_, Vec& x, const Vec& b, double eps) override
	{
		const PMat& A = M(LU_);
		if (A.block != 1) throw std::runtime_error("port oracle: ILU of block matrices not restated");
		double* xd = x.data(); const double* bd = b.data();
		const int64_t n = A.nrows;
		bool result = true;
		auto diag = [&](int64_t i) -> double {
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) if (A.ci[p] == i) return A.va[p];
			return 0.0;
		};
		if (n > 0) {
			const int64_t i = n - 1;
			const double s = bd[i];
			if (std::fabs(diag(i)) <= eps * std::fabs(s)) { xd[i] = 0; result = false; }    // :285-297
			else xd[i] = 1.0 * s / diag(i);                                                  // InverseMatMult(x[i], 1.0, A(i,i), s)
		}
		if (n <= 1) return result;
		for (int64_t i = n - 2;; --i) {
			double s = bd[i];
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) {
				if (A.ci[p] <= i) continue;
				s = 1.0 * s + (-1.0 * A.va[p]) * xd[A.ci[p]];
			}
			xd[i] = 1.0 * s / diag(i);
			if (i == 0) break;
		}
		return result;
	}
	// ordering_strategies/algorithms/native_cuthill_mckee.cpp:100-300 on the graph GetCuthillMcKeeOrder builds
	// (algebra_common/permutation_util.h:96-114: every stored column of a row, the diagonal included)
	void cuthill_mckee(const Mat& A_, bool reverse, bool preserveConsec, std::vector<size_t>& vNewIndex) override
	{
		const PMat& A = M(A_);
		const size_t nDoF = (size_t)A.nrows;
		std::vector<std::vector<size_t> > con(nDoF);
		for (size_t i = 0; i < nDoF; ++i) for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) con[i].push_back((size_t)A.ci[p]);
		auto lessDeg = [&](size_t i, size_t j) { return con[i].size() < con[j].size(); };
		std::vector<bool> handled(nDoF, false);
		for (size_t i = 0; i < nDoF; ++i) {
			if (con[i].empty()) handled[i] = true;
			else std::stable_sort(con[i].begin(), con[i].end(), lessDeg);                    // :121-133
		}
		std::vector<size_t> sorting(nDoF);
		for (size_t i = 0; i < nDoF; ++i) sorting[i] = i;
		std::stable_sort(sorting.begin(), sorting.end(), lessDeg);                           // :138-143
		std::vector<size_t> order;
		size_t firstNonHandled = 0;
		while (true) {
			size_t k = firstNonHandled;
			for (; k < nDoF; ++k) if (!handled[sorting[k]]) { firstNonHandled = k; break; }
			if (k == nDoF) break;
			std::vector<size_t> queue(1, sorting[firstNonHandled]);                          // breadth first, :172-199
			for (size_t head = 0; head < queue.size(); ++head) {
				const size_t front = queue[head];
				if (handled[front]) continue;
				order.push_back(front); handled[front] = true;
				for (size_t t = 0; t < con[front].size(); ++t) if (!handled[con[front][t]]) queue.push_back(con[front][t]);
			}
		}
		vNewIndex.assign(nDoF, (size_t)-1);
		if (preserveConsec) {                                                                // :205-262
			size_t cnt = 0;
			for (size_t newInd = 0; newInd < nDoF; ++newInd) {
				if (con[newInd].empty()) continue;
				const size_t oldInd = reverse ? order[order.size() - 1 - cnt] : order[cnt];
				++cnt;
				vNewIndex[oldInd] = newInd;
			}
			if (cnt != order.size()) throw std::runtime_error("OrderCuthillMcKee: Not all indices sorted that must be sorted");
			// findBlockSize (:68-95)
			auto gcd = [](size_t a, size_t b) { while (b) { size_t r = a % b; a = b; b = r; } return a; };
			size_t blockSize;
			{
				size_t cd = 0;
				while (cd < nDoF && con[cd].empty()) ++cd;
				if (cd == nDoF) blockSize = nDoF;
				else {
					size_t bs = 1;
					for (size_t i = cd + 1; i < nDoF; ++i) {
						if (con[i].empty()) { ++bs; continue; }
						cd = gcd(bs, cd);
						bs = 1;
					}
					blockSize = gcd(bs, cd);
				}
			}
			for (size_t i = 0; i < nDoF; i += blockSize) if (vNewIndex[i] == (size_t)-1) vNewIndex[i] = i;
			for (size_t i = 0; i < nDoF; i += blockSize) for (size_t j = 1; j < blockSize; ++j) vNewIndex[i + j] = vNewIndex[i] + j;
		} else {                                                                             // :264-291
			size_t sz = order.size();
			for (size_t i = 0; i < sz; ++i) vNewIndex[reverse ? order[sz - 1 - i] : order[i]] = i;
			size_t next = sz;
			for (size_t i = 0; i < nDoF; ++i) if (vNewIndex[i] == (size_t)-1) vNewIndex[i] = next++;
		}
	}

	double maxnorm(const Vec& a) override
	{ double d = 0; const double* v = a.data(); for (int64_t i = 0; i < a.len(); ++i) d = std::max(d, std::fabs(v[i])); return d; }
	// vector_impl.h:91-96 with urand of common/math/misc/math_util_impl.hpp:64-74
	void set_random(Vec& a, double from, double to) override
	{
		double* v = a.data();
		for (int64_t i = 0; i < a.len(); ++i) { long t = std::rand(); if (t == RAND_MAX) t -= 1; v[i] = from + (double)((to - from) * ((double)t / (double)RAND_MAX)); }
	}
	// sparsematrix_impl.h:341-370 with alpha1 = 0, beta1 = 1: dest = 0; for every stored, non-zero a_ij in row order:
	// dest[j] = 1.0 * dest[j] + (1.0 * a_ij^T) * w[i]   (MatMultTransposedAdd)
	void apply_transposed(const Mat& A_, Vec& y, const Vec& x) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		double* yd = y.data(); const double* xd = x.data();
		for (int64_t i = 0; i < y.len(); ++i) yd[i] = 0.0;
		for (int64_t i = 0; i < A.nrows; ++i)
			for (int64_t p = A.rp[i]; p != A.rp[i + 1]; ++p) {
				bool zero = true;
				for (int t = 0; t < BB; ++t) if (A.va[p * BB + t] != 0.0) { zero = false; break; }
				if (zero) continue;
				double* d = yd + (int64_t)A.ci[p] * B;
				for (int r = 0; r < B; ++r) {
					d[r] = 1.0 * d[r];
					for (int c = 0; c < B; ++c) d[r] = 1.0 * d[r] + 1.0 * A.va[p * BB + c + B * r] * xd[i * B + c];
				}
			}
	}
	Mat* matrix_script(int64_t, const double*, std::vector<unsigned char>&) override
	{ throw std::runtime_error("port oracle: the assembly-side matrix API is not restated (use the ref backend)"); }

	// operator/linear_solver/lu.h:122-140 (init_dense) with the non-LAPACK kernels
	// small_algebra/no_lapack/lu_decomp.h:45-75 (LUDecomp with row interchange)
	DenseLU* lu_init(const Mat& A_) override
	{
		const PMat& A = M(A_); const int B = A.block, BB = B * B;
		const int64_t n = A.nrows * B;
		PLU* L = new PLU; L->n = n; L->a.assign((size_t)n * n, 0.0); L->piv.assign(n, 0);
		std::vector<double>& a = L->a;
#define AA(i, j) a[(size_t)(i) * n + (j)]
		for (int64_t r = 0; r < A.nrows; ++r)
			for (int64_t p = A.rp[r]; p != A.rp[r + 1]; ++p)
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j)
					AA(r * B + i, (int64_t)A.ci[p] * B + j) = A.va[p * BB + i + B * j];
		for (int64_t k = 0; k < n; ++k) {
			int64_t biggest = k;
			for (int64_t j = k + 1; j < n; ++j) if (std::fabs(AA(biggest, k)) < std::fabs(AA(j, k))) biggest = j;
			if (biggest != k) for (int64_t j = 0; j < n; ++j) std::swap(AA(k, j), AA(biggest, j));
			L->piv[k] = (size_t)biggest;
			if (std::fabs(AA(k, k)) < 1e-10) { delete L; return nullptr; }
			for (int64_t i = k + 1; i < n; ++i) {
				AA(i, k) = AA(i, k) / AA(k, k);
				for (int64_t j = k + 1; j < n; ++j) AA(i, j) = AA(i, j) - AA(i, k) * AA(k, j);
			}
		}
		return L;
	}
	// lu.h:189-207 (solve_dense) + lu_decomp.h:160-195 (SolveLU)
	void lu_apply(const DenseLU& L_, Vec& x_, const Vec& b) override
	{
		const PLU& L = static_cast<const PLU&>(L_); const int64_t n = L.n; const std::vector<double>& a = L.a;
		if (&x_ != &b) assign(x_, b);
		double* x = x_.data();
		for (int64_t i = 0; i < n; ++i) if ((size_t)i < L.piv[i]) std::swap(x[i], x[L.piv[i]]);
		for (int64_t i = 0; i < n; ++i) {
			double s = x[i];
			for (int64_t k = 0; k < i; ++k) s -= AA(i, k) * x[k];
			x[i] = s;
		}
		for (int64_t i = n - 1;; --i) {
			double s = x[i];
			for (int64_t k = i + 1; k < n; ++k) s -= AA(i, k) * x[k];
			x[i] = s / AA(i, i);
			if (i == 0) break;
		}
#undef AA
	}
};

} // namespace

Backend* make_port_backend() { return new PortBackend; }

} // namespace oracle
