/*
 * oracle/ref_backend.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Adapter that runs the REAL ugcore templates (SparseMatrix<T>, Vector<T>,
 * gs_step_LL/UR, sgs_step, block inverses, LUDecomp/SolveLU) compiled straight
 * from /root/reference/ugbase where they lie — the same trick the reference's own
 * tests use (tests/sm_transpose.cc:1-10).  No reference source is copied into this
 * repository; this file only calls the reference API.  Built by oracle/Makefile
 * into oracle/_ref/ when /root/reference is present.
 */
#include "backend.h"

#include "lib_algebra/cpu_algebra/sparsematrix.h"
#include "lib_algebra/cpu_algebra/sparsematrix_impl.h"
#include "lib_algebra/cpu_algebra/vector.h"
#include "lib_algebra/small_algebra/small_algebra.h"
#include "lib_algebra/small_algebra/additional_math.h"
#include "lib_algebra/algebra_common/core_smoothers.h"
#include "lib_algebra/algebra_common/sparsematrix_util.h"
#include "lib_algebra/common/operations_vec.h"
// the real ILU kernels and Cuthill-McKee (made compilable without boost by ref_prelude.h)
#include "lib_algebra/operator/preconditioner/ilu.h"
#include "lib_algebra/algebra_common/permutation_util.h"

#include <stdexcept>
#include <map>

// common/util/crc32.cpp needs boost::crc (absent); only used for debug-id hashing.
namespace ug {
uint32 crc32(const char* s)
{
	uint32 h = 2166136261u;
	while (*s) { h ^= (unsigned char)*s++; h *= 16777619u; }
	return h;
}
}

namespace oracle {
namespace {

template <int B> struct T {
	typedef ug::DenseMatrix<ug::FixedArray2<double, B, B> > mval;
	typedef ug::DenseVector<ug::FixedArray1<double, B> > vval;
	static double& m(mval& a, int r, int c) { return a(r, c); }
	static double mc(const mval& a, int r, int c) { return a(r, c); }
};
template <> struct T<1> {
	typedef double mval;
	typedef double vval;
	static double& m(mval& a, int, int) { return a; }
	static double mc(const mval& a, int, int) { return a; }
};

struct RMatBase : Mat {
	// expanded copies (scalar on the block diagonal) used when a scalar transfer
	// matrix acts on block vectors — exactly how P/R look in a CPUBlockAlgebra<N>
	mutable std::map<int, std::unique_ptr<Mat> > expanded;
};
template <int B> struct RMat : RMatBase { ug::SparseMatrix<typename T<B>::mval> A; };

template <int B> struct RVec : Vec {
	ug::Vector<typename T<B>::vval> v;
	static_assert(sizeof(typename T<B>::vval) == sizeof(double) * B, "block vector must be a plain array");
	double* data() override { return v.size() ? reinterpret_cast<double*>(&v[0]) : nullptr; }
	const double* data() const override { return v.size() ? reinterpret_cast<const double*>(&v[0]) : nullptr; }
};
template <int B> struct RDiag : DiagInv {
	std::vector<typename ug::block_traits<typename T<B>::mval>::inverse_type> inv;
};
struct RLU : DenseLU {
	ug::DenseMatrixInverse<ug::DenseMatrix<ug::VariableArray2<double> > > mat;
	ug::DenseVector<ug::VariableArray1<double> > tmp;
	size_t size = 0;
};

template <int B> const ug::SparseMatrix<typename T<B>::mval>& SM(const Mat& m) { return static_cast<const RMat<B>&>(m).A; }
template <int B> ug::Vector<typename T<B>::vval>& V(Vec& v) { return static_cast<RVec<B>&>(v).v; }
template <int B> const ug::Vector<typename T<B>::vval>& V(const Vec& v) { return static_cast<const RVec<B>&>(v).v; }

#define DISPATCH(blk, CALL)                                          \
	switch (blk) {                                                   \
		case 1: { const int BB_ = 1; (void)BB_; CALL(1); break; }    \
		case 2: { CALL(2); break; }                                  \
		case 3: { CALL(3); break; }                                  \
		default: throw std::runtime_error("ref backend: block size must be 1, 2 or 3"); \
	}

template <int B> Mat* build(int64_t nr, int64_t nc, const int64_t* rp, const int* ci, const double* va)
{
	typedef ug::SparseMatrix<typename T<B>::mval> SMt;
	RMat<B>* m = new RMat<B>;
	m->nrows = nr; m->ncols = nc; m->block = B;
	m->A.resize_and_clear(nr, nc);
	std::vector<typename SMt::connection> row;
	for (int64_t r = 0; r < nr; ++r) {
		row.resize(rp[r + 1] - rp[r]);
		for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
			typename SMt::connection& c = row[p - rp[r]];
			c.iIndex = ci[p];
			for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) T<B>::m(c.dValue, i, j) = va[p * B * B + i + B * j];
		}
		if (!row.empty()) m->A.set_matrix_row(r, &row[0], row.size());
	}
	m->A.defragment();
	return m;
}

template <int B> Mat* expand(const Mat& S_)
{
	const ug::SparseMatrix<double>& S = SM<1>(S_);
	typedef ug::SparseMatrix<typename T<B>::mval> SMt;
	RMat<B>* m = new RMat<B>;
	m->nrows = S_.nrows; m->ncols = S_.ncols; m->block = B;
	m->A.resize_and_clear(S_.nrows, S_.ncols);
	std::vector<typename SMt::connection> row;
	for (size_t r = 0; r < S.num_rows(); ++r) {
		row.clear();
		for (ug::SparseMatrix<double>::const_row_iterator it = S.begin_row(r); it != S.end_row(r); ++it) {
			typename SMt::connection c;
			c.iIndex = it.index();
			c.dValue = 0.0;
			for (int i = 0; i < B; ++i) T<B>::m(c.dValue, i, i) = it.value();
			row.push_back(c);
		}
		if (!row.empty()) m->A.set_matrix_row(r, &row[0], row.size());
	}
	m->A.defragment();
	return m;
}

const Mat& for_vec(const Mat& A, int vblock)
{
	if (A.block == vblock) return A;
	if (A.block != 1) throw std::runtime_error("ref backend: matrix/vector block mismatch");
	const RMatBase& b = static_cast<const RMatBase&>(A);
	std::unique_ptr<Mat>& e = b.expanded[vblock];
	if (!e) {
		if (vblock == 2) e.reset(expand<2>(A));
		else if (vblock == 3) e.reset(expand<3>(A));
		else throw std::runtime_error("ref backend: block size must be 1, 2 or 3");
	}
	return *e;
}

struct RefBackend : Backend {
	const char* name() const override { return "ref"; }

	Mat* matrix(int block, int64_t nr, int64_t nc, const int64_t* rp, const int* ci, const double* va) override
	{
		Mat* m = nullptr;
#define CALL(B) m = build<B>(nr, nc, rp, ci, va)
		DISPATCH(block, CALL)
#undef CALL
		return m;
	}
	Vec* vector(int64_t n, int block) override
	{
		Vec* v = nullptr;
#define CALL(B) { RVec<B>* r = new RVec<B>; r->v.resize(n); r->v.set(0.0); v = r; }
		DISPATCH(block, CALL)
#undef CALL
		v->n = n; v->block = block;
		return v;
	}
	int64_t nnz(const Mat& A) override
	{
		int64_t n = 0;
#define CALL(B) n = (int64_t)SM<B>(A).total_num_connections()
		DISPATCH(A.block, CALL)
#undef CALL
		return n;
	}
	template <int B> void exp_(const Mat& A_, int64_t* rp, int* ci, double* va)
	{
		const ug::SparseMatrix<typename T<B>::mval>& A = SM<B>(A_);
		int64_t p = 0; rp[0] = 0;
		for (size_t r = 0; r < A.num_rows(); ++r) {
			for (typename ug::SparseMatrix<typename T<B>::mval>::const_row_iterator it = A.begin_row(r); it != A.end_row(r); ++it, ++p) {
				ci[p] = (int)it.index();
				for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) va[p * B * B + i + B * j] = T<B>::mc(it.value(), i, j);
			}
			rp[r + 1] = p;
		}
	}
	void export_crs(const Mat& A, int64_t* rp, int* ci, double* va) override
	{
#define CALL(B) exp_<B>(A, rp, ci, va)
		DISPATCH(A.block, CALL)
#undef CALL
	}
	Mat* transpose(const Mat& A, bool keep_zeros) override
	{
		Mat* t = nullptr;
#define CALL(B) { RMat<B>* m = new RMat<B>; m->nrows = A.ncols; m->ncols = A.nrows; m->block = B; \
		if (keep_zeros) m->A.set_as_transpose_of(SM<B>(A)); else m->A.set_as_transpose_of2(SM<B>(A)); t = m; }
		DISPATCH(A.block, CALL)
#undef CALL
		return t;
	}

	Mat* rap(const Mat& R0, const Mat& A, const Mat& P0) override
	{
		const Mat& R = for_vec(R0, A.block); const Mat& P = for_vec(P0, A.block);
		Mat* t = nullptr;
#define CALL(B) { RMat<B>* m = new RMat<B>; m->nrows = R.nrows; m->ncols = P.ncols; m->block = B; \
		m->A.resize_and_clear(R.nrows, P.ncols); ug::AddMultiplyOf(m->A, SM<B>(R), SM<B>(A), SM<B>(P)); m->A.defragment(); t = m; }
		DISPATCH(A.block, CALL)
#undef CALL
		return t;
	}

	void apply(const Mat& A0, Vec& y, const Vec& x) override
	{
		const Mat& A = for_vec(A0, y.block);
#define CALL(B) SM<B>(A).apply(V<B>(y), V<B>(x))
		DISPATCH(A.block, CALL)
#undef CALL
	}
	void matmul_minus(const Mat& A0, Vec& y, const Vec& x) override
	{
		const Mat& A = for_vec(A0, y.block);
#define CALL(B) SM<B>(A).matmul_minus(V<B>(y), V<B>(x))
		DISPATCH(A.block, CALL)
#undef CALL
	}
	double maxnorm(const Vec& a) override
	{
		double r = 0;
#define CALL(B) r = V<B>(a).maxnorm()
		DISPATCH(a.block, CALL)
#undef CALL
		return r;
	}
	void set_random(Vec& a, double from, double to) override
	{
#define CALL(B) V<B>(a).set_random(from, to)
		DISPATCH(a.block, CALL)
#undef CALL
	}
	void apply_transposed(const Mat& A, Vec& y, const Vec& x) override
	{
#define CALL(B) SM<B>(A).apply_transposed(V<B>(y), V<B>(x))
		DISPATCH(A.block, CALL)
#undef CALL
	}
	void axpy(const Mat& A0, Vec& dest, double alpha, const Vec& v, double beta, const Vec& w) override
	{
		const Mat& A = for_vec(A0, dest.block);
#define CALL(B) SM<B>(A).axpy(V<B>(dest), alpha, V<B>(v), beta, V<B>(w))
		DISPATCH(A.block, CALL)
#undef CALL
	}
	void apply_ignore_zero_rows(const Mat& A0, Vec& dest, double beta, const Vec& w) override
	{
		const Mat& A = for_vec(A0, dest.block);
#define CALL(B) SM<B>(A).apply_ignore_zero_rows(V<B>(dest), beta, V<B>(w))
		DISPATCH(A.block, CALL)
#undef CALL
	}

	double dot(const Vec& a, const Vec& b) override
	{
		double r = 0;
#define CALL(B) r = const_cast<ug::Vector<typename T<B>::vval>&>(V<B>(a)).dotprod(V<B>(b))
		DISPATCH(a.block, CALL)
#undef CALL
		return r;
	}
	double norm(const Vec& a) override
	{
		double r = 0;
#define CALL(B) r = V<B>(a).norm()
		DISPATCH(a.block, CALL)
#undef CALL
		return r;
	}
	void set(Vec& a, double v) override
	{
#define CALL(B) V<B>(a).set(v)
		DISPATCH(a.block, CALL)
#undef CALL
	}
	void assign(Vec& dst, const Vec& src) override
	{
#define CALL(B) V<B>(dst) = V<B>(src)
		DISPATCH(dst.block, CALL)
#undef CALL
	}
	void add(Vec& dst, const Vec& src) override
	{
#define CALL(B) V<B>(dst) += V<B>(src)
		DISPATCH(dst.block, CALL)
#undef CALL
	}
	void sub(Vec& dst, const Vec& src) override
	{
#define CALL(B) V<B>(dst) -= V<B>(src)
		DISPATCH(dst.block, CALL)
#undef CALL
	}
	void scale(Vec& dst, double s) override
	{
#define CALL(B) V<B>(dst) *= s
		DISPATCH(dst.block, CALL)
#undef CALL
	}
	void scale_add2(Vec& d, double a1, const Vec& v1, double a2, const Vec& v2) override
	{
#define CALL(B) ug::VecScaleAdd(V<B>(d), a1, V<B>(v1), a2, V<B>(v2))
		DISPATCH(d.block, CALL)
#undef CALL
	}
	void scale_add3(Vec& d, double a1, const Vec& v1, double a2, const Vec& v2, double a3, const Vec& v3) override
	{
#define CALL(B) ug::VecScaleAdd(V<B>(d), a1, V<B>(v1), a2, V<B>(v2), a3, V<B>(v3))
		DISPATCH(d.block, CALL)
#undef CALL
	}

	// jacobi.h:196-220, serial branch, executed with the reference block operations
	template <int B> DiagInv* jprep(const Mat& A_, double damp, bool block)
	{
		ug::SparseMatrix<typename T<B>::mval>& mat = const_cast<ug::SparseMatrix<typename T<B>::mval>&>(SM<B>(A_));
		RDiag<B>* D = new RDiag<B>;
		D->inv.resize(mat.num_rows());
		typename T<B>::mval m;
		for (size_t i = 0; i < mat.num_rows(); ++i) {
			typename T<B>::mval& d = mat(i, i);
			if (!block) ug::GetDiag(m, d); else m = d;
			m *= 1. / damp;
			ug::GetInverse(D->inv[i], m);
		}
		return D;
	}
	DiagInv* jacobi_prepare(const Mat& A, double damp, bool block) override
	{
		DiagInv* d = nullptr;
#define CALL(B) d = jprep<B>(A, damp, block)
		DISPATCH(A.block, CALL)
#undef CALL
		return d;
	}
	template <int B> void jstep(const DiagInv& D_, Vec& c_, const Vec& d_)
	{
		const RDiag<B>& D = static_cast<const RDiag<B>&>(D_);
		ug::Vector<typename T<B>::vval>& c = V<B>(c_); const ug::Vector<typename T<B>::vval>& d = V<B>(d_);
		for (size_t i = 0; i < D.inv.size(); ++i) ug::MatMult(c[i], 1.0, D.inv[i], d[i]); // jacobi.h:228-232
	}
	void jacobi_step(const DiagInv& D, Vec& c, const Vec& d) override
	{
#define CALL(B) jstep<B>(D, c, d)
		DISPATCH(c.block, CALL)
#undef CALL
	}
	void gs_step_LL(const Mat& A, Vec& c, const Vec& d, double relax) override
	{
#define CALL(B) ug::gs_step_LL(SM<B>(A), V<B>(c), V<B>(d), relax)
		DISPATCH(A.block, CALL)
#undef CALL
	}
	void gs_step_UR(const Mat& A, Vec& c, const Vec& d, double relax) override
	{
#define CALL(B) ug::gs_step_UR(SM<B>(A), V<B>(c), V<B>(d), relax)
		DISPATCH(A.block, CALL)
#undef CALL
	}
	void sgs_step(const Mat& A, Vec& c, const Vec& d, double relax) override
	{
#define CALL(B) ug::sgs_step(SM<B>(A), V<B>(c), V<B>(d), relax)
		DISPATCH(A.block, CALL)
#undef CALL
	}

	// ilu.h:563-576: m_ILU = mat; FactorizeILUBeta / FactorizeILUSorted (SparseMatrix::rows_sorted is true); defragment
	Mat* ilu_factorize(const Mat& A, double beta, double sortEps) override
	{
		Mat* t = nullptr;
#define CALL(B) { RMat<B>* m = new RMat<B>; m->nrows = A.nrows; m->ncols = A.ncols; m->block = B; m->A = SM<B>(A); \
		if (beta != 0.0) ug::FactorizeILUBeta(m->A, beta); else ug::FactorizeILUSorted(m->A, sortEps); m->A.defragment(); t = m; }
		DISPATCH(A.block, CALL)
#undef CALL
		return t;
	}
	void ilu_invert_L(const Mat& LU, Vec& x, const Vec& b) override
	{
#define CALL(B) ug::invert_L(SM<B>(LU), V<B>(x), V<B>(b))
		DISPATCH(LU.block, CALL)
#undef CALL
	}
	bool ilu_invert_U(const Mat& LU, Vec& x, const Vec& b, double eps) override
	{
		bool ok = true;
#define CALL(B) ok = ug::invert_U(SM<B>(LU), V<B>(x), V<B>(b), eps)
		DISPATCH(LU.block, CALL)
#undef CALL
		return ok;
	}
	void cuthill_mckee(const Mat& A, bool reverse, bool preserveConsec, std::vector<size_t>& newIndex) override
	{
#define CALL(B) ug::GetCuthillMcKeeOrder(SM<B>(A), newIndex, reverse, preserveConsec)
		DISPATCH(A.block, CALL)
#undef CALL
	}

	Mat* matrix_script(int64_t nops, const double* ops, std::vector<unsigned char>& isolated) override
	{
		typedef ug::SparseMatrix<double> SM1;
		std::unique_ptr<RMat<1> > m(new RMat<1>);
		for (int64_t k = 0; k < nops; ++k) {
			const int code = (int)ops[4 * k]; const size_t r = (size_t)ops[4 * k + 1], c = (size_t)ops[4 * k + 2]; const double v = ops[4 * k + 3];
			SM1& A = m->A;
			switch (code) {
				case 0: A.resize_and_clear(r, c); break;
				case 1: A(r, c) = v; break;
				case 2: A(r, c) += v; break;
				case 3: A.scale(v); break;
				case 4: A.clear_retain_structure(); break;
				case 5: A.resize_and_keep_values(r, c); break;
				case 6: A.defragment(); break;
				case 7: A.set(v); break;
				case 8: case 9: {
					std::unique_ptr<RMat<1> > t(new RMat<1>);
					if (code == 8) t->A.set_as_transpose_of(A, v); else t->A.set_as_copy_of(A, v);
					m.swap(t);
					break;
				}
				case 10: { const SM1& cA = A; volatile double sink = cA(r, c); (void)sink; break; }
				case 11: case 13: {
					std::vector<SM1::connection> row(c);
					for (size_t t = 0; t < c; ++t) { row[t].iIndex = (size_t)ops[4 * (k + 1 + t) + 2]; row[t].dValue = ops[4 * (k + 1 + t) + 3]; }
					if (code == 11) A.set_matrix_row(r, row.data(), c); else A.add_matrix_row(r, row.data(), c);
					k += (int64_t)c;
					break;
				}
				default: throw std::runtime_error("matrix script: unknown operation");
			}
		}
		m->nrows = (int64_t)m->A.num_rows(); m->ncols = (int64_t)m->A.num_cols(); m->block = 1;
		isolated.resize(m->A.num_rows());
		for (size_t i = 0; i < m->A.num_rows(); ++i) isolated[i] = m->A.is_isolated(i) ? 1 : 0;
		return m.release();
	}

	// lu.h:122-140 init_dense / :189-207 solve_dense with the reference's dense kernels
	template <int B> DenseLU* luinit(const Mat& A_)
	{
		RLU* L = new RLU;
		L->size = ug::GetDenseDoubleFromSparse(L->mat, SM<B>(A_));
		if (L->mat.invert() == false) { delete L; return nullptr; }
		return L;
	}
	DenseLU* lu_init(const Mat& A) override
	{
		DenseLU* l = nullptr;
#define CALL(B) l = luinit<B>(A)
		DISPATCH(A.block, CALL)
#undef CALL
		return l;
	}
	void lu_apply(const DenseLU& L_, Vec& x, const Vec& b) override
	{
		RLU& L = const_cast<RLU&>(static_cast<const RLU&>(L_));
		if (&x != &b) assign(x, b);
		L.tmp.resize(L.size);
		const double* bd = b.data(); double* xd = x.data();
		for (size_t k = 0; k < L.size; ++k) L.tmp[k] = bd[k];
		L.mat.apply(L.tmp);
		for (size_t k = 0; k < L.size; ++k) xd[k] = L.tmp[k];
	}
};

} // namespace

Backend* make_ref_backend() { return new RefBackend; }

} // namespace oracle
