/*
 * oracle/backend.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Kernel-level interface of the CPU oracle.  Two implementations exist:
 *   PortBackend (port_backend.cpp)  — plain C++ restatement of ugcore's
 *       SparseMatrix / Vector / core_smoothers arithmetic, no reference headers;
 *   RefBackend  (ref_backend.cpp)   — thin adapter around the REAL ugcore
 *       templates, compiled from /root/reference/ugbase where it lies
 *       (only built when that tree is present; output in oracle/_ref/).
 * The solver control flow (solvers.cpp) is written once against this interface.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may use anything under oracle/.
 */
#ifndef ORACLE_BACKEND_H
#define ORACLE_BACKEND_H
#include <cstdint>
#include <memory>
#include <vector>

namespace oracle {

struct Mat {
	virtual ~Mat() {}
	int64_t nrows = 0, ncols = 0;
	int block = 1;
};

struct Vec {
	virtual ~Vec() {}
	virtual double* data() = 0;
	virtual const double* data() const = 0;
	int64_t n = 0; // number of blocks
	int block = 1;
	int64_t len() const { return n * block; }
};

struct DiagInv { virtual ~DiagInv() {} };
struct DenseLU { virtual ~DenseLU() {} };

struct Backend {
	virtual ~Backend() {}
	virtual const char* name() const = 0;

	virtual Mat* matrix(int block, int64_t nrows, int64_t ncols, const int64_t* rowptr,
	                    const int* cols, const double* vals) = 0;
	virtual Vec* vector(int64_t nblocks, int block) = 0;
	virtual int64_t nnz(const Mat& A) = 0;
	virtual void export_crs(const Mat& A, int64_t* rowptr, int* cols, double* vals) = 0;
	// set_as_transpose_of (keeps explicit zeros) / set_as_transpose_of2 (drops them)
	virtual Mat* transpose(const Mat& A, bool keep_zeros) = 0;
	// Galerkin product M = R * A * P by AddMultiplyOf (algebra_common/sparsematrix_util.h:152-230), as
	// AssembledMultiGridCycle::init_rap_operator builds coarse operators (mg_solver_impl.hpp:959);
	// R, P scalar (expanded to the block diagonal for block A, as ugcore stores them)
	virtual Mat* rap(const Mat& R, const Mat& A, const Mat& P) = 0;

	virtual void apply(const Mat& A, Vec& y, const Vec& x) = 0;        // y = A x
	virtual void matmul_minus(const Mat& A, Vec& y, const Vec& x) = 0; // y -= A x
	virtual void axpy(const Mat& A, Vec& dest, double alpha, const Vec& v, double beta, const Vec& w) = 0;
	virtual void apply_ignore_zero_rows(const Mat& A, Vec& dest, double beta, const Vec& w) = 0;
	virtual void apply_transposed(const Mat& A, Vec& y, const Vec& x) = 0;  // y = A^T x, sparsematrix_impl.h:341-370

	virtual double dot(const Vec& a, const Vec& b) = 0;
	virtual double norm(const Vec& a) = 0;
	virtual double maxnorm(const Vec& a) = 0;                         // vector_impl.h:332-338
	virtual void set_random(Vec& a, double from, double to) = 0;      // vector_impl.h:91-96 (C rand(), seeded by the caller)
	virtual void set(Vec& a, double v) = 0;
	virtual void assign(Vec& dst, const Vec& src) = 0;
	virtual void add(Vec& dst, const Vec& src) = 0;   // dst += src
	virtual void sub(Vec& dst, const Vec& src) = 0;   // dst -= src
	virtual void scale(Vec& dst, double s) = 0;       // dst *= s
	virtual void scale_add2(Vec& d, double a1, const Vec& v1, double a2, const Vec& v2) = 0;
	virtual void scale_add3(Vec& d, double a1, const Vec& v1, double a2, const Vec& v2, double a3, const Vec& v3) = 0;

	virtual DiagInv* jacobi_prepare(const Mat& A, double damp, bool block) = 0;
	virtual void jacobi_step(const DiagInv& D, Vec& c, const Vec& d) = 0;
	virtual void gs_step_LL(const Mat& A, Vec& c, const Vec& d, double relax) = 0;
	virtual void gs_step_UR(const Mat& A, Vec& c, const Vec& d, double relax) = 0;
	virtual void sgs_step(const Mat& A, Vec& c, const Vec& d, double relax) = 0;

	// ILU(0) / ILU(beta) (lib_algebra/operator/preconditioner/ilu.h): factorisation of a copy of A with
	// FactorizeILUSorted (:174-228; beta == 0) or FactorizeILUBeta (:110-171), L below the diagonal with unit
	// diagonal implied, U on and above it; invert_L (:233-252), invert_U (:257-322, false = the last row's
	// near-zero check fired)
	virtual Mat* ilu_factorize(const Mat& A, double beta, double sortEps) = 0;
	virtual void ilu_invert_L(const Mat& LU, Vec& x, const Vec& b) = 0;
	virtual bool ilu_invert_U(const Mat& LU, Vec& x, const Vec& b, double eps) = 0;
	// Cuthill-McKee order of the matrix graph: GetCuthillMcKeeOrder (algebra_common/permutation_util.h:96-114) ->
	// ComputeCuthillMcKeeOrder (ordering_strategies/algorithms/native_cuthill_mckee.cpp:100-300); newIndex[old] = new
	virtual void cuthill_mckee(const Mat& A, bool reverse, bool preserveConsec, std::vector<size_t>& newIndex) = 0;

	// the assembly-side API of SparseMatrix (cpu_algebra/sparsematrix.h:116-343) driven by a small op script
	// (codes as in include/ug4b200_solver.h: ug4b200_host_matrix_script); isolated[i] = is_isolated(i).
	// Only the compiled reference implements it (the port restates the solve path, not the assembly side).
	virtual Mat* matrix_script(int64_t nops, const double* ops, std::vector<unsigned char>& isolated) = 0;

	virtual DenseLU* lu_init(const Mat& A) = 0;                  // nullptr if singular
	virtual void lu_apply(const DenseLU& lu, Vec& x, const Vec& b) = 0;
};

Backend* make_port_backend();
#ifdef ORACLE_WITH_UGREF
Backend* make_ref_backend();
#endif

} // namespace oracle
#endif
