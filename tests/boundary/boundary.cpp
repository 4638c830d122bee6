// tests/boundary/boundary.cpp — the GPU algebra compiled against ugcore's REAL headers (test code).
//
// Built with  g++ -DUG4B200_WITH_UGCORE -I/root/reference/ugbase -I ugcore_b200/csrc -I include  (tests/boundary/Makefile):
// csrc/host/ug_base.h and operators.h then pull ugcore's own common.h, smart_pointer.h, small_algebra.h,
// algebra_type.h, linear_operator.h, matrix_operator.h, linear_iterator.h, preconditioner.h and debug_writer.h
// instead of the stand-ins of the stand-alone build.  What this translation unit pins:
//
//   * struct GPUAlgebra / GPUBlockAlgebra<3> have the shape cpu_algebra_types.h:101-120 sketches and sit next to
//     CPUAlgebra (cpu_algebra_types.h:76-92): matrix_type, vector_type, blockSize, get_type() == AlgebraType(GPU, n);
//   * ug::MatrixOperator<GPUSparseMatrix<double>, GPUVector<double>> — ugcore's own template
//     (matrix_operator.h:46-75) — instantiates over the GPU types and is an ug::ILinearOperator;
//   * Jacobi / GaussSeidel / SymmetricGaussSeidel / ILU <GPUAlgebra> derive from ugcore's
//     ug::IPreconditioner<GPUAlgebra> (preconditioner.h:99-388), the GMG from ug::ILinearIterator (linear_iterator.h:79-198);
//   * the assembly-side API of GPUSparseMatrix takes the same calls as ugcore's SparseMatrix<double> (both are filled
//     by ONE template here) and y = A x through ug::ILinearOperator<..>* is bit-identical between the two algebras;
//   * a GMG-preconditioned CG solve driven only through ug::ILinearOperator / ug::ILinearIterator base-class pointers.
//
//   boundary_test types            host only (no device call): prints the checks above as JSON
//   boundary_test solve <refs>     needs a B200: SpMV parity CPUAlgebra vs GPUAlgebra + the solve; JSON with the history
#include "lib_algebra/cpu_algebra_types.h"                       // CPUAlgebra: SparseMatrix<double>, Vector<double>
#include "host/multigrid.h"                                      // the GPU algebra, in its UG4B200_WITH_UGCORE configuration
#include "synth/synth.h"
#include <cstdio>
#include <cstring>
#include <type_traits>

#if !defined(UG4B200_WITH_UGCORE)
#error "tests/boundary is the with-ugcore configuration"
#endif
// the include guards of the REAL interface headers: proof that no stand-in is in play
#if !defined(__H__LIB_ALGEBRA__OPERATOR__INTERFACE__MATRIX_OPERATOR__) || !defined(__H__LIB_ALGEBRA__OPERATOR__INTERFACE__PRECONDITIONER__) || \
    !defined(__H__LIB_ALGEBRA__OPERATOR__INTERFACE__OPERATOR_ITERATOR__) || !defined(__H__LIB_ALGEBRA__OPERATOR__INTERFACE__LINEAR_OPERATOR__)
#error "ugcore's interface headers were not included"
#endif

using namespace ug;

typedef GPUAlgebra::vector_type gvec_t;
typedef GPUAlgebra::matrix_type gmat_t;
typedef MatrixOperator<gmat_t, gvec_t> gop_t;                  // ugcore's MatrixOperator over the GPU types
typedef CPUAlgebra::vector_type cvec_t;
typedef CPUAlgebra::matrix_type cmat_t;
typedef MatrixOperator<cmat_t, cvec_t> cop_t;

// ---- compile-time shape of the boundary ----
static_assert(std::is_same<GPUAlgebra::matrix_type, GPUSparseMatrix<double> >::value, "cpu_algebra_types.h:108");
static_assert(std::is_same<GPUAlgebra::vector_type, GPUVector<double> >::value, "cpu_algebra_types.h:109");
static_assert(GPUAlgebra::blockSize == 1 && GPUBlockAlgebra<3>::blockSize == 3, "blockSize");
static_assert(std::is_base_of<ILinearOperator<gvec_t>, gop_t>::value, "MatrixOperator is an ILinearOperator");
static_assert(std::is_base_of<gmat_t, gop_t>::value, "MatrixOperator inherits the matrix (matrix_operator.h:46-48)");
static_assert(std::is_base_of<IPreconditioner<GPUAlgebra>, Jacobi<GPUAlgebra> >::value, "Jacobi : ug::IPreconditioner");
static_assert(std::is_base_of<IPreconditioner<GPUAlgebra>, GaussSeidel<GPUAlgebra> >::value, "GaussSeidel : ug::IPreconditioner");
static_assert(std::is_base_of<IPreconditioner<GPUAlgebra>, SymmetricGaussSeidel<GPUAlgebra> >::value, "SGS : ug::IPreconditioner");
static_assert(std::is_base_of<IPreconditioner<GPUAlgebra>, ILU<GPUAlgebra> >::value, "ILU : ug::IPreconditioner");
static_assert(std::is_base_of<IPreconditioner<GPUBlockAlgebra<3> >, Jacobi<GPUBlockAlgebra<3> > >::value, "block Jacobi");
static_assert(std::is_base_of<ILinearIterator<gvec_t>, AssembledMultiGridCycle<GPUAlgebra> >::value, "GMG : ug::ILinearIterator");
static_assert(std::is_base_of<DebugWritingObject<GPUAlgebra>, Jacobi<GPUAlgebra> >::value, "ugcore's IPreconditioner brings DebugWritingObject");

namespace {

/// ONE assembly routine for both algebras: the calls DomainDiscretization makes on matrix_type
/// (sparsematrix.h:116-343: resize_and_clear, operator()(r, c) inserting)
template <typename TMatrix>
void assemble(TMatrix& A, const synth_crs& c)
{
	A.resize_and_clear((size_t)c.nrows, (size_t)c.ncols);
	for (int64_t r = 0; r < c.nrows; ++r)
		for (int64_t p = c.rowptr[r]; p < c.rowptr[r + 1]; ++p) A((size_t)r, (size_t)c.cols[p]) = c.vals[p];
	A.defragment();
}

void fail(const char* what) { std::printf("{\"ok\": false, \"error\": \"%s\"}\n", what); std::exit(1); }

int run_types()
{
	// host only: nothing below touches the device
	SmartPtr<gop_t> A = make_sp<gop_t>();
	A->resize_and_clear(3, 3);
	(*A)(0, 0) = 2.0; (*A)(0, 1) = -1.0; (*A)(1, 0) = -1.0; (*A)(1, 1) = 2.0; (*A)(1, 2) = -1.0; (*A)(2, 1) = -1.0; (*A)(2, 2) = 2.0;
	SmartPtr<ILinearOperator<gvec_t> > L = A;                                       // derived -> base through ugcore's SmartPtr
	SmartPtr<gop_t> back = L.cast_dynamic<gop_t>();                                 // what IPreconditioner::init does (preconditioner.h:196-203)
	if (back.invalid() || back->num_rows() != 3 || back->total_num_connections() != 7) fail("cast_dynamic / host matrix API");
	SmartPtr<ILinearIterator<gvec_t> > it = make_sp<Jacobi<GPUAlgebra> >(0.66);     // the GPU smoother behind ugcore's iterator interface
	SmartPtr<ILinearIterator<gvec_t> > cl = it->clone();
	if (cl.invalid() || std::strcmp(it->name(), "Jacobi") != 0) fail("ILinearIterator::clone / name");
	if (it->damping()->damping() != 0.66 || !it->damping()->constant_damping()) fail("ILinearIterator::set_damp (linear_iterator.h:170-187)");
	const AlgebraType t1 = GPUAlgebra::get_type(), t3 = GPUBlockAlgebra<3>::get_type();
	if (t1.type() != AlgebraType::GPU || t1.blocksize() != 1 || t3.blocksize() != 3) fail("AlgebraType");
	bool threw = false;
	try { UG_THROW("probe " << 42); } catch (UGError& e) { threw = (e.get_msg().find("probe 42") != std::string::npos); }
	if (!threw) fail("UG_THROW does not raise ugcore's UGError");
	std::printf("{\"ok\": true, \"with_ugcore\": true, \"matrix_operator\": \"ug::MatrixOperator<GPUSparseMatrix<double>, GPUVector<double> >\", "
	            "\"algebra_type\": [%d, %d], \"block_algebra_type\": [%d, %d]}\n", t1.type(), t1.blocksize(), t3.type(), t3.blocksize());
	return 0;
}

int run_solve(int refs)
{
	synth_desc d; std::memset(&d, 0, sizeof(d));
	d.dim = 3; d.base[0] = d.base[1] = d.base[2] = 1; d.num_refs = refs; d.problem = SYNTH_POISSON; d.eps = 1.0;
	d.part[0] = d.part[1] = d.part[2] = 1; d.E = 1.0; d.nu = 0.3;
	synth_problem* P = nullptr;
	if (synth_create(&d, &P) != 0) fail(synth_last_error());
	synth_crs c;
	const double* rhs = nullptr; int64_t n = 0;
	synth_rhs(P, &rhs, &n);

	// ---- 1. the same assembled operator in both algebras; y = A x through ILinearOperator* ----
	synth_level_matrix(P, refs, &c);
	SmartPtr<cop_t> Ac = make_sp<cop_t>();
	SmartPtr<gop_t> Ag = make_sp<gop_t>();
	assemble(static_cast<cmat_t&>(*Ac), c);
	assemble(static_cast<gmat_t&>(*Ag), c);
	cvec_t xc((size_t)n), yc((size_t)n);
	gvec_t xg((size_t)n), yg((size_t)n);
	for (int64_t i = 0; i < n; ++i) { const double v = std::sin(0.37 * (double)i) + 1e-3 * (double)(i % 17); xc[(size_t)i] = v; xg[(size_t)i] = v; }   // host access: operator[]
	ILinearOperator<cvec_t>* Lc = Ac.get();
	ILinearOperator<gvec_t>* Lg = Ag.get();
	Lc->apply(yc, xc);
	Lg->apply(yg, xg);
	bool spmv_equal = true;
	for (int64_t i = 0; i < n; ++i) if (std::memcmp(&yc[(size_t)i], &yg[(size_t)i], sizeof(double)) != 0) { spmv_equal = false; break; }
	for (int64_t i = 0; i < n; ++i) { yc[(size_t)i] = rhs[i]; yg[(size_t)i] = rhs[i]; }
	Lc->apply_sub(yc, xc);                                                       // matmul_minus
	Lg->apply_sub(yg, xg);
	bool sub_equal = true;
	for (int64_t i = 0; i < n; ++i) if (std::memcmp(&yc[(size_t)i], &yg[(size_t)i], sizeof(double)) != 0) { sub_equal = false; break; }

	// ---- 2. GMG V(2,2) Jacobi(0.66) + CG, wired like util.solver does, driven through base-class pointers ----
	SmartPtr<AssembledMultiGridCycle<GPUAlgebra> > gmg = make_sp<AssembledMultiGridCycle<GPUAlgebra> >();
	gmg->set_base_level(0); gmg->set_surface_level(refs); gmg->set_cycle_type("V");
	gmg->set_num_presmooth(2); gmg->set_num_postsmooth(2);
	SmartPtr<ILinearIterator<gvec_t> > smoother = make_sp<Jacobi<GPUAlgebra> >(0.66);
	gmg->set_smoother(smoother);
	gmg->set_base_solver(make_sp<LU<GPUAlgebra> >());
	for (int l = 0; l <= refs; ++l) {
		if (l < refs) {
			synth_level_matrix(P, l, &c);
			SmartPtr<gop_t> Al = make_sp<gop_t>();
			assemble(static_cast<gmat_t&>(*Al), c);
			gmg->set_level_operator(l, Al);
		}
		if (l > 0) {
			SmartPtr<GPUTransferMatrix> Pm = make_sp<GPUTransferMatrix>(), Rm = make_sp<GPUTransferMatrix>();
			synth_prolongation(P, l, &c); assemble(*Pm, c);
			synth_restriction(P, l, &c); assemble(*Rm, c);
			gmg->set_level_transfer(l, Pm, Rm);
		}
	}
	SmartPtr<ILinearIterator<gvec_t> > precond = gmg;                             // ug::ILinearIterator from here on
	SmartPtr<ILinearOperator<gvec_t> > J = Ag;                                     // ug::ILinearOperator from here on
	SmartPtr<CG<gvec_t> > cg = make_sp<CG<gvec_t> >();
	SmartPtr<StdConvCheck<gvec_t> > cc = make_sp<StdConvCheck<gvec_t> >(100, 1e-12, 1e-10, false);
	cg->set_convergence_check(cc);
	cg->set_preconditioner(precond);
	gvec_t u((size_t)n), b((size_t)n);
	for (int64_t i = 0; i < n; ++i) b[(size_t)i] = rhs[i];
	u.set(0.0);
	if (!cg->init(J, u)) fail("solver:init(J, u)");
	const bool ok = cg->apply(u, b);
	// one more cycle directly through the iterator interface: c = B d, d -= A c (ILinearIterator::apply_update_defect)
	gvec_t dd((size_t)n), corr((size_t)n);
	for (int64_t i = 0; i < n; ++i) dd[(size_t)i] = rhs[i];
	const double d0 = dd.norm();
	ILinearIterator<gvec_t>* itp = precond.get();
	if (!itp->apply_update_defect(corr, dd)) fail("ILinearIterator::apply_update_defect");
	const double d1 = dd.norm();

	std::printf("{\"ok\": %s, \"n\": %lld, \"spmv_bit_exact\": %s, \"matmul_minus_bit_exact\": %s, \"converged\": %s, \"steps\": %d, "
	            "\"cycle_reduction\": %.17g, \"history\": [", (spmv_equal && sub_equal && ok) ? "true" : "false", (long long)n,
	            spmv_equal ? "true" : "false", sub_equal ? "true" : "false", ok ? "true" : "false", cc->step(), d1 / d0);
	const std::vector<number>& h = cc->get_defects();
	for (size_t i = 0; i < h.size(); ++i) std::printf("%s%.17g", i ? ", " : "", h[i]);
	std::printf("], \"solution_norm\": %.17g}\n", u.norm());
	synth_destroy(P);
	GPUManager::finalize();
	return (spmv_equal && sub_equal && ok) ? 0 : 1;
}

} // namespace

int main(int argc, char** argv)
{
	try {
		if (argc >= 2 && std::strcmp(argv[1], "types") == 0) return run_types();
		if (argc >= 2 && std::strcmp(argv[1], "solve") == 0) return run_solve(argc >= 3 ? std::atoi(argv[2]) : 3);
		std::fprintf(stderr, "usage: boundary_test types | solve <refs>\n");
		return 2;
	} catch (UGError& e) {
		std::printf("{\"ok\": false, \"error\": \"UGError: %s\"}\n", e.get_msg().c_str());
		return 1;
	} catch (std::exception& e) {
		std::printf("{\"ok\": false, \"error\": \"%s\"}\n", e.what());
		return 1;
	}
}
