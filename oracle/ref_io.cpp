/*
 * oracle/ref_io.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference's OWN ConnectionViewer writer and reader, compiled from /root/reference/ugbase
 * where they lie (lib_algebra/common/connection_viewer_output.h:84-145, 400-425,
 * connection_viewer_input.h:48-165), so that tests/test_matrix_io.py can check the product's
 * import / export (ugcore_b200/csrc/host/matrix_io.h) against files the reference itself
 * writes and reads.  (MatrixIOMtx needs boost::lexical_cast / boost::algorithm, which are
 * absent: the MatrixMarket side is checked against the format rules and scipy.io instead.)
 * Part of oracle/_ref/liboracle_ref.so only.
 */
#ifdef ORACLE_WITH_UGREF
#include "lib_algebra/cpu_algebra/sparsematrix.h"
#include "lib_algebra/cpu_algebra/vector.h"
#include "common/math/ugmath.h"
#include "lib_algebra/common/connection_viewer_output.h"
#include "lib_algebra/common/connection_viewer_input.h"
#include <cstdint>
#include <string>
#include <vector>

using namespace ug;

namespace {
std::vector<MathVector<3> > make_pos(const double* pos, int64_t n)
{
	std::vector<MathVector<3> > p((size_t)(n > 0 ? n : 1));
	for (int64_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) p[i][d] = pos ? pos[3 * i + d] : 0.0;
	return p;
}
void fill(SparseMatrix<double>& A, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals)
{
	A.resize_and_clear((size_t)nrows, (size_t)ncols);
	for (int64_t r = 0; r < nrows; ++r)
		for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) A((size_t)r, (size_t)cols[p]) = vals[p];
	A.defragment();
}
struct ReadResult { SparseMatrix<double> A; std::vector<MathVector<3> > grid; int dim = 0; };
}

extern "C" {

int oracle_ref_cv_write_matrix(const char* fn, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
                               const double* vals, const double* pos, int dim)
{
	SparseMatrix<double> A; fill(A, nrows, ncols, rowptr, cols, vals);
	std::vector<MathVector<3> > p = make_pos(pos, nrows);
	ConnectionViewer::WriteMatrix(std::string(fn), A, &p[0], dim);
	return 0;
}
/* from / to form: pos holds nrows "to" positions followed by ncols "from" positions */
int oracle_ref_cv_write_matrix_from_to(const char* fn, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols,
                                       const double* vals, const double* pos, int dim)
{
	SparseMatrix<double> A; fill(A, nrows, ncols, rowptr, cols, vals);
	std::vector<MathVector<3> > to = make_pos(pos, nrows), from = make_pos(pos ? pos + 3 * nrows : nullptr, ncols);
	to.resize((size_t)nrows); from.resize((size_t)ncols);
	return ConnectionViewer::WriteMatrix(std::string(fn), A, from, to, (size_t)dim) ? 0 : 1;
}
int oracle_ref_cv_write_vector(const char* fn, int64_t n, const double* v, const double* pos, int dim)
{
	Vector<double> b((size_t)n);
	for (int64_t i = 0; i < n; ++i) b[i] = v[i];
	std::vector<MathVector<3> > p = make_pos(pos, n);
	ConnectionViewer::WriteVector(std::string(fn), b, &p[0], dim);
	return 0;
}
void* oracle_ref_cv_read_matrix(const char* fn)
{
	ReadResult* r = new ReadResult;
	if (!ConnectionViewer::ReadMatrix(std::string(fn), r->A, r->grid, r->dim)) { delete r; return nullptr; }
	return r;
}
void oracle_ref_cv_read_info(void* h, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* dim)
{
	ReadResult* r = (ReadResult*)h;
	*nrows = (int64_t)r->A.num_rows(); *ncols = (int64_t)r->A.num_cols(); *nnz = (int64_t)r->A.total_num_connections(); *dim = r->dim;
}
void oracle_ref_cv_read_export(void* h, int64_t* rowptr, int* cols, double* vals, double* pos)
{
	ReadResult* r = (ReadResult*)h;
	const SparseMatrix<double>& A = r->A;
	int64_t p = 0;
	for (size_t i = 0; i < A.num_rows(); ++i) {
		rowptr[i] = p;
		for (SparseMatrix<double>::const_row_iterator it = A.begin_row(i); it != A.end_row(i); ++it) { cols[p] = (int)it.index(); vals[p] = it.value(); ++p; }
	}
	rowptr[r->A.num_rows()] = p;
	for (size_t i = 0; i < r->grid.size(); ++i) for (int d = 0; d < 3; ++d) pos[3 * i + d] = r->grid[i][d];
}
void oracle_ref_cv_read_free(void* h) { delete (ReadResult*)h; }
int oracle_ref_cv_read_vector(const char* fn, int64_t n, double* out)
{
	Vector<double> v;
	if (!ConnectionViewer::ReadVector(std::string(fn), v)) return 1;
	if ((int64_t)v.size() != n) return 2;
	for (int64_t i = 0; i < n; ++i) out[i] = v[i];
	return 0;
}

} // extern "C"
#endif
