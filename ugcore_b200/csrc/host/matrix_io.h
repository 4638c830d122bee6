// matrix_io.h — ConnectionViewer (.mat / .vec) and MatrixMarket (.mtx) import / export for the
// GPU algebra's host-side matrices and vectors.
//
// Why it sits next to the hot path: these are the two formats a ugcore installation can dump its
// assembled objects in — the surface matrix, every level matrix, P, R, the right-hand side and
// the per-iteration residuals (debug writers: ugbase/lib_algebra/operator/linear_solver/cg.h:274-280,
// ugbase/lib_disc/operator/linear_operator/multi_grid_solver/mg_solver_impl.hpp:692-696, 2181-2200) —
// so a real UG4 assembly can be solved here and compared, without boost / plugins in this tree.
//
// Reference (paths relative to /root/reference/ugbase/lib_algebra/common):
//   connection_viewer_output.h:84-111   WriteGridHeader   version 1 / dimension / #positions / positions / "1"
//   connection_viewer_output.h:121-145  WriteMatrix       "row col value" per connection, explicit zeros as " 0",
//                                                          values in the stream's default format (6 significant digits!)
//   connection_viewer_output.h:325-385  WriteMatrix (from / to positions): rows first, columns offset by #rows
//   connection_viewer_output.h:400-425  WriteVector       "i i value" with digits10 + 1 = 16 significant digits
//   connection_viewer_input.h:48-110    ReadMatrix        stops at a line starting with 'c' / 'v' (markers), drops zeros
//   connection_viewer_input.h:112-165   ReadVector
//   matrixio/matrix_io_mtx.h:224-256    MatrixIOMtx::read_into   coordinate real, 1-based, symmetric / skew expansion
//   matrixio/matrix_io_mtx.h:286-325    MatrixIOMtx::write_from  column-major, symmetry detection (:397-493),
//   matrixio/matrix_io_mtx.cpp:267-301  banner; entries "m n  v" in scientific notation with 13 digits
// Files written here with the reference's settings are byte-identical to the reference's writers
// (tests/test_matrix_io.py compares against the compiled reference writer); `precision` > 0 asks for
// lossless output instead (17 significant digits) — ugcore's default .mat precision of 6 digits is NOT
// enough to reproduce a residual history, so dumps meant for parity must be written with a raised
// stream precision or as vectors / MatrixMarket files.
#pragma once
#include "gpu_sparsematrix.h"
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

namespace ug {

/// one position per index, as the reference's MathVector<dim> arrays: pos[i][0..dim-1]
struct IOPositions {
	int dim = 3;
	std::vector<double> xyz;   // 3 doubles per position, unused components 0
	size_t size() const { return xyz.size() / 3; }
	const double* operator[](size_t i) const { return &xyz[3 * i]; }
	void resize(size_t n) { xyz.assign(3 * n, 0.0); }
};

namespace ConnectionViewer {

/// connection_viewer_output.h:84-111
inline bool WriteGridHeader(std::ostream& f, const IOPositions& positions, size_t N, int dimension)
{
	f << 1 << "\n";
	f << dimension << "\n";
	f << N << "\n";
	if (dimension == 1)
		for (size_t i = 0; i < N; i++) f << positions[i][0] << " 0.0\n";
	else if (dimension == 2)
		for (size_t i = 0; i < N; i++) f << positions[i][0] << " " << positions[i][1] << "\n";
	else
		for (size_t i = 0; i < N; i++) f << positions[i][0] << " " << positions[i][1] << " " << positions[i][2] << "\n";
	f << 1 << "\n"; // stringsInWindow
	return true;
}

/// connection_viewer_output.h:121-145.  precision 0: the reference's behaviour (stream default, 6 digits)
template <typename Matrix_type>
void WriteMatrix(std::ostream& file, const Matrix_type& A, const IOPositions& positions, int dimensions, int precision = 0)
{
	const size_t rows = A.num_rows();
	if (positions.size() < rows) UG_THROW("ConnectionViewer::WriteMatrix: one position per row needed");
	if (precision > 0) file << std::setprecision(precision);
	WriteGridHeader(file, positions, rows, dimensions);
	const std::vector<int64_t>& rp = A.crs_rowptr(); const std::vector<int>& ci = A.crs_cols(); const std::vector<double>& va = A.crs_vals();
	for (size_t i = 0; i < rows; i++)
		for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
			if (va[p] != 0.0) file << i << " " << ci[p] << " " << va[p] << std::endl;
			else file << i << " " << ci[p] << " 0" << std::endl;
		}
}
template <typename Matrix_type>
void WriteMatrix(const std::string& filename, const Matrix_type& A, const IOPositions& positions, int dimensions, int precision = 0)
{
	std::fstream file(filename.c_str(), std::ios::out);
	if (!file.is_open()) UG_THROW("ConnectionViewer::WriteMatrix: cannot open " << filename);
	WriteMatrix(file, A, positions, dimensions, precision);
}

/// connection_viewer_output.h:325-385: rectangular matrices (prolongation / restriction): the positions of
/// the rows come first, columns are numbered behind them
template <typename Matrix_type>
bool WriteMatrix(const std::string& filename, const Matrix_type& A, const IOPositions& positionsFrom, const IOPositions& positionsTo,
                 size_t dimensions, int precision = 0)
{
	if (positionsFrom.size() != A.num_cols() || positionsTo.size() != A.num_rows()) return false;
	const size_t fromOffset = positionsTo.size();
	std::fstream file(filename.c_str(), std::ios::out);
	if (!file.is_open()) UG_THROW("ConnectionViewer::WriteMatrix: cannot open " << filename);
	if (precision > 0) file << std::setprecision(precision);
	file << 1 << std::endl;
	file << dimensions << std::endl;
	file << positionsFrom.size() + positionsTo.size() << std::endl;
	for (int pass = 0; pass < 2; ++pass) {
		const IOPositions& P = pass == 0 ? positionsTo : positionsFrom;
		for (size_t i = 0; i < P.size(); i++) {
			if (dimensions == 1) file << P[i][0] << " 0.0" << std::endl;
			else if (dimensions == 2) file << P[i][0] << " " << P[i][1] << std::endl;
			else file << P[i][0] << " " << P[i][1] << " " << P[i][2] << std::endl;
		}
	}
	file << 1 << std::endl; // show all cons
	const std::vector<int64_t>& rp = A.crs_rowptr(); const std::vector<int>& ci = A.crs_cols(); const std::vector<double>& va = A.crs_vals();
	for (size_t i = 0; i < A.num_rows(); i++)
		for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
			if (va[p] != 0.0) file << i << " " << ci[p] + fromOffset << " " << va[p] << std::endl;
			else file << i << " " << ci[p] + fromOffset << " 0" << std::endl;
		}
	return true;
}

/// connection_viewer_output.h:400-425 (values always with 16 significant digits; precision > 16 raises it)
inline void WriteVector(const std::string& filename, const double* b, size_t rows, const IOPositions& positions, int dimensions,
                        int precision = 0)
{
	std::fstream file(filename.c_str(), std::ios::out);
	if (!file.is_open()) UG_THROW("ConnectionViewer::WriteVector: cannot open " << filename);
	if (positions.size() < rows) UG_THROW("ConnectionViewer::WriteVector: one position per entry needed");
	WriteGridHeader(file, positions, rows, dimensions);
	const int prec = precision > std::numeric_limits<double>::digits10 + 1 ? precision : std::numeric_limits<double>::digits10 + 1;
	for (size_t i = 0; i < rows; i++) file << i << " " << i << " " << std::setprecision(prec) << b[i] << std::endl;
}

struct Header { int version = -1, dimension = -1, gridsize = -1; };

inline Header read_header(std::istream& f, IOPositions& grid, const std::string& filename)
{
	Header h;
	f >> h.version >> h.dimension >> h.gridsize;
	if (!f || h.version != 1) UG_THROW("ConnectionViewer: " << filename << " is not a version-1 ConnectionViewer file");
	if (h.dimension < 1 || h.dimension > 3 || h.gridsize < 0) UG_THROW("ConnectionViewer: bad dimension / size in " << filename);
	grid.dim = h.dimension; grid.resize((size_t)h.gridsize);
	for (int i = 0; i < h.gridsize; i++) {
		double x = 0, y = 0, z = 0;
		f >> x >> y;                       // 1-d files carry "x 0.0"
		if (h.dimension == 3) f >> z;
		if (!f) UG_THROW("ConnectionViewer: " << filename << " ends inside the position block");
		grid.xyz[3 * (size_t)i] = x; grid.xyz[3 * (size_t)i + 1] = y; grid.xyz[3 * (size_t)i + 2] = z;
	}
	int printStringsInWindow = 0;
	f >> printStringsInWindow;
	return h;
}

/// connection_viewer_input.h:48-110.  keepZeros = false reproduces the reference (zero values are not
/// inserted); true keeps them as explicit zeros (the pattern UG4 retains on Dirichlet rows,
/// algebra_common/sparsematrix_util.h:850-861).  nTo > 0: the file was written with from / to positions
/// (rows 0..nTo-1, columns numbered behind them): the result is nTo x (gridsize - nTo).
template <typename Matrix_type>
bool ReadMatrix(const std::string& filename, Matrix_type& matrix, IOPositions& grid, int& dimension, bool keepZeros = false, size_t nTo = 0)
{
	std::fstream matfile(filename.c_str(), std::ios::in);
	if (!matfile.is_open()) return false;
	const Header h = read_header(matfile, grid, filename);
	dimension = h.dimension;
	const size_t n = (size_t)h.gridsize;
	if (nTo > n) UG_THROW("ConnectionViewer::ReadMatrix: nTo exceeds the number of positions");
	const size_t rows = nTo ? nTo : n, cols = nTo ? n - nTo : n, off = nTo ? nTo : 0;
	matrix.resize_and_clear(rows, cols);
	while (true) {
		matfile >> std::ws;
		const int c = matfile.peek();
		if (c == EOF || c == 'c' || c == 'v' || matfile.eof()) break;
		long long from, to; double value;
		matfile >> from >> to >> value;
		if (!matfile) UG_THROW("ConnectionViewer::ReadMatrix: malformed connection line in " << filename);
		if (from < 0 || (size_t)from >= rows || to < (long long)off || (size_t)to >= off + cols)
			UG_THROW("ConnectionViewer::ReadMatrix: connection (" << from << ", " << to << ") outside the matrix in " << filename);
		if (value != 0.0 || keepZeros) matrix((size_t)from, (size_t)to - off) = value;
	}
	matrix.defragment();
	return true;
}

/// connection_viewer_input.h:112-165
inline bool ReadVector(const std::string& filename, std::vector<double>& vec, IOPositions& grid, int& dimension)
{
	std::fstream matfile(filename.c_str(), std::ios::in);
	if (!matfile.is_open()) return false;
	const Header h = read_header(matfile, grid, filename);
	dimension = h.dimension;
	vec.assign((size_t)h.gridsize, 0.0);
	while (true) {
		matfile >> std::ws;
		const int c = matfile.peek();
		if (c == EOF || c == 'c' || c == 'v' || matfile.eof()) break;
		long long from, to; double value;
		matfile >> from >> to >> value;
		if (!matfile) UG_THROW("ConnectionViewer::ReadVector: malformed line in " << filename);
		if (from != to || from < 0 || from >= h.gridsize) UG_THROW("ConnectionViewer::ReadVector: bad index in " << filename);
		vec[(size_t)from] = value;
	}
	return true;
}

} // namespace ConnectionViewer

/// MatrixMarket exchange files, coordinate real general / symmetric / skew-symmetric
/// (matrixio/matrix_io_mtx.{h,cpp}; same member names)
class MatrixIOMtx {
  public:
	enum AlgebraicType { GENERAL = 0, SYMMETRIC = 1, SKEW = 2 };
	explicit MatrixIOMtx(const std::string& mFile) : m_file(mFile) {}

	/// matrix_io_mtx.cpp:141-236: banner + size line
	void query()
	{
		std::ifstream f(m_file.c_str());
		if (!f.is_open()) UG_THROW("MatrixIOMtx: cannot open " << m_file);
		std::string line;
		std::getline(f, line);
		std::stringstream first(line);
		std::vector<std::string> it; std::string w;
		while (first >> w) it.push_back(w);
		if (it.size() < 5 || it[0] != "%%MatrixMarket" || it[1] != "matrix") UG_THROW("MatrixIOMtx: " << m_file << " is not a valid Matrix Market file");
		if (lower(it[2]) != "coordinate") UG_THROW("Other than sparse MatrixMarket matrices are not yet implemented.");
		const std::string num = lower(it[3]);
		if (num != "real" && num != "integer") UG_THROW("MatrixIOMtx: numeric type '" << it[3] << "' not supported (real only)");
		const std::string alg = lower(it[4]);
		if (alg == "general") m_type = GENERAL; else if (alg == "symmetric") m_type = SYMMETRIC; else if (alg == "skew-symmetric") m_type = SKEW;
		else UG_THROW("MatrixIOMtx: algebraic type '" << it[4] << "' not supported");
		m_firstDataLine = 1;
		do {
			if (!std::getline(f, line)) UG_THROW("MatrixIOMtx: unexpected end of file in " << m_file);
			m_firstDataLine++;
		} while (line.empty() || line[0] == '%');
		std::stringstream dims(line);
		long long r = 0, c = 0, l = 0;
		dims >> r >> c >> l;
		if (!dims || r <= 0 || c <= 0 || l < 0) UG_THROW("MatrixIOMtx: bad size line in " << m_file);
		m_rows = (size_t)r; m_cols = (size_t)c; m_lines = (size_t)l;
		m_queried = true;
	}
	size_t get_num_rows() { if (!m_queried) query(); return m_rows; }
	size_t get_num_cols() { if (!m_queried) query(); return m_cols; }
	size_t get_num_lines() { if (!m_queried) query(); return m_lines; }
	AlgebraicType algebraic_type() { if (!m_queried) query(); return m_type; }

	/// matrix_io_mtx.h:224-256
	template <typename matrix_type>
	void read_into(matrix_type& matrix)
	{
		if (!m_queried) query();
		std::ifstream f(m_file.c_str());
		std::string line;
		for (size_t i = 0; i < m_firstDataLine; i++) std::getline(f, line);
		matrix.resize_and_clear(m_rows, m_cols);
		for (size_t i = 0; i < m_lines; i++) {
			if (!std::getline(f, line)) UG_THROW("MatrixIOMtx: " << m_file << " has fewer data lines than announced");
			std::stringstream ss(line);
			long long x = 0, y = 0; double val = 0.0;
			ss >> x >> y >> val;
			if (!ss) UG_THROW("Sparse matrix requires three values per line. Found: '" << line << "'");
			if (x < 1 || (size_t)x > m_rows || y < 1 || (size_t)y > m_cols) UG_THROW("MatrixIOMtx: entry (" << x << ", " << y << ") outside the matrix");
			matrix((size_t)x - 1, (size_t)y - 1) = val;           // MM is 1-based
			if (m_type == SYMMETRIC && x != y) matrix((size_t)y - 1, (size_t)x - 1) = val;
			else if (m_type == SKEW && x != y) matrix((size_t)y - 1, (size_t)x - 1) = -val;
		}
		matrix.defragment();
	}

	/// matrix_io_mtx.h:286-325 + determine_matrix_characteristics (:397-493): non-zero entries only,
	/// column-major, lower triangle for (skew-)symmetric matrices, "%.13e" values
	template <typename matrix_type>
	void write_from(const matrix_type& matrix, std::string comment = "%Generated with ug4.")
	{
		const size_t rows = matrix.num_rows(), cols = matrix.num_cols();
		const std::vector<int64_t>& rp = matrix.crs_rowptr(); const std::vector<int>& ci = matrix.crs_cols(); const std::vector<double>& va = matrix.crs_vals();
		auto entry = [&](size_t r, size_t c) -> double {   // matrix(r, c), 0 if not stored
			if (r >= rows) return 0.0;
			const int* b = ci.data() + rp[r]; const int* e = ci.data() + rp[r + 1];
			const int* it = std::lower_bound(b, e, (int)c);
			return (it != e && (size_t)*it == c) ? va[it - ci.data()] : 0.0;
		};
		bool isSymmetric = true, isSkew = true;
		for (size_t r = 0; r < rows && (isSymmetric || isSkew); r++)
			for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
				if (va[p] == 0.0 || (size_t)ci[p] == r) continue;
				const double t = entry((size_t)ci[p], r);
				if (va[p] != t) isSymmetric = false;
				if (va[p] != -1.0 * t) isSkew = false;
			}
		if (isSymmetric) isSkew = false;   // a diagonal matrix is written as symmetric
		std::vector<std::vector<size_t> > rowIndexPerCol(cols);
		size_t diagEntries = 0, offDiagEntries = 0;
		for (size_t r = 0; r < rows; r++)
			for (int64_t p = rp[r]; p < rp[r + 1]; ++p) {
				if (va[p] == 0.0) continue;
				const size_t c = (size_t)ci[p];
				if (!(isSymmetric || isSkew) || c <= r) rowIndexPerCol[c].push_back(r);
				(c == r) ? diagEntries++ : offDiagEntries++;
			}
		m_type = isSymmetric ? SYMMETRIC : (isSkew ? SKEW : GENERAL);
		m_rows = rows; m_cols = cols;
		m_lines = (m_type == GENERAL) ? offDiagEntries + diagEntries : offDiagEntries / 2 + diagEntries;
		m_queried = true;
		std::ofstream f(m_file.c_str(), std::ios_base::out | std::ios_base::trunc);
		if (!f.is_open()) UG_THROW("MatrixIOMtx: cannot open " << m_file << " for writing");
		f << "%%MatrixMarket matrix coordinate real " << (m_type == GENERAL ? "general" : (m_type == SYMMETRIC ? "symmetric" : "skew-symmetric")) << "\n";
		if (!comment.empty()) {
			if (comment.find_first_of('%') != 0) comment.insert(0, "%");
			f << comment << "\n";
		}
		f << m_rows << " " << m_cols << " " << m_lines << "\n";
		for (size_t col = 0; col < cols; col++)
			for (size_t k = 0; k < rowIndexPerCol[col].size(); k++) {
				const size_t row = rowIndexPerCol[col][k];
				const double val = entry(row, col);
				f.unsetf(std::ios_base::scientific);
				f << row + 1 << " " << col + 1;
				f << ((val < 0) ? " " : "  ");
				f.setf(std::ios_base::scientific);
				f << std::setprecision(m_precision) << val << "\n";
			}
	}
	/// digits behind the point of the scientific notation (reference: 13; 16 is lossless for fp64)
	void set_precision(int digits) { m_precision = digits; }

  private:
	static std::string lower(std::string s) { for (char& c : s) c = (char)std::tolower((unsigned char)c); return s; }
	std::string m_file;
	bool m_queried = false;
	AlgebraicType m_type = GENERAL;
	size_t m_rows = 0, m_cols = 0, m_lines = 0, m_firstDataLine = 0;
	int m_precision = 13;
};

/// Writes every vector it is handed as <dir>/<name> in ConnectionViewer form (what GridFunctionDebugWriter does with
/// its base directory, lib_disc/function_spaces/grid_function_util.h; values with 16+ digits like WriteVector)
template <typename TVector>
class ConnectionViewerVectorWriter : public IVectorDebugWriter<TVector> {
  public:
	ConnectionViewerVectorWriter(const std::string& dir, const IOPositions& pos, int dim, int precision = 0)
	    : m_dir(dir), m_pos(pos), m_dim(dim), m_precision(precision) {}
	virtual void write_vector(const TVector& vec, const char* name)
	{
		std::vector<double> h(vec.len());
		if (!h.empty()) vec.copy_to_host(h.data());
		IOPositions pos = m_pos;
		if (pos.size() < h.size()) {   // block vectors: one entry per component, the node's position repeated
			const size_t B = TVector::blockSize;
			IOPositions p; p.dim = m_pos.dim; p.resize(h.size());
			for (size_t i = 0; i < h.size(); ++i) for (int d = 0; d < 3; ++d) p.xyz[3 * i + d] = (i / B) < m_pos.size() ? m_pos[i / B][d] : 0.0;
			pos = p;
		}
		ConnectionViewer::WriteVector(m_dir + "/" + name, h.data(), h.size(), pos, m_dim, m_precision);
	}
  private:
	std::string m_dir;
	IOPositions m_pos;
	int m_dim, m_precision;
};

} // namespace ug
