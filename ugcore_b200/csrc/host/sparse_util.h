// sparse_util.h — Galerkin coarse operators for the GPU algebra's host-side matrices.
//
//   AddMultiplyOf(M, A, B, C):  M += A * B * C
//       ugbase/lib_algebra/algebra_common/sparsematrix_util.h:152-230
//   used by AssembledMultiGridCycle::init_rap_operator (gmg:set_rap(true)):
//       A_{l-1} += R_l * A_l * P_l for l = top ... base + 1
//       ugbase/lib_disc/operator/linear_operator/multi_grid_solver/mg_solver_impl.hpp:828-1013 (:959)
//
// Runs on the host at init time (everything before solver:apply stays on the CPU;
// the result is uploaded like an assembled level matrix).  The triple loop is the reference's:
// for row i, for A_ik != 0, for B_kl != 0: ab = A_ik * B_kl; for C_lj != 0: M_ij += ab * C_lj — the
// same accumulation order, so the coarse matrices are bit-identical to ugcore's
// (tests/test_rap.py checks them against the reference's own AddMultiplyOf, compiled for the tests).
// Rows are independent: OpenMP over i, each thread with its own sparse accumulator.
// A and C are scalar matrices (transfers: for block algebras ugcore stores the scalar on the
// block diagonal, which multiplies every component of the block), B and M carry blocks.
#pragma once
#include "gpu_sparsematrix.h"
#include <algorithm>
#include <vector>

namespace ug {

template <typename TBlockMatrix>
void AddMultiplyOf(TBlockMatrix& M, const GPUSparseMatrix<double>& A, const TBlockMatrix& B, const GPUSparseMatrix<double>& C)
{
	enum { Bs = TBlockMatrix::blockSize, BB = Bs * Bs };
	if (C.num_rows() != B.num_cols()) UG_THROW("AddMultiplyOf: sizes must match: nRows(C) = " << C.num_rows() << " != " << B.num_cols() << " = nCols(B)");
	if (B.num_rows() != A.num_cols()) UG_THROW("AddMultiplyOf: sizes must match: nRows(B) = " << B.num_rows() << " != " << A.num_cols() << " = nCols(A)");
	if (M.num_rows() != A.num_rows()) UG_THROW("AddMultiplyOf: row sizes mismatch: M.num_rows = " << M.num_rows() << ", A.num_rows = " << A.num_rows());
	if (M.num_cols() != C.num_cols()) UG_THROW("AddMultiplyOf: column sizes mismatch: M.num_cols = " << M.num_cols() << ", C.num_cols = " << C.num_cols());
	const std::vector<int64_t>& arp = A.crs_rowptr(); const std::vector<int>& aci = A.crs_cols(); const std::vector<double>& ava = A.crs_vals();
	const std::vector<int64_t>& brp = B.crs_rowptr(); const std::vector<int>& bci = B.crs_cols(); const std::vector<double>& bva = B.crs_vals();
	const std::vector<int64_t>& crp = C.crs_rowptr(); const std::vector<int>& cci = C.crs_cols(); const std::vector<double>& cva = C.crs_vals();
	const std::vector<int64_t>& mrp = M.crs_rowptr(); const std::vector<int>& mci = M.crs_cols(); const std::vector<double>& mva = M.crs_vals();
	const int64_t n = (int64_t)A.num_rows(), nc = (int64_t)C.num_cols();
	std::vector<std::vector<int> > rcols((size_t)n);
	std::vector<std::vector<double> > rvals((size_t)n);
#pragma omp parallel
	{
		std::vector<int> slot((size_t)nc, -1), touched;      // sparse accumulator (UnsortedSparseVector in the reference)
		std::vector<double> acc;
#pragma omp for schedule(dynamic, 64)
		for (int64_t i = 0; i < n; ++i) {
			touched.clear(); acc.clear();
			for (int64_t pa = arp[i]; pa < arp[i + 1]; ++pa) {
				const double a = ava[pa];
				if (a == 0.0) continue;
				const int64_t k = aci[pa];
				for (int64_t pb = brp[k]; pb < brp[k + 1]; ++pb) {
					const double* bv = &bva[(size_t)pb * BB];
					bool zero = true;
					for (int t = 0; t < BB; ++t) if (bv[t] != 0.0) { zero = false; break; }
					if (zero) continue;
					const int64_t l = bci[pb];
					double ab[BB];
					for (int t = 0; t < BB; ++t) ab[t] = a * bv[t];                      // AssignMult(ab, A_ik, B_kl)
					for (int64_t pc = crp[l]; pc < crp[l + 1]; ++pc) {
						const double c = cva[pc];
						if (c == 0.0) continue;
						const int j = cci[pc];
						int s = slot[j];
						if (s < 0) { s = slot[j] = (int)touched.size(); touched.push_back(j); acc.insert(acc.end(), BB, 0.0); }
						double* dst = &acc[(size_t)s * BB];
						for (int t = 0; t < BB; ++t) dst[t] = dst[t] + ab[t] * c;         // AddMult(row(j), ab, C_lj)
					}
				}
			}
			// M.add_matrix_row(i, row): added to what M holds (empty in init_rap_operator)
			std::vector<int>& oc = rcols[(size_t)i]; std::vector<double>& ov = rvals[(size_t)i];
			std::vector<std::pair<int, int> > merged;   // (column, source: >= 0 accumulator slot, < 0: -(position in M) - 1)
			for (size_t t = 0; t < touched.size(); ++t) merged.push_back(std::make_pair(touched[t], (int)t));
			std::sort(merged.begin(), merged.end());
			size_t q = 0;
			for (int64_t pm = mrp[i]; pm < mrp[i + 1] || q < merged.size();) {
				const int cm = pm < mrp[i + 1] ? mci[pm] : 2147483647, cq = q < merged.size() ? merged[q].first : 2147483647;
				const int col = std::min(cm, cq);
				oc.push_back(col);
				for (int t = 0; t < BB; ++t) {
					double v = 0.0;
					if (cm == col) v = mva[(size_t)pm * BB + t];
					if (cq == col) v = (cm == col) ? v + acc[(size_t)merged[q].second * BB + t] : acc[(size_t)merged[q].second * BB + t];
					ov.push_back(v);
				}
				if (cm == col) ++pm;
				if (cq == col) ++q;
			}
			for (int j : touched) slot[j] = -1;
		}
	}
	std::vector<int64_t> rp((size_t)n + 1, 0);
	for (int64_t i = 0; i < n; ++i) rp[(size_t)i + 1] = rp[(size_t)i] + (int64_t)rcols[(size_t)i].size();
	std::vector<int> ci((size_t)rp[(size_t)n]); std::vector<double> va((size_t)rp[(size_t)n] * BB);
	for (int64_t i = 0; i < n; ++i) {
		std::copy(rcols[(size_t)i].begin(), rcols[(size_t)i].end(), ci.begin() + rp[(size_t)i]);
		std::copy(rvals[(size_t)i].begin(), rvals[(size_t)i].end(), va.begin() + rp[(size_t)i] * BB);
	}
	M.set_from_crs((size_t)n, (size_t)nc, rp.data(), ci.data(), va.data());
}

} // namespace ug
