// ilu_factor.h — host-side pieces of ILU for the GPU algebra: factorisation, level sets, ordering.
//
//   FactorizeILUSorted / FactorizeILUBeta   ugbase/lib_algebra/operator/preconditioner/ilu.h:174-228, :110-171
//   GetCuthillMcKeeOrder                    ugbase/lib_algebra/algebra_common/permutation_util.h:96-114 ->
//   ComputeCuthillMcKeeOrder                ugbase/lib_algebra/ordering_strategies/algorithms/native_cuthill_mckee.cpp:100-300
//
// The factorisation is init-time work on the assembled matrix (everything before
// solver:apply stays on the CPU); it runs here on the defragmented CRS arrays with the reference's loop
// order and operations, so the factors are bit-identical to ugcore's (tests/test_ilu.py compares them with
// the reference's own FactorizeILUSorted / FactorizeILUBeta, compiled for the tests).  The triangular
// solves run on the device (preconditioners.h: ILU) as level-scheduled sweeps: level_sets() below groups the
// rows of a triangular factor so that a row only depends on rows of earlier groups.
#pragma once
#include "ug_base.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace ug {

namespace ilu_detail {
/// position of entry (r, c) in a CRS row with sorted columns, -1 if not stored (get_connection)
inline int64_t find(const std::vector<int64_t>& rp, const std::vector<int>& ci, int64_t r, int c)
{
	const int* b = ci.data() + rp[(size_t)r];
	const int* e = ci.data() + rp[(size_t)r + 1];
	const int* p = std::lower_bound(b, e, c);
	return (p != e && *p == c) ? (int64_t)(p - ci.data()) : -1;
}
} // namespace ilu_detail

// ---- B x B blocks, column-major like FixedArray2 (small_algebra/storage/fixed_array_impl.h:182-203): the few
//      DenseMatrix operations the factorisation uses, with the reference's evaluation order ----
namespace ilu_detail {
#define UG_BLK(m, r, c) (m)[(r) + B * (c)]
/// Invert(mat) for 1x1 / 2x2 / 3x3 (small_matrix/densematrix_inverse.h:75-86, 121-147, 172-229); false if singular
template <int B> inline bool blk_invert(double* m)
{
	if (B == 1) { if (m[0] == 0.0) return false; m[0] = 1.0 / m[0]; return true; }
	if (B == 2) {
		double invdet = UG_BLK(m, 0, 0) * UG_BLK(m, 1, 1) - UG_BLK(m, 1, 0) * UG_BLK(m, 0, 1);
		if (invdet == 0.0) return false;
		invdet = 1.0 / invdet;
		std::swap(UG_BLK(m, 0, 0), UG_BLK(m, 1, 1));
		UG_BLK(m, 0, 0) *= invdet; UG_BLK(m, 0, 1) *= -invdet; UG_BLK(m, 1, 0) *= -invdet; UG_BLK(m, 1, 1) *= invdet;
		return true;
	}
	double invdet = UG_BLK(m, 0, 0) * UG_BLK(m, 1, 1) * UG_BLK(m, 2, 2) + UG_BLK(m, 0, 1) * UG_BLK(m, 1, 2) * UG_BLK(m, 2, 0) +
	                UG_BLK(m, 0, 2) * UG_BLK(m, 1, 0) * UG_BLK(m, 2, 1) - UG_BLK(m, 0, 0) * UG_BLK(m, 1, 2) * UG_BLK(m, 2, 1) -
	                UG_BLK(m, 0, 1) * UG_BLK(m, 1, 0) * UG_BLK(m, 2, 2) - UG_BLK(m, 0, 2) * UG_BLK(m, 1, 1) * UG_BLK(m, 2, 0);
	if (invdet == 0.0) return false;
	invdet = 1.0 / invdet;
	double inv[9];
	UG_BLK(inv, 0, 0) = ( UG_BLK(m, 1, 1) * UG_BLK(m, 2, 2) - UG_BLK(m, 1, 2) * UG_BLK(m, 2, 1)) * invdet;
	UG_BLK(inv, 0, 1) = (-UG_BLK(m, 0, 1) * UG_BLK(m, 2, 2) + UG_BLK(m, 0, 2) * UG_BLK(m, 2, 1)) * invdet;
	UG_BLK(inv, 0, 2) = ( UG_BLK(m, 0, 1) * UG_BLK(m, 1, 2) - UG_BLK(m, 0, 2) * UG_BLK(m, 1, 1)) * invdet;
	UG_BLK(inv, 1, 0) = (-UG_BLK(m, 1, 0) * UG_BLK(m, 2, 2) + UG_BLK(m, 1, 2) * UG_BLK(m, 2, 0)) * invdet;
	UG_BLK(inv, 1, 1) = ( UG_BLK(m, 0, 0) * UG_BLK(m, 2, 2) - UG_BLK(m, 0, 2) * UG_BLK(m, 2, 0)) * invdet;
	UG_BLK(inv, 1, 2) = (-UG_BLK(m, 0, 0) * UG_BLK(m, 1, 2) + UG_BLK(m, 0, 2) * UG_BLK(m, 1, 0)) * invdet;
	UG_BLK(inv, 2, 0) = ( UG_BLK(m, 1, 0) * UG_BLK(m, 2, 1) - UG_BLK(m, 1, 1) * UG_BLK(m, 2, 0)) * invdet;
	UG_BLK(inv, 2, 1) = (-UG_BLK(m, 0, 0) * UG_BLK(m, 2, 1) + UG_BLK(m, 0, 1) * UG_BLK(m, 2, 0)) * invdet;
	UG_BLK(inv, 2, 2) = ( UG_BLK(m, 0, 0) * UG_BLK(m, 1, 1) - UG_BLK(m, 0, 1) * UG_BLK(m, 1, 0)) * invdet;
	for (int t = 0; t < 9; ++t) m[t] = inv[t];
	return true;
}
/// erg = a * b (DenseMatrix::operator*, densematrix_impl.h:251-267: erg(r,c) = 0; erg(r,c) += a(r,i) * b(i,c))
template <int B> inline void blk_mul(double* erg, const double* a, const double* b)
{
	for (int r = 0; r < B; ++r)
		for (int c = 0; c < B; ++c) {
			double e = 0.0;
			for (int i = 0; i < B; ++i) e += UG_BLK(a, r, i) * UG_BLK(b, i, c);
			UG_BLK(erg, r, c) = e;
		}
}
/// a /= d: a = a * d^-1 (densematrix_impl.h:191-200)
template <int B> inline void blk_div(double* a, const double* d)
{
	double tmp[B * B], erg[B * B];
	for (int t = 0; t < B * B; ++t) tmp[t] = d[t];
	if (!blk_invert<B>(tmp)) UG_THROW("Failed to invert dense matrix.");
	blk_mul<B>(erg, a, tmp);
	for (int t = 0; t < B * B; ++t) a[t] = erg[t];
}
/// a -= b * c
template <int B> inline void blk_sub_mul(double* a, const double* b, const double* c)
{
	double erg[B * B];
	blk_mul<B>(erg, b, c);
	for (int r = 0; r < B; ++r) for (int cc = 0; cc < B; ++cc) UG_BLK(a, r, cc) -= UG_BLK(erg, r, cc);
}
/// BlockNorm (small_algebra/blocks.h:51-60): Frobenius norm
template <int B> inline double blk_norm(const double* a)
{
	double s = 0.0;
	for (int t = 0; t < B * B; ++t) s += a[t] * a[t];
	return std::sqrt(s);
}
#undef UG_BLK
} // namespace ilu_detail

/// ILU(0) on the stored pattern, rows sorted (ilu.h:174-228); B x B block entries, B*B doubles each
template <int B>
inline void FactorizeILUSortedBlock(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va,
                                    number eps = 1e-50)
{
	const int BB = B * B;
	std::vector<int64_t> diag((size_t)n);
	for (int64_t i = 0; i < n; ++i) diag[(size_t)i] = ilu_detail::find(rp, ci, i, (int)i);
	for (int64_t i = 1; i < n; ++i) {
		for (int64_t pik = rp[(size_t)i]; pik != rp[(size_t)i + 1] && ci[(size_t)pik] < i; ++pik) {
			const int k = ci[(size_t)pik];
			if (diag[(size_t)k] < 0) UG_THROW("ILU: row " << k << " has no diagonal entry");
			double* a_ik = &va[(size_t)pik * BB];
			const double* a_kk = &va[(size_t)diag[(size_t)k] * BB];
			if (std::fabs(ilu_detail::blk_norm<B>(a_kk)) < eps * ilu_detail::blk_norm<B>(a_ik))
				UG_THROW("ILU: Blocknorm of diagonal is near-zero for k=" << k << " with eps: " << eps);
			ilu_detail::blk_div<B>(a_ik, a_kk);
			int64_t pij = pik + 1, pkj = rp[(size_t)k];
			const int64_t ei = rp[(size_t)i + 1], ek = rp[(size_t)k + 1];
			while (pij != ei && pkj != ek) {
				if (ci[(size_t)pij] > ci[(size_t)pkj]) ++pkj;
				else if (ci[(size_t)pij] < ci[(size_t)pkj]) ++pij;
				else { ilu_detail::blk_sub_mul<B>(&va[(size_t)pij * BB], a_ik, &va[(size_t)pkj * BB]); ++pkj; ++pij; }
			}
		}
	}
}

/// ILU(beta) with blocks (ilu.h:110-171)
template <int B>
inline void FactorizeILUBetaBlock(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va, number beta)
{
	const int BB = B * B;
	for (int64_t i = 1; i < n; ++i) {
		const int64_t dii = ilu_detail::find(rp, ci, i, (int)i);
		if (dii < 0) UG_THROW("ILU: row " << i << " has no diagonal entry");
		double Nii[BB];
		for (int t = 0; t < BB; ++t) { Nii[t] = va[(size_t)dii * BB + t]; Nii[t] *= 0.0; }
		for (int64_t pik = rp[(size_t)i]; pik != rp[(size_t)i + 1] && ci[(size_t)pik] < i; ++pik) {
			const int k = ci[(size_t)pik];
			const int64_t dkk = ilu_detail::find(rp, ci, k, k);
			if (dkk < 0) UG_THROW("ILU: row " << k << " has no diagonal entry");
			double* a_ik = &va[(size_t)pik * BB];
			ilu_detail::blk_div<B>(a_ik, &va[(size_t)dkk * BB]);
			for (int64_t pkj = rp[(size_t)k]; pkj != rp[(size_t)k + 1]; ++pkj) {
				const int j = ci[(size_t)pkj];
				if (j <= k) continue;
				const int64_t pij = ilu_detail::find(rp, ci, i, j);
				if (pij >= 0) ilu_detail::blk_sub_mul<B>(&va[(size_t)pij * BB], a_ik, &va[(size_t)pkj * BB]);
				else ilu_detail::blk_sub_mul<B>(Nii, a_ik, &va[(size_t)pkj * BB]);
			}
		}
		for (int t = 0; t < BB; ++t) va[(size_t)dii * BB + t] += beta * Nii[t];   // AddMult(Aii, beta, Nii)
	}
}

/// dispatch on the block size (1: the scalar routines below)
inline void FactorizeILUSorted(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va, number eps);
inline void FactorizeILUBeta(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va, number beta);
inline void FactorizeILU(int block, int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va, number beta,
                         number eps)
{
	if (block == 1) { if (beta != 0.0) FactorizeILUBeta(n, rp, ci, va, beta); else FactorizeILUSorted(n, rp, ci, va, eps); }
	else if (block == 2) { if (beta != 0.0) FactorizeILUBetaBlock<2>(n, rp, ci, va, beta); else FactorizeILUSortedBlock<2>(n, rp, ci, va, eps); }
	else if (block == 3) { if (beta != 0.0) FactorizeILUBetaBlock<3>(n, rp, ci, va, beta); else FactorizeILUSortedBlock<3>(n, rp, ci, va, eps); }
	else UG_THROW("ILU: block size must be 1, 2 or 3");
}

/// ILU(0) on the stored pattern, rows sorted (ilu.h:174-228); scalar entries
inline void FactorizeILUSorted(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va,
                               number eps)
{
	std::vector<int64_t> diag((size_t)n);
	for (int64_t i = 0; i < n; ++i) diag[(size_t)i] = ilu_detail::find(rp, ci, i, (int)i);
	for (int64_t i = 1; i < n; ++i) {
		// eliminate all entries A(i, k) with k < i using the rows A(k, .)
		for (int64_t pik = rp[(size_t)i]; pik != rp[(size_t)i + 1] && ci[(size_t)pik] < i; ++pik) {
			const int k = ci[(size_t)pik];
			if (diag[(size_t)k] < 0) UG_THROW("ILU: row " << k << " has no diagonal entry");
			const double a_kk = va[(size_t)diag[(size_t)k]];
			if (std::fabs(a_kk) < eps * std::fabs(va[(size_t)pik]))
				UG_THROW("ILU: Blocknorm of diagonal is near-zero for k=" << k << " with eps: " << eps << ", ||A_kk||=" << std::fabs(a_kk)
				         << ", ||A_ik||=" << std::fabs(va[(size_t)pik]));
			va[(size_t)pik] /= a_kk;                       // A(i,k) /= A(k,k), kept as the entry of L
			const double a_ik = va[(size_t)pik];
			// A(i, j) -= A(i, k) * A(k, j) for the j > k stored in both rows: merge of two sorted rows
			int64_t pij = pik + 1, pkj = rp[(size_t)k];
			const int64_t ei = rp[(size_t)i + 1], ek = rp[(size_t)k + 1];
			while (pij != ei && pkj != ek) {
				if (ci[(size_t)pij] > ci[(size_t)pkj]) ++pkj;
				else if (ci[(size_t)pij] < ci[(size_t)pkj]) ++pij;
				else { va[(size_t)pij] -= a_ik * va[(size_t)pkj]; ++pkj; ++pij; }
			}
		}
	}
}

/// ILU(0) whose fill-in is lumped onto the diagonal with weight beta (ilu.h:110-171); scalar entries
inline void FactorizeILUBeta(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, std::vector<double>& va, number beta)
{
	for (int64_t i = 1; i < n; ++i) {
		const int64_t dii = ilu_detail::find(rp, ci, i, (int)i);
		if (dii < 0) UG_THROW("ILU: row " << i << " has no diagonal entry");
		double Nii = va[(size_t)dii]; Nii *= 0.0;
		for (int64_t pik = rp[(size_t)i]; pik != rp[(size_t)i + 1] && ci[(size_t)pik] < i; ++pik) {
			const int k = ci[(size_t)pik];
			const int64_t dkk = ilu_detail::find(rp, ci, k, k);
			if (dkk < 0) UG_THROW("ILU: row " << k << " has no diagonal entry");
			va[(size_t)pik] /= va[(size_t)dkk];
			const double a_ik = va[(size_t)pik];
			for (int64_t pkj = rp[(size_t)k]; pkj != rp[(size_t)k + 1]; ++pkj) {
				const int j = ci[(size_t)pkj];
				if (j <= k) continue;                      // L part of row k
				const double a_kj = va[(size_t)pkj];
				const int64_t pij = ilu_detail::find(rp, ci, i, j);
				if (pij >= 0) va[(size_t)pij] -= a_ik * a_kj;   // inside the pattern: standard elimination
				else Nii -= a_ik * a_kj;                        // outside: lumped onto the diagonal
			}
		}
		va[(size_t)dii] += beta * Nii;                     // AddMult(Aii, beta, Nii)
	}
}

/// Level sets of a triangular factor given by its stored pattern.  lower = true: row i depends on the stored
/// columns j < i, level(i) = 1 + max level(j); lower = false: on the stored j > i.  Returns the number of
/// levels and fills level[].
inline int level_sets(int64_t n, const std::vector<int64_t>& rp, const std::vector<int>& ci, bool lower, std::vector<int>& level)
{
	level.assign((size_t)n, 0);
	int nlev = n > 0 ? 1 : 0;
	if (lower) {
		for (int64_t i = 0; i < n; ++i) {
			int l = 0;
			for (int64_t p = rp[(size_t)i]; p != rp[(size_t)i + 1] && ci[(size_t)p] < i; ++p) l = std::max(l, level[(size_t)ci[(size_t)p]] + 1);
			level[(size_t)i] = l; nlev = std::max(nlev, l + 1);
		}
	} else {
		for (int64_t i = n - 1; i >= 0; --i) {
			int l = 0;
			for (int64_t p = rp[(size_t)i + 1] - 1; p >= rp[(size_t)i] && ci[(size_t)p] > i; --p) l = std::max(l, level[(size_t)ci[(size_t)p]] + 1);
			level[(size_t)i] = l; nlev = std::max(nlev, l + 1);
		}
	}
	return nlev;
}

/// newIndex[old] = new.  Cuthill-McKee on the graph of the stored pattern, every stored column of a row being
/// a neighbour (the diagonal included), exactly as GetCuthillMcKeeOrder / ComputeCuthillMcKeeOrder do it:
/// neighbour lists and start candidates stable-sorted by degree, breadth-first numbering, optional reversal;
/// unconnected indices go to the end (bPreserveConsec = false) or keep their place (true).
inline void GetCuthillMcKeeOrder(int64_t n, const int64_t* rp, const int* ci, std::vector<size_t>& newIndex, bool reverse = true,
                                 bool bPreserveConsec = false)
{
	const size_t nDoF = (size_t)n;
	std::vector<std::vector<size_t> > con(nDoF);
	for (size_t i = 0; i < nDoF; ++i) con[i].assign(ci + rp[i], ci + rp[i + 1]);
	auto byDegree = [&con](size_t a, size_t b) { return con[a].size() < con[b].size(); };
	std::vector<char> handled(nDoF, 0);
	for (size_t i = 0; i < nDoF; ++i) {
		if (con[i].empty()) handled[i] = 1;
		else std::stable_sort(con[i].begin(), con[i].end(), byDegree);
	}
	std::vector<size_t> start(nDoF);
	for (size_t i = 0; i < nDoF; ++i) start[i] = i;
	std::stable_sort(start.begin(), start.end(), byDegree);
	std::vector<size_t> order, queue;
	order.reserve(nDoF);
	size_t firstNonHandled = 0;
	for (;;) {
		while (firstNonHandled < nDoF && handled[start[firstNonHandled]]) ++firstNonHandled;
		if (firstNonHandled == nDoF) break;
		queue.assign(1, start[firstNonHandled]);
		for (size_t head = 0; head < queue.size(); ++head) {
			const size_t front = queue[head];
			if (handled[front]) continue;
			order.push_back(front); handled[front] = 1;
			for (size_t t = 0; t < con[front].size(); ++t) if (!handled[con[front][t]]) queue.push_back(con[front][t]);
		}
	}
	const size_t none = (size_t)-1;
	newIndex.assign(nDoF, none);
	const size_t sz = order.size();
	if (bPreserveConsec) {
		size_t cnt = 0;
		for (size_t newInd = 0; newInd < nDoF; ++newInd) {
			if (con[newInd].empty()) continue;
			newIndex[reverse ? order[sz - 1 - cnt] : order[cnt]] = newInd;
			++cnt;
		}
		if (cnt != sz) UG_THROW("OrderCuthillMcKee: Not all indices sorted that must be sorted: " << cnt << " written, but should write: " << sz);
		// smallest block of consecutive indices of which only the first carries connections (findBlockSize)
		auto gcd = [](size_t a, size_t b) { while (b) { const size_t r = a % b; a = b; b = r; } return a; };
		size_t cd = 0, blockSize;
		while (cd < nDoF && con[cd].empty()) ++cd;
		if (cd == nDoF) blockSize = nDoF;
		else {
			size_t run = 1;
			for (size_t i = cd + 1; i < nDoF; ++i) {
				if (con[i].empty()) { ++run; continue; }
				cd = gcd(run, cd);
				run = 1;
			}
			blockSize = gcd(run, cd);
		}
		for (size_t i = 0; i < nDoF; i += blockSize) if (newIndex[i] == none) newIndex[i] = i;
		for (size_t i = 0; i < nDoF; i += blockSize) for (size_t j = 1; j < blockSize; ++j) newIndex[i + j] = newIndex[i] + j;
	} else {
		for (size_t i = 0; i < sz; ++i) newIndex[reverse ? order[sz - 1 - i] : order[i]] = i;
		size_t next = sz;
		for (size_t i = 0; i < nDoF; ++i) if (newIndex[i] == none) newIndex[i] = next++;
	}
}

} // namespace ug
