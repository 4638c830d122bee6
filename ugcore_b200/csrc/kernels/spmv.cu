// spmv.cu — SELL-32 upload and the SpMV family (SparseMatrix<T>::axpy and friends).
//
// Reference semantics (paths relative to /root/reference/ugbase):
//   lib_algebra/cpu_algebra/sparsematrix_impl.h:291-339  axpy (apply / matmul_minus)
//   lib_algebra/cpu_algebra/sparsematrix_impl.h:271-288  apply_ignore_zero_rows
//   lib_algebra/cpu_algebra/sparsematrix_impl.h:257-268  mat_mult_add_row
//   lib_algebra/small_algebra/small_matrix/densematrix_operations.h:56-79 (block MatMult/MatMultAdd)
// Every row is accumulated by ONE thread in ascending column order with separate
// multiply and add (-fmad=false), i.e. exactly the CPU operation sequence: results
// are bit-identical to ugcore's CPUAlgebra / CPUBlockAlgebra<N>.
//
// Layout: SELL-32.  A slice is 32 consecutive rows; entry k of lane l of slice s sits
// at slice_ptr[s] + 32*k + l, so a warp reads 256 contiguous bytes of values and 128
// contiguous bytes of column indices per k (fully coalesced, streamed with evict-first
// loads); the x-gather goes through the read-only path and lives in L1/L2.
// Roofline: HBM.  Algorithmic bytes per launch:
//   y = A x : 12*nnz + 4*(n+1) + 8*ncols + 8*n      y -= A x : + 8*n
// Grid: persistent, (#SMs * 8) CTAs of 8 warps striding over the slices.
#include "../common.cuh"
#include <algorithm>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

enum { MODE_ASSIGN = 0, MODE_ASSIGN_SKIP_EMPTY = 1, MODE_INPLACE = 2, MODE_GENERAL = 3 };
enum { FUSE_NONE = 0, FUSE_DOT = 1, FUSE_JACOBI = 2, FUSE_RESTRICT_JACOBI = 3 };

struct Sell {
	const int64_t* slice_ptr; const int* rowlen; const int* cols; const double* vals;
	int64_t nrows, num_slices;
	// value-indexed stream (COMP kernels): word = column offset << 16 | dictionary index << vshift
	const unsigned int* vc; const int* colbase; const double* dict; int ndict; int vshift;
	// x-staged stream (spmv1_xs_kernel): word = staging position << 16 | dictionary index << 3
	const unsigned int* xw; const int4* xs_hdr; const int2* xs_runs; int xs_rmax;
};
struct Fuse {
	// FUSE_DOT (ar.nranks > 1: the last block also sums over the ranks through the peer windows)
	double* partials; unsigned int* counter; ug4b200_fin fin; UgAr ar;
	// FUSE_JACOBI
	const double* diaginv; double* st_out; double* sc; int flags;
	// FUSE_JACOBI / FUSE_RESTRICT_JACOBI in partitioned runs: interface rows of st_out are also stored
	// into the neighbours' peer windows (common.cuh: ug_push_row / ug_push_finish); nullptr otherwise
	const UgPushDev* push;
};

template <int BETAK> __device__ __forceinline__ double mulbeta(double a, double beta)
{ return BETAK == 1 ? a : (BETAK == -1 ? -a : beta * a); }

// ---------------------------------------------------------------- scalar (block 1)
// One warp per slice.  All per-row streams (dest, sc, diaginv, own st entry) are requested
// BEFORE the entry loop so their HBM latency overlaps the matrix stream; entries are fetched
// in batches of UNR with the next batch's values/columns requested before the current
// batch's x-gather is consumed (software pipelining: two dependent round trips overlap).
#ifndef UG_UNR
#define UG_UNR 4
#endif
#ifndef UG_PIPE
#define UG_PIPE 0
#endif
#ifndef UG_MINBLK
#define UG_MINBLK 1
#endif
constexpr int UNR = UG_UNR;

template <int BETAK, int MODE, int FUSE, bool COMP>
__global__ void __launch_bounds__(kThreads, UG_MINBLK)
spmv1_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
             Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	const int lane = threadIdx.x & 31;
	const int64_t gwarp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
	const int64_t nwarps = (int64_t)gridDim.x * kWarps;
	double dot = 0.0;
	const UgPushDev* push = (FUSE == FUSE_JACOBI || FUSE == FUSE_RESTRICT_JACOBI) ? fz.push : nullptr;
	const unsigned long long pe = push ? *(volatile unsigned long long*)push->epoch + 1ull : 0ull;
	for (int64_t s = gwarp; s < A.num_slices; s += nwarps) {
		const int64_t row = s * 32 + lane;
		const int64_t base = A.slice_ptr[s];
		const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
		const int len = A.rowlen[row];
		const unsigned int pmask = push ? push->rowmask[s] : 0u;
		const double* vp = A.vals + base + lane;
		const int* cp = A.cols + base + lane;
		const unsigned int* vcp = A.vc + base + lane;
		const int cbase = COMP ? A.colbase[s] : 0;
		const bool live = row < A.nrows;
		// entry k of this lane: value and column (COMP: dictionary look-up, column = slice base + offset)
		auto ld_entry = [&](int64_t k, double& av, int& cv) {
			if (COMP) { const unsigned int e = __ldcs(vcp + k * 32); av = __ldg(A.dict + ((e & 0xffffu) >> A.vshift)); cv = cbase + (int)(e >> 16); }
			else { av = ug_ld_stream(vp + k * 32); cv = ug_ld_stream(cp + k * 32); }
		};
		// ---- first batch of the matrix stream
		double a[UNR]; int c[UNR];
#pragma unroll
		for (int u = 0; u < UNR; ++u)
			if (u < width) ld_entry(u, a[u], c[u]);
		// ---- per-row streams, hoisted
		double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
		if (MODE == MODE_INPLACE) { if (live) acc = dest[row]; }
		else if (MODE == MODE_GENERAL) { if (live) acc = alpha * v[row]; }
		if (FUSE == FUSE_DOT) { if (live) own = w[row]; }
		if (FUSE == FUSE_RESTRICT_JACOBI && live) { dinv = fz.diaginv[row]; if (len == 0) own = dest[row]; }
		if (FUSE == FUSE_JACOBI && live) {
			if (fz.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
			if ((fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) scv = fz.sc[row];
			if (fz.flags & UG4B200_SMOOTH_JACOBI) dinv = fz.diaginv[row];
		}
		for (int k = 0; k < width; k += UNR) {
			double x[UNR];
#pragma unroll
			for (int u = 0; u < UNR; ++u)
				if (k + u < len) x[u] = __ldg(w + c[u]);
			double an[UNR]; int cn[UNR];
#if UG_PIPE
#pragma unroll
			for (int u = 0; u < UNR; ++u)
				if (k + UNR + u < width) ld_entry(k + UNR + u, an[u], cn[u]);
#endif
#pragma unroll
			for (int u = 0; u < UNR; ++u) {
				if (k + u < len) {
					const double t = mulbeta<BETAK>(a[u], beta) * x[u];
					// MatMult(dest[i], beta, a_0, w[c_0]): the first connection is assigned, not added
					if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && k + u == 0) acc = t;
					else acc = acc + t;
				}
			}
#if !UG_PIPE
#pragma unroll
			for (int u = 0; u < UNR; ++u)
				if (k + UNR + u < width) ld_entry(k + UNR + u, an[u], cn[u]);
#endif
#pragma unroll
			for (int u = 0; u < UNR; ++u) { a[u] = an[u]; c[u] = cn[u]; }
		}
		if (FUSE == FUSE_JACOBI) {
			if (live) {
				dest[row] = acc;
				if (fz.flags & UG4B200_SMOOTH_ADD_IN) scv = scv + own;
				if (fz.flags & UG4B200_SMOOTH_JACOBI) {
					const double st = dinv * acc;   // MatMult(c[i], 1.0, diagInv[i], d[i])
					fz.st_out[row] = st;
					if ((pmask >> lane) & 1u) ug_push_row(push, pe, s, lane, pmask, st);
					if (fz.flags & UG4B200_SMOOTH_ADD_OUT) scv = scv + st;
				}
				if (fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) fz.sc[row] = scv;
			}
		} else if (FUSE == FUSE_RESTRICT_JACOBI) {
			// coarse defect by restriction (rows without connections keep their value), then the first
			// Jacobi step of the coarse level on it: st = diagInv * sd
			if (live) {
				double dv = own;
				if (len > 0) { dest[row] = acc; dv = acc; }
				const double st = dinv * dv;
				fz.st_out[row] = st;
				if ((pmask >> lane) & 1u) ug_push_row(push, pe, s, lane, pmask, st);
			}
		} else {
			if (live && (MODE != MODE_ASSIGN_SKIP_EMPTY || len > 0)) dest[row] = acc;
			if (FUSE == FUSE_DOT && live) dot += acc * own;
		}
	}
	if (FUSE == FUSE_DOT) ug_block_reduce_fin(dot, fz.partials, fz.counter, fz.fin, fz.ar);
}

// ---------------------------------------------------------------- block B x B
// values of entry e: component (r,c) at vals[(e - lane)*B*B + (r + B*c)*32 + lane]
template <int B, int BETAK, int MODE, int FUSE>
__global__ void __launch_bounds__(kThreads)
spmvB_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
             Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	constexpr int BB = B * B;
	const int lane = threadIdx.x & 31;
	const int64_t gwarp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
	const int64_t nwarps = (int64_t)gridDim.x * kWarps;
	double dot = 0.0;
	for (int64_t s = gwarp; s < A.num_slices; s += nwarps) {
		const int64_t row = s * 32 + lane;
		const int len = A.rowlen[row];
		const int64_t base = A.slice_ptr[s];
		const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
		const double* vp = A.vals + base * BB + lane;
		const int* cp = A.cols + base + lane;
		const bool live = row < A.nrows;
		double acc[B];
		int k = 0;
		if (MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) {
#pragma unroll
			for (int r = 0; r < B; ++r) acc[r] = 0.0;
			if (width > 0 && len > 0) {
				const int c0 = ug_ld_stream(cp);
				double x[B];
#pragma unroll
				for (int t = 0; t < B; ++t) x[t] = __ldg(w + (int64_t)c0 * B + t);
#pragma unroll
				for (int r = 0; r < B; ++r) {
					acc[r] = mulbeta<BETAK>(ug_ld_stream(vp + (r + B * 0) * 32), beta) * x[0];
#pragma unroll
					for (int c = 1; c < B; ++c)
						acc[r] = acc[r] + mulbeta<BETAK>(ug_ld_stream(vp + (r + B * c) * 32), beta) * x[c];
				}
			}
			k = 1;
		} else if (MODE == MODE_INPLACE) {
#pragma unroll
			for (int r = 0; r < B; ++r) acc[r] = live ? dest[row * B + r] : 0.0;
		} else {
#pragma unroll
			for (int r = 0; r < B; ++r) acc[r] = live ? alpha * v[row * B + r] : 0.0;
		}
		for (; k < width; ++k) {
			if (k < len) {
				const int c0 = ug_ld_stream(cp + (int64_t)k * 32);
				const double* ve = vp + (int64_t)k * 32 * BB;
				double a[BB]; double x[B];
#pragma unroll
				for (int q = 0; q < BB; ++q) a[q] = ug_ld_stream(ve + q * 32);
#pragma unroll
				for (int t = 0; t < B; ++t) x[t] = __ldg(w + (int64_t)c0 * B + t);
#pragma unroll
				for (int r = 0; r < B; ++r)
#pragma unroll
					for (int c = 0; c < B; ++c) acc[r] = acc[r] + mulbeta<BETAK>(a[r + B * c], beta) * x[c];
			}
		}
		if (FUSE == FUSE_JACOBI) {
			if (live) {
				const bool touch_sc = fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT);
				double scv[B];
#pragma unroll
				for (int r = 0; r < B; ++r) {
					dest[row * B + r] = acc[r];
					scv[r] = (touch_sc && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) ? fz.sc[row * B + r] : 0.0;
					if (fz.flags & UG4B200_SMOOTH_ADD_IN) scv[r] = scv[r] + w[row * B + r];
				}
				if (fz.flags & UG4B200_SMOOTH_JACOBI) {
					const double* D = fz.diaginv + row * BB;
#pragma unroll
					for (int r = 0; r < B; ++r) {
						double st = D[r] * acc[0];
#pragma unroll
						for (int c = 1; c < B; ++c) st = st + D[r + B * c] * acc[c];
						fz.st_out[row * B + r] = st;
						if (fz.flags & UG4B200_SMOOTH_ADD_OUT) scv[r] = scv[r] + st;
					}
				}
				if (touch_sc) {
#pragma unroll
					for (int r = 0; r < B; ++r) fz.sc[row * B + r] = scv[r];
				}
			}
		} else {
			if (live && (MODE != MODE_ASSIGN_SKIP_EMPTY || len > 0)) {
#pragma unroll
				for (int r = 0; r < B; ++r) dest[row * B + r] = acc[r];
			}
			if (FUSE == FUSE_DOT && live) {
				double l = 0.0;
#pragma unroll
				for (int r = 0; r < B; ++r) l += acc[r] * w[row * B + r];
				dot += l;
			}
		}
	}
	if (FUSE == FUSE_DOT) ug_block_reduce_fin(dot, fz.partials, fz.counter, fz.fin, fz.ar);
}

// ------------------------------------------ scalar matrix acting on VB-block vectors
// (P / R of a block algebra: the scalar sits on the block diagonal, so each component
// sees the scalar row product; zero off-diagonal terms add exact zeros)
template <int VB, int BETAK, int MODE>
__global__ void __launch_bounds__(kThreads)
spmv1xV_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
               const int* guard)
{
	if (ug_guarded(guard)) return;
	const int lane = threadIdx.x & 31;
	const int64_t gwarp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
	const int64_t nwarps = (int64_t)gridDim.x * kWarps;
	for (int64_t s = gwarp; s < A.num_slices; s += nwarps) {
		const int64_t row = s * 32 + lane;
		const int len = A.rowlen[row];
		const int64_t base = A.slice_ptr[s];
		const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
		const double* vp = A.vals + base + lane;
		const int* cp = A.cols + base + lane;
		const bool live = row < A.nrows;
		double acc[VB];
		int k = 0;
		if (MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) {
#pragma unroll
			for (int t = 0; t < VB; ++t) acc[t] = 0.0;
			if (width > 0 && len > 0) {
				const double a0 = mulbeta<BETAK>(ug_ld_stream(vp), beta); const int c0 = ug_ld_stream(cp);
#pragma unroll
				for (int t = 0; t < VB; ++t) acc[t] = a0 * __ldg(w + (int64_t)c0 * VB + t);
			}
			k = 1;
		} else if (MODE == MODE_INPLACE) {
#pragma unroll
			for (int t = 0; t < VB; ++t) acc[t] = live ? dest[row * VB + t] : 0.0;
		} else {
#pragma unroll
			for (int t = 0; t < VB; ++t) acc[t] = live ? alpha * v[row * VB + t] : 0.0;
		}
		for (; k < width; ++k) {
			if (k < len) {
				const double a = mulbeta<BETAK>(ug_ld_stream(vp + (int64_t)k * 32), beta);
				const int c = ug_ld_stream(cp + (int64_t)k * 32);
#pragma unroll
				for (int t = 0; t < VB; ++t) acc[t] = acc[t] + a * __ldg(w + (int64_t)c * VB + t);
			}
		}
		if (live && (MODE != MODE_ASSIGN_SKIP_EMPTY || len > 0)) {
#pragma unroll
			for (int t = 0; t < VB; ++t) dest[row * VB + t] = acc[t];
		}
	}
}

// One warp per slice, many small CTAs: the hardware block scheduler balances the tail
// (a fixed persistent grid was measured at 1.6 waves: 44 % of HBM peak, profiles/r01).
// Only the fused-reduction variant is capped by the size of the partial-sum buffer.
inline int spmv_grid(const ug4b200_ctx*, int64_t num_slices)
{
	int64_t b = (num_slices + kWarps - 1) / kWarps;
	if (b > kMaxReduceBlocks) b = kMaxReduceBlocks;
	if (b < 1) b = 1;
	return (int)b;
}

inline Sell view(const ug4b200_matrix* A)
{ return Sell{A->slice_ptr, A->rowlen, A->cols, A->vals, A->nrows, A->num_slices, A->vc, A->colbase, A->dict, A->ndict, A->vshift,
              A->xw, A->xs_hdr, A->xs_runs, A->xs_rmax}; }

#include "spmv_tma.cuh"


// Large matrices: persistent bulk-copy-staged kernels, one wave of (#SMs x resident CTAs).
template <typename K>
int launch_persistent(ug4b200_ctx* ctx, K kernel, int* ctas_per_sm, int threads, int smem, int wpb, const Sell& S, double* dest,
                      const double* v, double alpha, double beta, const double* w, const Fuse& fz, bool* used)
{
	*used = false;
	if (*ctas_per_sm == -1) {
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		int n = 0;
		if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
		if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
		*ctas_per_sm = n;
	}
	if (*ctas_per_sm <= 0) return UG4B200_OK;
	int64_t grid = (int64_t)ctx->num_sms * *ctas_per_sm;
	if (grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;
	if (S.num_slices < grid * wpb * ctx->tma_min_slices_per_warp) return UG4B200_OK; // too small to fill the pipeline
	UG_LAUNCH(ctx, kernel, (int)grid, threads, smem, S, dest, v, alpha, beta, w, fz, ctx->guard);
	*used = true;
	return UG4B200_OK;
}
template <int BETAK, int MODE, int FUSE, bool COMP>
int launch_tma(ug4b200_ctx* ctx, const Sell& S, double* dest, const double* v, double alpha, double beta,
               const double* w, const Fuse& fz, bool* used)
{
	if constexpr (COMP) {
		typedef tma::VCfg C;
		if (S.ndict <= tma::SDICT_MAX && S.vshift == 3) {
			static int cps = -1;   // per instantiation
			return launch_persistent(ctx, tma::spmv1_vi_kernel<BETAK, MODE, FUSE, true>, &cps, C::WPB * 32, C::SMEM_BYTES_SDICT, C::WPB,
			                         S, dest, v, alpha, beta, w, fz, used);
		}
		static int cps = -1;
		return launch_persistent(ctx, tma::spmv1_vi_kernel<BETAK, MODE, FUSE, false>, &cps, C::WPB * 32, C::SMEM_BYTES, C::WPB,
		                         S, dest, v, alpha, beta, w, fz, used);
	} else {
		static int cps = -1;
		return launch_persistent(ctx, tma::spmv1_tma_kernel<BETAK, MODE, FUSE>, &cps, tma::WPB * 32, tma::Cfg::SMEM_BYTES, tma::WPB,
		                         S, dest, v, alpha, beta, w, fz, used);
	}
}

template <int BETAK, int MODE, int FUSE, bool COMP>
int launch_scalar(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, const double* v, double alpha, double beta,
                  const double* w, const Fuse& fz)
{
	const int grid = spmv_grid(ctx, A->num_slices);
	Sell S = view(A);
	if (ctx->no_xs) S.xw = nullptr;
	bool used = false;
	// x-staged stream (x operand through shared memory by bulk copies: needs a 16-byte aligned x — anything from
	// ug4b200_alloc is): taken when it exists, i.e. when the value-indexed stream does not (or on request)
	if (S.xw != nullptr && !ctx->no_tma && ((uintptr_t)w & 15u) == 0) {
		typedef tma::XCfg X;
		static int cps = -1;   // per instantiation
		const int rc = launch_persistent(ctx, tma::spmv1_xs_kernel<BETAK, MODE, FUSE>, &cps, X::WPB * 32, X::SMEM_BYTES, X::WPB,
		                                 S, dest, v, alpha, beta, w, fz, &used);
		if (rc || used) return rc;
	}
	// measured (profiles/r01b): the bulk-copy kernel wins for the fused variants, the register-staged
	// one for the plain sweep; the value-indexed stream always goes through the bulk-copy kernel when
	// the matrix is large enough (its 2-byte loads are a poor fit for register staging)
	// (the fused restriction is implemented by the value-indexed bulk-copy kernel and the register-staged one)
	if constexpr (FUSE != FUSE_RESTRICT_JACOBI || COMP) {
		if (!ctx->no_tma && (COMP || FUSE != FUSE_NONE || ctx->tma_min_slices_per_warp == 0 || ctx->tma_all)) {
			const int rc = launch_tma<BETAK, MODE, FUSE, COMP>(ctx, S, dest, v, alpha, beta, w, fz, &used);
			if (rc) return rc;
		}
	}
	if (!used)
		UG_LAUNCH(ctx, (spmv1_kernel<BETAK, MODE, FUSE, COMP>), grid, kThreads, 0, S, dest, v, alpha, beta, w, fz, ctx->guard);
	return UG4B200_OK;
}

template <int BETAK, int MODE, int FUSE>
int launch_mode(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, const double* v, double alpha, double beta,
                const double* w, int vblock, const Fuse& fz)
{
	const int grid = spmv_grid(ctx, A->num_slices);
	const Sell S = view(A);
	// fused interface push: only the stand-alone scalar smoothing kernels take it; any other path drops the arm
	ug4b200_interface* armed = ctx->armed_iface; const double* armed_vec = ctx->armed_vec;
	ctx->armed_iface = nullptr; ctx->armed_vec = nullptr;
	Fuse fzp = fz;
	if ((FUSE == FUSE_JACOBI || FUSE == FUSE_RESTRICT_JACOBI) && armed && armed_vec == fz.st_out && A->block == 1 && vblock == 1 &&
	    !ug_batchable(ctx, A->nrows) && (FUSE != FUSE_JACOBI || (fz.flags & UG4B200_SMOOTH_JACOBI)))
		fzp.push = ug_iface_push_begin(ctx, armed, armed_vec);
	if (FUSE != FUSE_DOT && A->block == 1 && vblock == 1 && ug_batchable(ctx, A->nrows)) {
		// small operand: record instead of launching (batch.cu)
		if (A->nrows == 0) return UG4B200_OK;
		UgBatchOp o{};
		o.kind = UG_OP_SPMV; o.sub = MODE | (FUSE << 4); o.flags = fz.flags;
		o.comp = (A->comp && !ctx->no_comp) ? 1 : 0; o.vshift = A->vshift; o.n = A->nrows;
		o.slice_ptr = A->slice_ptr; o.rowlen = A->rowlen; o.cols = A->cols; o.vals = A->vals;
		o.vc = A->vc; o.colbase = A->colbase; o.dict = A->dict;
		o.dest = dest; o.v = v; o.w = w; o.diaginv = fz.diaginv; o.st_out = fz.st_out; o.sc = fz.sc;
		o.alpha = alpha; o.beta = beta;
		return ug_batch_push(ctx, o);
	}
	if (A->block == 1 && vblock == 1) {
		if (A->comp && !ctx->no_comp) return launch_scalar<BETAK, MODE, FUSE, true>(ctx, A, dest, v, alpha, beta, w, fzp);
		return launch_scalar<BETAK, MODE, FUSE, false>(ctx, A, dest, v, alpha, beta, w, fzp);
	} else if (A->block == 1) {
		if (FUSE != FUSE_NONE) return ug4b200_fail(ctx, UG4B200_ERR_ARG, "fused SpMV needs matrix block == vector block");
		if (vblock == 2) { UG_LAUNCH(ctx, (spmv1xV_kernel<2, BETAK, MODE>), grid, kThreads, 0, S, dest, v, alpha, beta, w, ctx->guard); }
		else if (vblock == 3) { UG_LAUNCH(ctx, (spmv1xV_kernel<3, BETAK, MODE>), grid, kThreads, 0, S, dest, v, alpha, beta, w, ctx->guard); }
		else return ug4b200_fail(ctx, UG4B200_ERR_ARG, "vector block size must be 1, 2 or 3");
	} else if (FUSE == FUSE_RESTRICT_JACOBI) {
		return ug4b200_fail(ctx, UG4B200_ERR_ARG, "restrict_jacobi_fused is implemented for scalar vectors only");
	} else if (A->block == vblock) {
		if (vblock == 2) { UG_LAUNCH(ctx, (spmvB_kernel<2, BETAK, MODE, FUSE>), grid, kThreads, 0, S, dest, v, alpha, beta, w, fz, ctx->guard); }
		else if (vblock == 3) { UG_LAUNCH(ctx, (spmvB_kernel<3, BETAK, MODE, FUSE>), grid, kThreads, 0, S, dest, v, alpha, beta, w, fz, ctx->guard); }
		else return ug4b200_fail(ctx, UG4B200_ERR_ARG, "matrix block size must be 1, 2 or 3");
	} else return ug4b200_fail(ctx, UG4B200_ERR_ARG, "matrix / vector block size mismatch");
	return UG4B200_OK;
}

template <int MODE>
int launch_beta(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, const double* v, double alpha, double beta,
                const double* w, int vblock)
{
	Fuse fz{};
	if (beta == 1.0) return launch_mode<1, MODE, FUSE_NONE>(ctx, A, dest, v, alpha, beta, w, vblock, fz);
	if (beta == -1.0) return launch_mode<-1, MODE, FUSE_NONE>(ctx, A, dest, v, alpha, beta, w, vblock, fz);
	return launch_mode<0, MODE, FUSE_NONE>(ctx, A, dest, v, alpha, beta, w, vblock, fz);
}

__global__ void get_diag_kernel(Sell A, const int* __restrict__ diagpos, int B, double* diag)
{
	ug_pdl_sync();
	const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= A.nrows) return;
	const int BB = B * B;
	const int64_t s = row >> 5; const int lane = (int)(row & 31);
	const int dp = diagpos[row];
	for (int q = 0; q < BB; ++q) diag[row * BB + q] = 0.0;
	if (dp >= 0) {
		const double* ve = A.vals + (A.slice_ptr[s] + (int64_t)dp * 32) * BB + lane;
		for (int q = 0; q < BB; ++q) diag[row * BB + q] = ve[q * 32];
	}
}

__global__ void jacobi_invert_kernel(int64_t nrows, int B, double invdamp, int block_inverse, const double* diag,
                                     double* diaginv)
{
	ug_pdl_sync();
	const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= nrows) return;
	const int BB = B * B;
	double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	for (int q = 0; q < BB; ++q) m[q] = diag[row * BB + q];
	// jacobi.h:210-215: GetDiag (if !block) ; m *= 1./damp ; GetInverse(diagInv, m)
	if (!block_inverse && B > 1)
		for (int r = 0; r < B; ++r) for (int c = 0; c < B; ++c) if (r != c) m[r + B * c] = 0.0;
	for (int q = 0; q < BB; ++q) m[q] = m[q] * invdamp;
	double* inv = diaginv + row * BB;
#define MM(r, c) m[(r) + B * (c)]
#define II(r, c) inv[(r) + B * (c)]
	if (B == 1) inv[0] = 1.0 / m[0];
	else if (B == 2) {
		// GetInverse2, densematrix_inverse.h:96-108
		double invdet = MM(0,0) * MM(1,1) - MM(1,0) * MM(0,1);
		if (invdet == 0.0) return;
		invdet = 1.0 / invdet;
		II(0,0) = MM(1,1) * invdet; II(1,1) = MM(0,0) * invdet;
		II(0,1) = MM(0,1) * -invdet; II(1,0) = MM(1,0) * -invdet;
	} else {
		// GetDet3 / GetInverse3, densematrix_inverse.h:155-181
		double invdet = MM(0,0)*MM(1,1)*MM(2,2) + MM(0,1)*MM(1,2)*MM(2,0) + MM(0,2)*MM(1,0)*MM(2,1)
		              - MM(0,0)*MM(1,2)*MM(2,1) - MM(0,1)*MM(1,0)*MM(2,2) - MM(0,2)*MM(1,1)*MM(2,0);
		if (invdet == 0.0) return;
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
	}
#undef MM
#undef II
}

// ---- value dictionary, host side ----------------------------------------------------------------------------------
// The distinct values of a matrix as bit patterns (-0.0 and 0.0 stay distinct), numbered in the order a sequential scan
// of the CRS value array first meets them.  Open addressing on a multiplicative hash; at most 65536 entries.
// (The first version used std::unordered_map and a serial scan: 0.45 s of the 1.0 s solver:init at 129^3.)
class ValueDict {
public:
	static constexpr int kMax = 65536;
	ValueDict() { rehash(256); }
	int size() const { return (int)m_bits.size(); }
	uint64_t bits(int i) const { return m_bits[i]; }
	int find(uint64_t b) const
	{
		for (uint64_t i = hash(b);; i = (i + 1) & m_mask) {
			const int s = m_slot[i];
			if (s == 0) return -1;
			if (m_bits[s - 1] == b) return s - 1;
		}
	}
	// index of b, appended if new; -1 if the dictionary is full
	int insert(uint64_t b)
	{
		const int f = find(b);
		if (f >= 0) return f;
		if (size() >= kMax) return -1;
		if ((uint64_t)(size() + 1) * 2 > m_mask + 1) rehash((m_mask + 1) * 4);
		m_bits.push_back(b);
		place(b, size());
		return size() - 1;
	}
	void values(std::vector<double>& out) const
	{
		out.resize(m_bits.size());
		if (!m_bits.empty()) std::memcpy(out.data(), m_bits.data(), 8 * m_bits.size());
	}
private:
	uint64_t hash(uint64_t b) const { return ((b * 0x9E3779B97F4A7C15ull) >> 20) & m_mask; }
	void place(uint64_t b, int slot1)
	{
		uint64_t i = hash(b);
		while (m_slot[i] != 0) i = (i + 1) & m_mask;
		m_slot[i] = slot1;
	}
	void rehash(uint64_t n)
	{
		m_mask = n - 1;
		m_slot.assign(n, 0);
		for (int k = 0; k < size(); ++k) place(m_bits[k], k + 1);
	}
	std::vector<uint64_t> m_bits;   // in order of first occurrence
	std::vector<int> m_slot;        // index + 1, 0 = empty
	uint64_t m_mask = 0;
};

// look-ups of consecutive entries: neighbouring entries of a stencil matrix often repeat the value
struct DictCursor {
	const ValueDict& d; uint64_t last = 0; int idx = -1;
	explicit DictCursor(const ValueDict& dd) : d(dd) {}
	unsigned int operator()(double v)
	{
		uint64_t b; std::memcpy(&b, &v, 8);
		if (idx < 0 || b != last) { last = b; idx = d.find(b); }
		return (unsigned int)idx;
	}
};

inline uint64_t value_bits(double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }

// false: more than 65536 distinct values.  The scan runs over contiguous chunks in parallel; merging the chunks'
// finds in chunk order reproduces the numbering of the sequential scan exactly (a value's first occurrence lies in the
// first chunk that holds it, and inside a chunk the finds are in scan order).
bool build_value_dict(const double* vals, int64_t nnz, ValueDict& dict)
{
	int T = 1;
#ifdef _OPENMP
	if (nnz >= (1 << 16)) T = omp_get_max_threads();
#endif
	std::vector<std::vector<uint64_t>> found((size_t)T);
	bool ok = true;
#pragma omp parallel num_threads(T) reduction(&& : ok)
	{
		int t = 0, nt = 1;
#ifdef _OPENMP
		t = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
		const int64_t lo = nnz * t / nt, hi = nnz * (t + 1) / nt;
		ValueDict local;
		std::vector<uint64_t>& mine = found[(size_t)t];
		uint64_t last = 0; bool have = false;
		for (int64_t p = lo; p < hi; ++p) {
			const uint64_t b = value_bits(vals[p]);
			if (have && b == last) continue;
			last = b; have = true;
			if (local.find(b) >= 0) continue;
			if (local.insert(b) < 0) { ok = false; break; }
			mine.push_back(b);
		}
	}
	if (!ok) return false;
	for (int t = 0; t < T; ++t)
		for (const uint64_t b : found[(size_t)t])
			if (dict.insert(b) < 0) return false;
	return true;
}

// value-indexed words: (column - smallest column of the slice) << 16 | dictionary index << vshift; padding words 0
void encode_value_indexed(int64_t nrows, int64_t ns, int64_t pnnz, const int64_t* sp, const int64_t* rowptr, const int* cols,
                          const double* vals, const ValueDict& dict, const int* colbase, int vshift, std::vector<unsigned int>& out)
{
	out.assign((size_t)pnnz, 0u);
#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < ns; ++s) {
		const int64_t base = sp[s];
		DictCursor idx(dict);
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			for (int64_t p = rowptr[r], k = 0; p < rowptr[r + 1]; ++p, ++k)
				out[base + k * 32 + l] = ((unsigned int)(cols[p] - colbase[s]) << 16) | (idx(vals[p]) << vshift);
		}
	}
}

inline int value_indexed_shift(int ndict) { return ndict <= 1024 ? 3 : 0; }

// ---- x-staged stream, host side (shared by ug4b200_matrix_upload_crs and ug4b200_host_stream_plan) ----
// Per slice the sorted distinct columns are grouped into runs of consecutive columns (gaps of <= 2 merged, ends aligned
// to 16 bytes); the words then carry the position of their column in the concatenation of the runs.  Any banded numbering
// of a structured grid gives a handful of runs per slice (27-point operator, lexicographic: 9 runs of 34 columns).
struct XsPlan {
	bool ok = false;
	int rmax = 0;                  // run slots per slice actually needed (stride of `runs`)
	int max_doubles = 0;           // widest staged x segment
	int64_t doubles = 0;           // staged doubles summed over the slices
	std::vector<unsigned int> xw;  // [pnnz] position << 16 | dictionary index << 3
	std::vector<int4> hdr;         // [ns] {entry offset / 32, width, staged bytes, runs}
	std::vector<int2> runs;        // [ns * rmax] {first column (even), doubles (even) | position << 16}
};
void build_xs_plan(int64_t nrows, int64_t ns, int64_t pnnz, const int64_t* sp, const int64_t* rowptr, const int* cols,
                   const double* vals, const ValueDict& dict, XsPlan& out)
{
	const int RMAX = tma_xs_max_runs(), XCAP = tma_xs_max_doubles();
	std::vector<unsigned int>& hxw = out.xw; hxw.assign((size_t)pnnz, 0u);
	std::vector<int4>& hh = out.hdr; hh.assign((size_t)ns, make_int4(0, 0, 0, 0));
	std::vector<int2> hr((size_t)ns * RMAX, make_int2(0, 0));
	bool xok = true; int rmaxUsed = 1, maxDoubles = 0; int64_t xtotal = 0;
#pragma omp parallel for schedule(static) reduction(&& : xok) reduction(max : rmaxUsed, maxDoubles) reduction(+ : xtotal)
	for (int64_t s = 0; s < ns; ++s) {
		const int64_t base = sp[s];
		const int width = (int)((sp[s + 1] - base) >> 5);
		hh[s] = make_int4((int)(base >> 5), width, 0, 0);
		std::vector<int> cs;
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) cs.push_back(cols[p]);
		}
		if (cs.empty()) continue;
		std::sort(cs.begin(), cs.end());
		cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
		int ra[64], rb[64], rd[64]; int nr = 0, tot = 0;   // run [ra, rb) staged at double index rd
		bool fit = true;
		for (size_t i = 0; i < cs.size(); ++i) {
			const int c = cs[i];
			if (nr > 0 && c < rb[nr - 1] + 3) { rb[nr - 1] = (c + 2) & ~1; continue; }   // extend (gap <= 2: cheaper than a new copy)
			if (nr == RMAX) { fit = false; break; }
			ra[nr] = c & ~1; rb[nr] = (c + 2) & ~1; ++nr;
		}
		if (fit) for (int q = 0; q < nr; ++q) { rd[q] = tot; tot += rb[q] - ra[q]; }
		if (!fit || tot > XCAP) { xok = false; continue; }
		for (int q = 0; q < nr; ++q) hr[(size_t)s * RMAX + q] = make_int2(ra[q], (rb[q] - ra[q]) | (rd[q] << 16));
		hh[s].z = tot * 8; hh[s].w = nr;
		if (nr > rmaxUsed) rmaxUsed = nr;
		if (tot > maxDoubles) maxDoubles = tot;
		xtotal += tot;
		DictCursor idx(dict);
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			int q = 0;
			for (int64_t p = rowptr[r], k = 0; p < rowptr[r + 1]; ++p, ++k) {
				while (cols[p] >= rb[q]) ++q;     // columns ascend inside a row, runs ascend
				hxw[base + k * 32 + l] = ((unsigned int)(rd[q] + cols[p] - ra[q]) << 16) | (idx(vals[p]) << 3);
			}
		}
	}
	out.ok = xok; out.rmax = rmaxUsed; out.max_doubles = maxDoubles; out.doubles = xtotal;
	if (!xok) return;
	// keep only the run slots any slice uses (stride of the run table)
	out.runs.resize((size_t)ns * rmaxUsed);
	for (int64_t s2 = 0; s2 < ns; ++s2) for (int q = 0; q < rmaxUsed; ++q) out.runs[(size_t)s2 * rmaxUsed + q] = hr[(size_t)s2 * RMAX + q];
}

} // namespace

extern "C" {

int ug4b200_matrix_upload_crs(ug4b200_ctx* ctx, int block, int64_t nrows, int64_t ncols, const int64_t* rowptr,
                              const int* cols, const double* vals, int flags, ug4b200_matrix** out)
{
	UG_ARG(ctx, out != nullptr, "out is NULL");
	*out = nullptr;
	UG_ARG(ctx, block >= 1 && block <= 3, "block size must be 1, 2 or 3");
	UG_ARG(ctx, nrows >= 0 && ncols >= 0 && rowptr != nullptr, "bad dimensions");
	UG_ARG(ctx, ncols < 2147483647LL && nrows < 2147483647LL, "dimensions exceed int32 columns");
	const int BB = block * block;
	const int64_t nnz = rowptr[nrows];
	const int64_t ns = (nrows + 31) / 32;
	std::vector<int64_t> sp(ns + 1, 0);
	std::vector<int> rl(ns * 32 > 0 ? ns * 32 : 1, 0), dp(ns * 32 > 0 ? ns * 32 : 1, -1);
	int maxlen = 0;
	bool sorted = true, alldiag = (nrows == ncols);
#pragma omp parallel for schedule(static) reduction(max : maxlen) reduction(&& : sorted, alldiag)
	for (int64_t s = 0; s < ns; ++s) {
		int w = 0;
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			const int64_t len = rowptr[r + 1] - rowptr[r];
			rl[r] = (int)len;
			if (len > w) w = (int)len;
			bool hasdiag = false;
			for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
				if (p > rowptr[r] && cols[p] <= cols[p - 1]) sorted = false;
				if (cols[p] < 0 || cols[p] >= ncols) sorted = false;
				if (cols[p] == r) { dp[r] = (int)(p - rowptr[r]); hasdiag = true; }
			}
			if (!hasdiag) alldiag = false;
		}
		sp[s + 1] = (int64_t)w * 32;
		if (w > maxlen) maxlen = w;
	}
	UG_ARG(ctx, sorted, "columns must be sorted ascending inside every row and lie in [0, ncols)");
	for (int64_t s = 0; s < ns; ++s) sp[s + 1] += sp[s];
	const int64_t pnnz = sp[ns];
	std::vector<int> hc((size_t)(pnnz > 0 ? pnnz : 1), 0);
	std::vector<double> hv((size_t)(pnnz > 0 ? pnnz : 1) * BB, 0.0);
#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < ns; ++s) {
		const int64_t base = sp[s];
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			for (int64_t p = rowptr[r], k = 0; p < rowptr[r + 1]; ++p, ++k) {
				hc[base + k * 32 + l] = cols[p];
				for (int q = 0; q < BB; ++q) hv[(base + k * 32) * BB + (int64_t)q * 32 + l] = vals[p * BB + q];
			}
		}
	}
	ug4b200_matrix* A = new ug4b200_matrix;
	A->block = block; A->nrows = nrows; A->ncols = ncols; A->nnz = nnz; A->padded_nnz = pnnz; A->num_slices = ns;
	A->max_row_len = maxlen; A->has_all_diag = alldiag;
	auto up = [&](void** d, const void* h, size_t bytes) -> int {
		if (bytes == 0) bytes = 8;
		cudaError_t e = cudaMalloc(d, bytes);
		if (e != cudaSuccess) { cudaGetLastError(); return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "matrix upload: out of device memory"); }
		A->device_bytes += bytes;
		e = cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
		if (e != cudaSuccess) return ug4b200_fail(ctx, UG4B200_ERR_CUDA, cudaGetErrorString(e));
		return 0;
	};
	int rc = 0;
	if (!rc) rc = up((void**)&A->slice_ptr, sp.data(), sizeof(int64_t) * (ns + 1));
	if (!rc) rc = up((void**)&A->rowlen, rl.data(), sizeof(int) * rl.size());
	if (!rc) rc = up((void**)&A->diagpos, dp.data(), sizeof(int) * dp.size());
	if (!rc) rc = up((void**)&A->cols, hc.data(), sizeof(int) * hc.size());
	if (!rc) rc = up((void**)&A->vals, hv.data(), sizeof(double) * hv.size());
	// ---- dictionary-encoded copies of the entry stream (scalar matrices only) ----
	// With at most 65536 distinct values (as bit patterns: -0.0 and 0.0 stay distinct) an entry needs an index instead
	// of 8 bytes.  Two lossless 4-byte-per-entry streams are built on that dictionary:
	//   value-indexed (vc):  word = column offset from the slice's smallest column << 16 | index << vshift; needs every
	//                        slice's columns within 65535 of its smallest column (129^3 lexicographic: yes; 257^3: no)
	//   x-staged (xw):       word = position of the column in the slice's staged x segment << 16 | index << 3; needs
	//                        <= 256 values, rows of <= 27 entries and per slice <= 32 runs / 384 doubles of columns
	//                        (any banded numbering of a structured grid, whatever its size)
	// Otherwise the plain 12-byte stream is used.
	std::vector<unsigned int> hvc; std::vector<int> hcb; std::vector<double> hdict;
	if (!rc && block == 1 && !ctx->no_comp && !(flags & UG4B200_MAT_NO_COMPRESS) && pnnz > 0) {
		bool ok = true, window = true;
		hcb.assign((size_t)ns, 0);
#pragma omp parallel for schedule(static) reduction(&& : window)
		for (int64_t s = 0; s < ns; ++s) {
			int lo = 2147483647, hi = -1;
			for (int l = 0; l < 32; ++l) {
				const int64_t r = s * 32 + l;
				if (r >= nrows) break;
				if (rowptr[r + 1] > rowptr[r]) { lo = std::min(lo, cols[rowptr[r]]); hi = std::max(hi, cols[rowptr[r + 1] - 1]); }
			}
			if (hi < 0) lo = 0;
			hcb[s] = lo;
			if (hi >= 0 && (int64_t)hi - lo > 65535) window = false;
		}
		ValueDict dict;
		ok = build_value_dict(vals, nnz, dict);
		if (ok) {
			dict.values(hdict);
			if (hdict.empty()) hdict.push_back(0.0);
			if (!rc) rc = up((void**)&A->dict, hdict.data(), sizeof(double) * hdict.size());
			if (!rc) A->ndict = (int)hdict.size();
		}
		if (ok && window) {
			const int vshift = value_indexed_shift((int)hdict.size());
			encode_value_indexed(nrows, ns, pnnz, sp.data(), rowptr, cols, vals, dict, hcb.data(), vshift, hvc);
			if (!rc) rc = up((void**)&A->vc, hvc.data(), sizeof(unsigned int) * hvc.size());
			if (!rc) rc = up((void**)&A->colbase, hcb.data(), sizeof(int) * hcb.size());
			if (!rc) { A->comp = true; A->vshift = vshift; }
		}
		if (ok) {
			// the x-staged stream is the fall-back of the value-indexed one (same speed where both apply, measured):
			// built when the column window is too wide for 16 bits, or on request (UG4B200_XSTAGE=1, tests)
			const bool want_xs = !A->comp || ctx->force_xs;
			// ---- x-staged copy (spmv1_xs_kernel): per slice the sorted distinct columns are grouped into runs of
			// consecutive columns (gaps of <= 2 merged, ends aligned to 16 bytes); the words then carry the position
			// of their column in the concatenation of the runs.  Any banded numbering of a structured grid gives a
			// handful of runs per slice (27-point operator, lexicographic: 9 runs of 34 columns).
			if (!rc && want_xs && !ctx->no_xs && !(flags & UG4B200_MAT_NO_XSTAGE) && hdict.size() <= (size_t)tma_xs_max_dict() && maxlen <= tma_xs_max_width()) {
				XsPlan xp;
				build_xs_plan(nrows, ns, pnnz, sp.data(), rowptr, cols, vals, dict, xp);
				if (xp.ok) {
					if (!rc) rc = up((void**)&A->xw, xp.xw.data(), sizeof(unsigned int) * xp.xw.size());
					if (!rc) rc = up((void**)&A->xs_hdr, xp.hdr.data(), sizeof(int4) * xp.hdr.size());
					if (!rc) rc = up((void**)&A->xs_runs, xp.runs.data(), sizeof(int2) * xp.runs.size());
					if (!rc) { A->xs = true; A->xs_rmax = xp.rmax; A->xs_doubles = xp.doubles; }
					if (!rc) { const cudaError_t e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) rc = ug4b200_fail(ctx, UG4B200_ERR_CUDA, cudaGetErrorString(e)); }
				}
			}
		}
	}
	if (rc) { ug4b200_matrix_destroy(ctx, A); return rc; }
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host staging buffers die here
	*out = A;
	return UG4B200_OK;
}

int ug4b200_host_stream_plan(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals,
                             ug4b200_stream_plan* plan, unsigned int* xw, int* hdr, int* runs, double* dictOut)
{
	if (!plan || !rowptr || nrows < 0 || ncols < 0) return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_host_stream_plan: bad argument");
	const int64_t nnz = rowptr[nrows];
	if (nnz > 0 && (!cols || !vals)) return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_host_stream_plan: cols / vals are NULL");
	const int64_t ns = (nrows + 31) / 32;
	std::vector<int64_t> sp((size_t)ns + 1, 0);
	int maxlen = 0; int64_t window = 0;
	for (int64_t s = 0; s < ns; ++s) {
		int w = 0, lo = 2147483647, hi = -1;
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			const int len = (int)(rowptr[r + 1] - rowptr[r]);
			if (len > w) w = len;
			if (len > 0) { lo = std::min(lo, cols[rowptr[r]]); hi = std::max(hi, cols[rowptr[r + 1] - 1]); }
		}
		sp[s + 1] = sp[s] + (int64_t)w * 32;
		if (w > maxlen) maxlen = w;
		if (hi >= 0 && (int64_t)hi - lo > window) window = (int64_t)hi - lo;
	}
	*plan = ug4b200_stream_plan{};
	plan->num_slices = ns; plan->padded_nnz = sp[ns]; plan->max_row_len = maxlen; plan->max_column_window = window;
	ValueDict dict;
	std::vector<double> hdict;
	const bool ok = build_value_dict(vals, nnz, dict);
	if (ok) dict.values(hdict);
	plan->num_distinct_values = ok ? (int)hdict.size() : -1;
	if (!ok || sp[ns] == 0) return UG4B200_OK;
	plan->value_indexed = window <= 65535 ? 1 : 0;
	if (dictOut) std::memcpy(dictOut, hdict.data(), sizeof(double) * hdict.size());
	if (hdict.size() > (size_t)tma_xs_max_dict() || maxlen > tma_xs_max_width()) return UG4B200_OK;
	XsPlan xp;
	build_xs_plan(nrows, ns, sp[ns], sp.data(), rowptr, cols, vals, dict, xp);
	plan->x_staged = xp.ok ? 1 : 0; plan->x_staged_runs = xp.rmax; plan->x_staged_max_doubles = xp.max_doubles;
	plan->x_staged_doubles = xp.doubles;
	if (!xp.ok) return UG4B200_OK;
	if (xw) std::memcpy(xw, xp.xw.data(), sizeof(unsigned int) * xp.xw.size());
	if (hdr) std::memcpy(hdr, xp.hdr.data(), sizeof(int4) * xp.hdr.size());
	if (runs) std::memcpy(runs, xp.runs.data(), sizeof(int2) * xp.runs.size());
	return UG4B200_OK;
}

int ug4b200_host_value_indexed_stream(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int* cols, const double* vals,
                                      unsigned int* words, int* colbase, int* vshift)
{
	if (!rowptr || !vshift || nrows < 0 || ncols < 0) return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_host_value_indexed_stream: bad argument");
	*vshift = -1;
	const int64_t nnz = rowptr[nrows];
	if (nnz > 0 && (!cols || !vals)) return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_host_value_indexed_stream: cols / vals are NULL");
	const int64_t ns = (nrows + 31) / 32;
	std::vector<int64_t> sp((size_t)ns + 1, 0);
	std::vector<int> hcb((size_t)ns, 0);
	bool window = true;
	for (int64_t s = 0; s < ns; ++s) {
		int w = 0, lo = 2147483647, hi = -1;
		for (int l = 0; l < 32; ++l) {
			const int64_t r = s * 32 + l;
			if (r >= nrows) break;
			const int len = (int)(rowptr[r + 1] - rowptr[r]);
			if (len > w) w = len;
			if (len > 0) { lo = std::min(lo, cols[rowptr[r]]); hi = std::max(hi, cols[rowptr[r + 1] - 1]); }
		}
		sp[s + 1] = sp[s] + (int64_t)w * 32;
		if (hi < 0) lo = 0;
		hcb[s] = lo;
		if (hi >= 0 && (int64_t)hi - lo > 65535) window = false;
	}
	ValueDict dict;
	if (sp[ns] == 0 || !window || !build_value_dict(vals, nnz, dict)) return UG4B200_OK;
	const int sh = value_indexed_shift(std::max(dict.size(), 1));
	if (words) {
		std::vector<unsigned int> w;
		encode_value_indexed(nrows, ns, sp[ns], sp.data(), rowptr, cols, vals, dict, hcb.data(), sh, w);
		std::memcpy(words, w.data(), sizeof(unsigned int) * w.size());
	}
	if (colbase && ns > 0) std::memcpy(colbase, hcb.data(), sizeof(int) * hcb.size());
	*vshift = sh;
	return UG4B200_OK;
}

int ug4b200_matrix_destroy(ug4b200_ctx* ctx, ug4b200_matrix* A)
{
	if (!A) return UG4B200_OK;
	if (ctx) { ug_batch_flush(ctx); cudaStreamSynchronize(ctx->stream); }
	cudaFree(A->slice_ptr); cudaFree(A->rowlen); cudaFree(A->diagpos); cudaFree(A->cols); cudaFree(A->vals);
	cudaFree(A->vc); cudaFree(A->colbase); cudaFree(A->dict);
	cudaFree(A->xw); cudaFree(A->xs_hdr); cudaFree(A->xs_runs);
	delete A;
	return UG4B200_OK;
}

int ug4b200_matrix_get_info(const ug4b200_matrix* A, ug4b200_matrix_info* info)
{
	info->nrows = A->nrows; info->ncols = A->ncols; info->nnz = A->nnz; info->padded_nnz = A->padded_nnz;
	info->num_slices = A->num_slices; info->device_bytes = (int64_t)A->device_bytes; info->block = A->block;
	info->max_row_len = A->max_row_len;
	info->value_indexed = A->comp ? 1 : 0; info->num_distinct_values = A->ndict;
	info->x_staged = A->xs ? 1 : 0; info->x_staged_runs = A->xs_rmax; info->x_staged_doubles = A->xs_doubles;
	return UG4B200_OK;
}

int ug4b200_matrix_axpy(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, double alpha, const double* v,
                        double beta, const double* w, int vblock)
{
	UG_ARG(ctx, A && dest && w, "NULL argument");
	UG_ARG(ctx, dest != w, "dest must not alias w");
	if (alpha == 0.0) return launch_beta<MODE_ASSIGN>(ctx, A, dest, v, alpha, beta, w, vblock);
	UG_ARG(ctx, v != nullptr, "v is NULL with alpha != 0");
	if (v == dest && alpha == 1.0) return launch_beta<MODE_INPLACE>(ctx, A, dest, v, alpha, beta, w, vblock);
	// dest == v with alpha != 1: dest[i] *= alpha, then accumulate == general branch read of v[i]
	return launch_beta<MODE_GENERAL>(ctx, A, dest, v, alpha, beta, w, vblock);
}
int ug4b200_matrix_apply(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x, int vblock)
{ return ug4b200_matrix_axpy(ctx, A, y, 0.0, nullptr, 1.0, x, vblock); }
int ug4b200_matrix_matmul_minus(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x, int vblock)
{ return ug4b200_matrix_axpy(ctx, A, y, 1.0, y, -1.0, x, vblock); }
int ug4b200_matrix_apply_ignore_zero_rows(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* dest, double beta,
                                          const double* w, int vblock)
{
	UG_ARG(ctx, A && dest && w && dest != w, "bad argument");
	return launch_beta<MODE_ASSIGN_SKIP_EMPTY>(ctx, A, dest, nullptr, 0.0, beta, w, vblock);
}
int ug4b200_matrix_apply_dot_ds(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x, ug4b200_fin fin)
{
	UG_ARG(ctx, A && y && x && y != x, "bad argument");
	UG_ARG(ctx, A->nrows == A->ncols, "apply_dot needs a square matrix");
	Fuse fz{}; fz.partials = ctx->partials; fz.counter = ctx->counter; fz.fin = fin; fz.ar = ug_ar_none();
	return launch_mode<1, MODE_ASSIGN, FUSE_DOT>(ctx, A, y, nullptr, 0.0, 1.0, x, A->block, fz);
}
int ug4b200_matrix_apply_dot_allreduce_ds(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* y, const double* x,
                                          ug4b200_fin fin, double* scratch_dev)
{
	UG_ARG(ctx, A && y && x && y != x, "bad argument");
	UG_ARG(ctx, A->nrows == A->ncols, "apply_dot needs a square matrix");
	if (ctx->nranks <= 1) return ug4b200_matrix_apply_dot_ds(ctx, A, y, x, fin);
	if (ctx->p2p && ctx->p2p->nranks > 1) {
		Fuse fz{}; fz.partials = ctx->partials; fz.counter = ctx->counter; fz.fin = fin; fz.ar = ug_ar_of(ctx);
		return launch_mode<1, MODE_ASSIGN, FUSE_DOT>(ctx, A, y, nullptr, 0.0, 1.0, x, A->block, fz);
	}
	UG_ARG(ctx, scratch_dev != nullptr, "scratch_dev needed for the NCCL transport");
	ug4b200_fin st{UG4B200_FIN_STORE, scratch_dev, nullptr, nullptr, nullptr};
	int rc = ug4b200_matrix_apply_dot_ds(ctx, A, y, x, st);
	if (!rc) rc = ug4b200_allreduce_sum(ctx, scratch_dev, 1);
	if (!rc) rc = ug4b200_scalar_fin_ds(ctx, scratch_dev, fin);
	return rc;
}
int ug4b200_jacobi_smooth_fused_src(ug4b200_ctx* ctx, const ug4b200_matrix* A, const double* diaginv, double* sd,
                                    const double* sd_in, const double* st_in, double* st_out, double* sc, int flags)
{
	UG_ARG(ctx, A && sd && st_in, "NULL argument");
	UG_ARG(ctx, A->nrows == A->ncols, "square matrix needed");
	UG_ARG(ctx, !(flags & UG4B200_SMOOTH_JACOBI) || (diaginv && st_out), "JACOBI needs diaginv and st_out");
	UG_ARG(ctx, !(flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) || sc, "ADD_* needs sc");
	UG_ARG(ctx, st_in != st_out && sd != st_in && sc != st_in && (sd_in == nullptr || sd_in != st_out), "st_in must not alias an output");
	Fuse fz{}; fz.diaginv = diaginv; fz.st_out = st_out; fz.sc = sc; fz.flags = flags;
	if (sd_in == nullptr || sd_in == sd)
		return launch_mode<-1, MODE_INPLACE, FUSE_JACOBI>(ctx, A, sd, sd, 1.0, -1.0, st_in, A->block, fz);
	// sd = 1.0*sd_in - A*st_in: the defect is read from another vector (no copy beforehand)
	return launch_mode<-1, MODE_GENERAL, FUSE_JACOBI>(ctx, A, sd, sd_in, 1.0, -1.0, st_in, A->block, fz);
}
int ug4b200_jacobi_smooth_fused(ug4b200_ctx* ctx, const ug4b200_matrix* A, const double* diaginv, double* sd,
                                const double* st_in, double* st_out, double* sc, int flags)
{ return ug4b200_jacobi_smooth_fused_src(ctx, A, diaginv, sd, nullptr, st_in, st_out, sc, flags); }

int ug4b200_restrict_jacobi_fused(ug4b200_ctx* ctx, const ug4b200_matrix* R, const double* diaginv_coarse,
                                  double* sd_coarse, double beta, const double* sd_fine, double* st_coarse)
{
	UG_ARG(ctx, R && diaginv_coarse && sd_coarse && sd_fine && st_coarse, "NULL argument");
	UG_ARG(ctx, R->block == 1, "scalar transfer matrix needed");
	UG_ARG(ctx, sd_coarse != sd_fine && st_coarse != sd_fine && st_coarse != sd_coarse, "arguments must not alias");
	Fuse fz{}; fz.diaginv = diaginv_coarse; fz.st_out = st_coarse;
	if (beta == 1.0) return launch_mode<1, MODE_ASSIGN_SKIP_EMPTY, FUSE_RESTRICT_JACOBI>(ctx, R, sd_coarse, nullptr, 0.0, beta, sd_fine, 1, fz);
	return launch_mode<0, MODE_ASSIGN_SKIP_EMPTY, FUSE_RESTRICT_JACOBI>(ctx, R, sd_coarse, nullptr, 0.0, beta, sd_fine, 1, fz);
}

int ug4b200_jacobi_prepare(ug4b200_ctx* ctx, const ug4b200_matrix* A, double damp, int block_inverse, double* diaginv)
{
	UG_ARG(ctx, A && diaginv, "NULL argument");
	UG_ARG(ctx, A->nrows == A->ncols, "Square Matrix needed for Jacobi Iteration.");
	if (A->nrows == 0) return UG4B200_OK;
	// diaginv doubles as scratch for the extracted diagonal (each thread reads its block before writing it)
	int rc = ug4b200_matrix_get_diag(ctx, A, diaginv);
	if (rc) return rc;
	return ug4b200_jacobi_invert_diag(ctx, A->nrows, A->block, damp, block_inverse, diaginv, diaginv);
}

int ug4b200_matrix_get_diag(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* diag)
{
	UG_ARG(ctx, A && diag, "NULL argument");
	UG_ARG(ctx, A->nrows == A->ncols, "square matrix needed");
	if (A->nrows == 0) return UG4B200_OK;
	UG_LAUNCH(ctx, get_diag_kernel, (int)((A->nrows + 255) / 256), 256, 0, view(A), A->diagpos, A->block, diag);
	return UG4B200_OK;
}

int ug4b200_jacobi_invert_diag(ug4b200_ctx* ctx, int64_t nrows, int block, double damp, int block_inverse,
                               const double* diag, double* diaginv)
{
	UG_ARG(ctx, diag && diaginv && block >= 1 && block <= 3, "bad argument");
	if (nrows <= 0) return UG4B200_OK;
	UG_LAUNCH(ctx, jacobi_invert_kernel, (int)((nrows + 255) / 256), 256, 0, nrows, block, 1. / damp, block_inverse, diag, diaginv);
	return UG4B200_OK;
}

} // extern "C"
