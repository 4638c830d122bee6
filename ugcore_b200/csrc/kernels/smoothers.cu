// smoothers.cu — Jacobi step, multicolour Gauss-Seidel sweeps, dense-LU and
// single-CTA CG base solvers.
//
// Reference semantics (paths relative to /root/reference/ugbase):
//   lib_algebra/operator/preconditioner/jacobi.h:222-232           Jacobi::step
//   lib_algebra/algebra_common/core_smoothers.h:105-206            gs_step_LL / gs_step_UR / sgs_step
//   lib_algebra/small_algebra/double.h:157-161, small_matrix/densematrix_inverse.h:137-147, 229-245
//                                                                   InverseMatMult (scalar, 2x2, 3x3 Cramer)
//   lib_algebra/small_algebra/no_lapack/lu_decomp.h:160-195        SolveLU
//   lib_algebra/operator/linear_solver/cg.h:103-242                CG (base solver variant, no preconditioner)
// Multicolour GS: ugcore only has the lexicographic sweep; over a colour-sorted matrix
// that sweep is multicolour GS, each colour being one data-parallel launch here.
#include "../common.cuh"
#include <vector>

namespace {

struct Sell {
	const int64_t* slice_ptr; const int* rowlen; const int* diagpos; const int* cols; const double* vals;
	int64_t nrows, num_slices;
};
inline Sell view(const ug4b200_matrix* A)
{ return Sell{A->slice_ptr, A->rowlen, A->diagpos, A->cols, A->vals, A->nrows, A->num_slices}; }

template <int B, bool ADD>
__global__ void __launch_bounds__(256)
jacobi_step_kernel(int64_t n, const double* __restrict__ dinv, double* c, const double* d, double* sc, const int* guard)
{
	if (ug_guarded(guard)) return;
	constexpr int BB = B * B;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		double x[B], r[B];
#pragma unroll
		for (int t = 0; t < B; ++t) x[t] = d[i * B + t];
#pragma unroll
		for (int q = 0; q < B; ++q) {
			r[q] = dinv[i * BB + q] * x[0];
#pragma unroll
			for (int t = 1; t < B; ++t) r[q] = r[q] + dinv[i * BB + q + B * t] * x[t];
		}
#pragma unroll
		for (int q = 0; q < B; ++q) {
			c[i * B + q] = r[q];
			if (ADD) sc[i * B + q] = sc[i * B + q] + r[q];
		}
	}
}

// dest = beta * mat^{-1} * vec, same operation order as the reference
template <int B> __device__ __forceinline__ void inverse_mat_mult(double* dest, double beta, const double* m, const double* v)
{
#define MM(r, c) m[(r) + B * (c)]
	if (B == 1) { dest[0] = beta * v[0] / m[0]; }
	else if (B == 2) {
		const double det = MM(0,0) * MM(1,1) - MM(1,0) * MM(0,1);
		if (det == 0.0) return;
		const double d0 = beta * (MM(1,1) * v[0] - MM(0,1) * v[1]) / det;
		const double d1 = beta * (-MM(1,0) * v[0] + MM(0,0) * v[1]) / det;
		dest[0] = d0; dest[1] = d1;
	} else {
		const double det = MM(0,0)*MM(1,1)*MM(2,2) + MM(0,1)*MM(1,2)*MM(2,0) + MM(0,2)*MM(1,0)*MM(2,1)
		                 - MM(0,0)*MM(1,2)*MM(2,1) - MM(0,1)*MM(1,0)*MM(2,2) - MM(0,2)*MM(1,1)*MM(2,0);
		if (det == 0.0) return;
		const double d0 = (( MM(1,1)*MM(2,2) - MM(1,2)*MM(2,1)) * v[0] +
		                   (-MM(0,1)*MM(2,2) + MM(0,2)*MM(2,1)) * v[1] +
		                   ( MM(0,1)*MM(1,2) - MM(0,2)*MM(1,1)) * v[2]) * beta / det;
		const double d1 = ((-MM(1,0)*MM(2,2) + MM(1,2)*MM(2,0)) * v[0] +
		                   ( MM(0,0)*MM(2,2) - MM(0,2)*MM(2,0)) * v[1] +
		                   (-MM(0,0)*MM(1,2) + MM(0,2)*MM(1,0)) * v[2]) * beta / det;
		const double d2 = (( MM(1,0)*MM(2,1) - MM(1,1)*MM(2,0)) * v[0] +
		                   (-MM(0,0)*MM(2,1) + MM(0,1)*MM(2,0)) * v[1] +
		                   ( MM(0,0)*MM(1,1) - MM(0,1)*MM(1,0)) * v[2]) * beta / det;
		dest[0] = d0; dest[1] = d1; dest[2] = d2;
	}
#undef MM
}

// One colour of a Gauss-Seidel sweep.  DIR 0: forward (entries left of the diagonal),
// DIR 1: backward (entries right of it), DIR 2: c_i = A_ii * c_i (middle step of sgs_step).
template <int B, int DIR>
__global__ void __launch_bounds__(256)
gs_color_kernel(Sell A, int64_t r0, int64_t r1, double relax, double* c, const double* d, const int* guard)
{
	if (ug_guarded(guard)) return;
	constexpr int BB = B * B;
	const int64_t row = (r0 & ~(int64_t)31) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (row < r0 || row >= r1) return;
	const int64_t s = row >> 5; const int lane = (int)(row & 31);
	const int64_t base = A.slice_ptr[s];
	const int len = A.rowlen[row], dp = A.diagpos[row];
	const double* vp = A.vals + base * BB + lane;
	const int* cp = A.cols + base + lane;
	double sv[B];
#pragma unroll
	for (int t = 0; t < B; ++t) sv[t] = d[row * B + t];
	double aii[BB];
#pragma unroll
	for (int q = 0; q < BB; ++q) aii[q] = dp >= 0 ? vp[((int64_t)dp * BB + q) * 32] : 0.0;
	if (DIR == 2) {
		double r[B];
#pragma unroll
		for (int q = 0; q < B; ++q) {
			r[q] = aii[q] * sv[0];
#pragma unroll
			for (int t = 1; t < B; ++t) r[q] = r[q] + aii[q + B * t] * sv[t];
		}
#pragma unroll
		for (int q = 0; q < B; ++q) c[row * B + q] = r[q];
		return;
	}
	const int k0 = DIR == 0 ? 0 : dp + 1;
	const int k1 = DIR == 0 ? (dp >= 0 ? dp : len) : len;
	for (int k = k0; k < k1; ++k) {
		const int col = cp[(int64_t)k * 32];
		if (DIR == 0 && dp < 0 && col >= row) break;
		double x[B];
#pragma unroll
		for (int t = 0; t < B; ++t) x[t] = c[(int64_t)col * B + t];
		// MatMultAdd(s, 1.0, s, -1.0, a, c[col])
#pragma unroll
		for (int r = 0; r < B; ++r)
#pragma unroll
			for (int t = 0; t < B; ++t) sv[r] = sv[r] + (-vp[((int64_t)k * BB + r + B * t) * 32]) * x[t];
	}
	double out[B];
#pragma unroll
	for (int t = 0; t < B; ++t) out[t] = c[row * B + t];
	inverse_mat_mult<B>(out, relax, aii, sv);
#pragma unroll
	for (int t = 0; t < B; ++t) c[row * B + t] = out[t];
}

// ---- dense LU apply, one CTA ---------------------------------------------------------
__global__ void __launch_bounds__(256)
lu_apply_kernel(int n, const double* __restrict__ lu, const int* __restrict__ piv, double* x, const double* b,
                const int* guard)
{
	if (ug_guarded(guard)) return;
	extern __shared__ double sx[];
	for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = b[i];
	__syncthreads();
	if (threadIdx.x == 0)
		for (int i = 0; i < n; ++i) if (i < piv[i]) { const double t = sx[i]; sx[i] = sx[piv[i]]; sx[piv[i]] = t; }
	__syncthreads();
	// forward substitution, column oriented: row i receives its updates in ascending k,
	// the same order as the reference's row loop
	for (int k = 0; k < n - 1; ++k) {
		const double xk = sx[k];
		for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) sx[i] = sx[i] - lu[(size_t)i * n + k] * xk;
		__syncthreads();
	}
	// backward substitution.  SolveLU (no_lapack/lu_decomp.h:160-195) subtracts, for every row, in ASCENDING k — starting
	// with the unknown that was finished last — so a bit-identical evaluation is one serial chain of n^2/2 dependent
	// operations (measured: 28 ms per base solve at n = 1029, the 7^3 x 3 base grid of the partitioned elasticity run).
	// Up to kLuExactMax unknowns (every base grid of one to 2x2x2 cells, scalar) that chain is kept here and the result
	// equals SolveLU bit for bit; larger systems go to lu_apply_large_kernel (column-oriented, parallel).
	if (threadIdx.x == 0) {
		for (int i = n - 1; i >= 0; --i) {
			double s = sx[i];
			for (int k = i + 1; k < n; ++k) s = s - lu[(size_t)i * n + k] * sx[k];
			sx[i] = s / lu[(size_t)i * n + i];
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = sx[i];
}

// Dense LU apply for more than kLuExactMax unknowns: column-oriented forward and backward substitution (all rows
// receive the update of an unknown as soon as it is final).  Every step needs one column of the row-major factor — a
// strided global read whose ~1 us latency, paid 2 n times, was the whole cost (2.5 ms per apply at n = 1029).  The
// columns are therefore prefetched D steps ahead into a shared-memory ring with cp.async (no registers involved); a step
// is then two CTA barriers and n shared-memory updates.  Same terms per row as SolveLU, backward part in descending
// instead of ascending k: equal to the reference to round-off (tests/test_gpu_kernels.py).
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
	const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(1024)
lu_apply_large_kernel(int n, const double* __restrict__ lu, const int* __restrict__ piv, double* x, const double* b,
                      const int* guard)
{
	if (ug_guarded(guard)) return;
	extern __shared__ double sm[];
	double* sx = sm;                 // [n]
	double* ring = sm + n;           // [D][n]
	for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = b[i];
	__syncthreads();
	if (threadIdx.x == 0)
		for (int i = 0; i < n; ++i) if (i < piv[i]) { const double t = sx[i]; sx[i] = sx[piv[i]]; sx[piv[i]] = t; }
	// column k of the factor -> ring slot k % D (every thread copies the entries of its own rows)
	auto fetch = [&](int k) {
		if (k >= 0 && k < n) {
			double* dst = ring + (size_t)(k % D) * n;
			for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async8(dst + i, lu + (size_t)i * n + k);
		}
		cp_async_commit();           // (an empty group keeps the group count in step)
	};
	// ---- forward: for k = 0 .. n-2: rows i > k: sx[i] -= L[i][k] * sx[k]
	for (int d = 0; d < D; ++d) fetch(d);
	for (int k = 0; k < n - 1; ++k) {
		cp_async_wait<D - 1>();
		__syncthreads();             // column k has landed for every thread; sx[k] is final
		const double xk = sx[k];
		const double* col = ring + (size_t)(k % D) * n;
		for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) sx[i] = sx[i] - col[i] * xk;
		__syncthreads();             // slot k % D is free, sx updated
		fetch(k + D);
	}
	cp_async_wait<0>();
	__syncthreads();
	// ---- backward: for k = n-1 .. 0: sx[k] /= U[k][k]; rows i < k: sx[i] -= U[i][k] * sx[k]
	for (int d = 0; d < D; ++d) fetch(n - 1 - d);
	for (int k = n - 1; k >= 0; --k) {
		cp_async_wait<D - 1>();
		__syncthreads();
		const double* col = ring + (size_t)(k % D) * n;
		const double xk = sx[k] / col[k];      // every thread computes the same value; thread 0 stores it below
		__syncthreads();                       // all have read sx[k] before it is overwritten
		if (threadIdx.x == 0) sx[k] = xk;
		for (int i = threadIdx.x; i < k; i += blockDim.x) sx[i] = sx[i] - col[i] * xk;
		__syncthreads();
		fetch(k - D);
	}
	cp_async_wait<0>();
	__syncthreads();
	for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = sx[i];
}

// ---- single-CTA CG for tiny coarse systems ----------------------------------------------
__device__ double cta_sum(double v, double* s_w)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	v = ug_warp_sum(v);
	__syncthreads();
	if (lane == 0) s_w[wid] = v;
	__syncthreads();
	double r = 0.0;
	for (int i = 0; i < nw; ++i) r += s_w[i];
	return r;
}
__global__ void __launch_bounds__(1024)
coarse_cg_kernel(Sell A, int B, double* x, const double* b, double* work, int max_steps, double min_defect,
                 double rel_reduction, const int* guard)
{
	if (ug_guarded(guard)) return;
	__shared__ double s_w[32];
	const int64_t n = A.nrows * B;
	const int BB = B * B;
	double* r = work; double* p = work + n; double* q = work + 2 * n;
	double acc = 0.0;
	for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { x[i] = 0.0; const double v = b[i]; r[i] = v; p[i] = v; acc += v * v; }
	double rho = cta_sum(acc, s_w);
	const double norm0 = sqrt(rho);
	if (!(norm0 >= min_defect) || norm0 == 0.0) return;
	for (int it = 0; it < max_steps; ++it) {
		__syncthreads();
		// q = A p, lambda = (q, p)
		acc = 0.0;
		for (int64_t row = threadIdx.x; row < A.nrows; row += blockDim.x) {
			const int64_t s = row >> 5; const int lane = (int)(row & 31);
			const int64_t base = A.slice_ptr[s];
			const int len = A.rowlen[row];
			for (int rr = 0; rr < B; ++rr) {
				double a = 0.0;
				for (int k = 0; k < len; ++k) {
					const int col = A.cols[base + (int64_t)k * 32 + lane];
					for (int t = 0; t < B; ++t)
						a = a + A.vals[((base + (int64_t)k * 32) * BB) + (int64_t)(rr + B * t) * 32 + lane] * p[(int64_t)col * B + t];
				}
				q[row * B + rr] = a;
				acc += a * p[row * B + rr];
			}
		}
		const double lambda = cta_sum(acc, s_w);
		if (lambda == 0.0) return;
		const double alpha = rho / lambda;
		acc = 0.0;
		for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
			x[i] = x[i] + alpha * p[i];
			const double v = r[i] - alpha * q[i];
			r[i] = v; acc += v * v;
		}
		const double rho_new = cta_sum(acc, s_w);
		const double nr = sqrt(rho_new);
		if (nr < min_defect || nr / norm0 < rel_reduction) return;
		const double beta = rho_new / rho;
		for (int64_t i = threadIdx.x; i < n; i += blockDim.x) p[i] = beta * p[i] + r[i];
		rho = rho_new;
	}
}

template <int B>
int gs_launch(ug4b200_ctx* ctx, const ug4b200_matrix* A, int ncolors, const int64_t* cp, int kind, double relax,
              double* c, const double* d)
{
	const Sell S = view(A);
	auto grid_of = [](int64_t r0, int64_t r1) { return (int)(((r1 - (r0 & ~(int64_t)31)) + 255) / 256); };
	if (kind == 0 || kind == 2) {
		for (int k = 0; k < ncolors; ++k) {
			if (cp[k + 1] <= cp[k]) continue;
			UG_LAUNCH(ctx, (gs_color_kernel<B, 0>), grid_of(cp[k], cp[k + 1]), 256, 0, S, cp[k], cp[k + 1], relax, c, d, ctx->guard);
		}
	}
	if (kind == 2 && A->nrows > 0) {
		UG_LAUNCH(ctx, (gs_color_kernel<B, 2>), grid_of(0, A->nrows), 256, 0, S, (int64_t)0, A->nrows, relax, c, c, ctx->guard);
	}
	if (kind == 1 || kind == 2) {
		const double* dd = kind == 2 ? c : d;
		for (int k = ncolors - 1; k >= 0; --k) {
			if (cp[k + 1] <= cp[k]) continue;
			UG_LAUNCH(ctx, (gs_color_kernel<B, 1>), grid_of(cp[k], cp[k + 1]), 256, 0, S, cp[k], cp[k + 1], relax, c, dd, ctx->guard);
		}
	}
	return UG4B200_OK;
}

} // namespace

extern "C" {

int ug4b200_jacobi_step(ug4b200_ctx* ctx, int64_t n, int block, const double* diaginv, double* c, const double* d)
{
	if (n <= 0) return UG4B200_OK;
	if (block == 1 && ug_batchable(ctx, n)) {
		UgBatchOp o{}; o.kind = UG_OP_JACOBI; o.sub = 0; o.n = n; o.diaginv = diaginv; o.dest = c; o.w = d;
		return ug_batch_push(ctx, o);
	}
	int grid = (int)((n + 255) / 256); if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
	if (block == 1) { UG_LAUNCH(ctx, (jacobi_step_kernel<1, false>), grid, 256, 0, n, diaginv, c, d, nullptr, ctx->guard); }
	else if (block == 2) { UG_LAUNCH(ctx, (jacobi_step_kernel<2, false>), grid, 256, 0, n, diaginv, c, d, nullptr, ctx->guard); }
	else if (block == 3) { UG_LAUNCH(ctx, (jacobi_step_kernel<3, false>), grid, 256, 0, n, diaginv, c, d, nullptr, ctx->guard); }
	else return ug4b200_fail(ctx, UG4B200_ERR_ARG, "block size must be 1, 2 or 3");
	return UG4B200_OK;
}
int ug4b200_jacobi_step_add(ug4b200_ctx* ctx, int64_t n, int block, const double* diaginv, double* c, const double* d,
                            double* sc)
{
	if (n <= 0) return UG4B200_OK;
	if (block == 1 && ug_batchable(ctx, n)) {
		UgBatchOp o{}; o.kind = UG_OP_JACOBI; o.sub = 1; o.n = n; o.diaginv = diaginv; o.dest = c; o.w = d; o.sc = sc;
		return ug_batch_push(ctx, o);
	}
	int grid = (int)((n + 255) / 256); if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
	if (block == 1) { UG_LAUNCH(ctx, (jacobi_step_kernel<1, true>), grid, 256, 0, n, diaginv, c, d, sc, ctx->guard); }
	else if (block == 2) { UG_LAUNCH(ctx, (jacobi_step_kernel<2, true>), grid, 256, 0, n, diaginv, c, d, sc, ctx->guard); }
	else if (block == 3) { UG_LAUNCH(ctx, (jacobi_step_kernel<3, true>), grid, 256, 0, n, diaginv, c, d, sc, ctx->guard); }
	else return ug4b200_fail(ctx, UG4B200_ERR_ARG, "block size must be 1, 2 or 3");
	return UG4B200_OK;
}

// greedy first-fit colouring in row order over the stored pattern (host)
int ug4b200_color_greedy(int64_t nrows, const int64_t* rowptr, const int* cols, int* color, int* ncolors)
{
	int nc = 0;
	std::vector<int> mark;
	for (int64_t i = 0; i < nrows; ++i) color[i] = -1;
	for (int64_t i = 0; i < nrows; ++i) {
		mark.assign(nc + 1, 0);
		for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
			const int j = cols[p];
			if (j != i && j < nrows && color[j] >= 0) mark[color[j]] = 1;
		}
		int k = 0;
		while (k < nc && mark[k]) ++k;
		color[i] = k;
		if (k == nc) ++nc;
	}
	*ncolors = nc;
	return UG4B200_OK;
}
int ug4b200_color_check(int64_t nrows, const int64_t* rowptr, const int* cols, int ncolors, const int64_t* color_ptr)
{
	if (ncolors < 1 || color_ptr[0] != 0 || color_ptr[ncolors] != nrows) return UG4B200_ERR_ARG;
	for (int k = 0; k < ncolors; ++k) {
		if (color_ptr[k + 1] < color_ptr[k]) return UG4B200_ERR_ARG;
		for (int64_t i = color_ptr[k]; i < color_ptr[k + 1]; ++i)
			for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
				if (cols[p] != i && cols[p] >= color_ptr[k] && cols[p] < color_ptr[k + 1]) return UG4B200_ERR_ARG;
	}
	return UG4B200_OK;
}

int ug4b200_gs_step(ug4b200_ctx* ctx, const ug4b200_matrix* A, int ncolors, const int64_t* color_ptr_host, int kind,
                    double relax, double* c, const double* d)
{
	UG_ARG(ctx, A && c && d && color_ptr_host, "NULL argument");
	UG_ARG(ctx, A->nrows == A->ncols, "square matrix needed");
	UG_ARG(ctx, A->has_all_diag, "Gauss-Seidel: A has noninvertible diagonal (missing diagonal entry)");
	UG_ARG(ctx, kind >= 0 && kind <= 2, "kind must be 0, 1 or 2");
	UG_ARG(ctx, ncolors >= 1 && color_ptr_host[0] == 0 && color_ptr_host[ncolors] == A->nrows, "bad colour pointer");
	if (A->block == 1) return gs_launch<1>(ctx, A, ncolors, color_ptr_host, kind, relax, c, d);
	if (A->block == 2) return gs_launch<2>(ctx, A, ncolors, color_ptr_host, kind, relax, c, d);
	return gs_launch<3>(ctx, A, ncolors, color_ptr_host, kind, relax, c, d);
}

int ug4b200_lu_apply(ug4b200_ctx* ctx, int n, const double* lu_dev, const int* piv_dev, double* x, const double* b)
{
	if (n <= 0) return UG4B200_OK;
	UG_ARG(ctx, n <= 4096, "dense LU base solver limited to 4096 unknowns");
	if (n > kLuExactMax) {
		// large base system: its own kernel (shared-memory ring of prefetched factor columns; not recorded into the batch)
		const int D = n <= 1536 ? 8 : 4;
		const int smem = (int)sizeof(double) * n * (D + 1);
		static bool attr = false;
		if (!attr) {
			cudaFuncSetAttribute(lu_apply_large_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			cudaFuncSetAttribute(lu_apply_large_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			attr = true;
		}
		if (D == 8) { UG_LAUNCH(ctx, lu_apply_large_kernel<8>, 1, 1024, smem, n, lu_dev, piv_dev, x, b, ctx->guard); }
		else { UG_LAUNCH(ctx, lu_apply_large_kernel<4>, 1, 1024, smem, n, lu_dev, piv_dev, x, b, ctx->guard); }
		return UG4B200_OK;
	}
	if (ug_batchable(ctx, n)) {
		UgBatchOp o{}; o.kind = UG_OP_LU; o.n = n; o.vals = lu_dev; o.cols = piv_dev; o.dest = x; o.w = b;
		return ug_batch_push(ctx, o);
	}
	UG_LAUNCH(ctx, lu_apply_kernel, 1, 256, sizeof(double) * n, n, lu_dev, piv_dev, x, b, ctx->guard);
	return UG4B200_OK;
}

int ug4b200_coarse_cg(ug4b200_ctx* ctx, const ug4b200_matrix* A, double* x, const double* b, double* work4n,
                      int max_steps, double min_defect, double rel_reduction)
{
	UG_ARG(ctx, A && x && b && work4n, "NULL argument");
	UG_ARG(ctx, A->nrows == A->ncols, "square matrix needed");
	if (A->nrows == 0) return UG4B200_OK;
	int threads = 1024;
	if (A->nrows * A->block <= 256) threads = 256;
	UG_LAUNCH(ctx, coarse_cg_kernel, 1, threads, 0, view(A), A->block, x, b, work4n, max_steps, min_defect, rel_reduction,
	          ctx->guard);
	return UG4B200_OK;
}

} // extern "C"
