// batch.cu — recorded small operations executed by ONE thread-block cluster.
//
// The coarse levels of the multigrid hierarchy (a few thousand rows) are launch- and
// latency-bound: a smoothing sweep there is ~1 us of dependent L2 round trips, but as a kernel
// of its own it costs 4-6 us (measured: levels 0..5 of the 129^3 hierarchy took 194 us of a
// 660 us CG iteration in ~32 launches).  The C-ABI entry points therefore do not launch
// operations on small operands; they record them (ug_batch_push), and the record is flushed as
// one kernel when anything else is enqueued on the stream (UG_FLUSH in every launch / copy /
// event / sync / graph boundary), which preserves stream order exactly.
//
// The kernel is one cluster of up to 16 CTAs x 384 threads (distributed over 16 SMs).  It runs
// the operations in recorded order; between two operations all CTAs meet at the hardware cluster
// barrier (barrier.cluster.arrive.release / wait.acquire: a few hundred ns instead of the
// ~2 us node-to-node latency of a graph launch), which also orders their global-memory accesses.
// Mutable vectors are read with ordinary loads (never ld.global.nc) so that the acquire makes
// the other CTAs' stores visible.
//
// Arithmetic: the same per-row operation sequence as the stand-alone kernels (spmv.cu,
// smoothers.cu, blas1.cu): one thread per row, ascending column order, separate multiply and
// add (-fmad=false), i.e. bit-identical to ugcore's CPU algebra
// (ugbase/lib_algebra/cpu_algebra/sparsematrix_impl.h:257-339, operator/preconditioner/jacobi.h:222-232,
// small_algebra/no_lapack/lu_decomp.h:160-195, common/operations_vec.h:49-175).
#include "common.cuh"

namespace {

enum { MODE_ASSIGN = 0, MODE_ASSIGN_SKIP_EMPTY = 1, MODE_INPLACE = 2, MODE_GENERAL = 3 };
enum { FUSE_NONE = 0, FUSE_DOT = 1, FUSE_JACOBI = 2, FUSE_RESTRICT_JACOBI = 3 };

constexpr int kBatchThreads = 384;
constexpr int kBatchUB = 14;        // entries per register batch of the plain stream: a 27-point row is two batches
constexpr int kBatchRowMax = 28;    // value-indexed stream: rows up to this length are fetched in ONE round trip
constexpr int kLuMax = 4096;

__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

// one row of an SpMV-family operation (all variants of spmv1_kernel, decided at run time)
// (the record is addressed as P.op[i].field in place: a reference to it would be copied to local memory)
#define o (P.op[opi])
__device__ __forceinline__ void spmv_row(const UgBatchParams& P, const int opi, int64_t row)
{
	const int mode = o.sub & 15, fuse = o.sub >> 4;
	const int64_t s = row >> 5; const int lane = (int)(row & 31);
	// matrix data is immutable: read-only path (stays in L1 across the cluster barriers)
	const int64_t base = __ldg(o.slice_ptr + s);
	const int len = __ldg(o.rowlen + row);
	const int cbase = o.comp ? __ldg(o.colbase + s) : 0;
	const double beta = o.beta;
	const double* w = o.w;
	double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
	if (mode == MODE_INPLACE) acc = o.dest[row];
	else if (mode == MODE_GENERAL) acc = o.alpha * o.v[row];
	if (fuse == FUSE_RESTRICT_JACOBI) { dinv = o.diaginv[row]; if (len == 0) own = o.dest[row]; }
	if (fuse == FUSE_JACOBI) {
		if (o.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
		if ((o.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(o.flags & UG4B200_SMOOTH_SC_ZERO)) scv = o.sc[row];
		if (o.flags & UG4B200_SMOOTH_JACOBI) dinv = o.diaginv[row];
	}
	const bool assign = (mode == MODE_ASSIGN || mode == MODE_ASSIGN_SKIP_EMPTY);
	if (o.comp && len <= kBatchRowMax) {
		// All entry words of the row in one round trip, then the x-gathers; the dictionary look-ups hit L1.
		// The dependent chain of a sweep is slice offset -> words -> x (first half) -> x (second half).
		const unsigned int* vcp = o.vc + base + lane;
		unsigned int e[kBatchRowMax];
#pragma unroll
		for (int u = 0; u < kBatchRowMax; ++u)
			if (u < len) e[u] = __ldg(vcp + (int64_t)u * 32);
		// x-gathers in two halves (register budget of 512 threads per CTA)
#pragma unroll
		for (int h = 0; h < kBatchRowMax; h += kBatchRowMax / 2) {
			double x[kBatchRowMax / 2];
#pragma unroll
			for (int u = 0; u < kBatchRowMax / 2; ++u)
				if (h + u < len) x[u] = w[cbase + (int)(e[h + u] >> 16)];
#pragma unroll
			for (int u = 0; u < kBatchRowMax / 2; ++u) {
				if (h + u < len) {
					const double t = (beta * __ldg(o.dict + ((e[h + u] & 0xffffu) >> o.vshift))) * x[u];
					if (assign && h + u == 0) acc = t;
					else acc = acc + t;
				}
			}
		}
	} else
	for (int k0 = 0; k0 < len; k0 += kBatchUB) {
		double a[kBatchUB], x[kBatchUB]; int c[kBatchUB];
		if (o.comp) {
			const unsigned int* vcp = o.vc + base + lane;
#pragma unroll
			for (int u = 0; u < kBatchUB; ++u)
				if (k0 + u < len) {
					const unsigned int e = __ldg(vcp + (int64_t)(k0 + u) * 32);
					c[u] = cbase + (int)(e >> 16);
					a[u] = __ldg(o.dict + ((e & 0xffffu) >> o.vshift));
				}
		} else {
			const double* vp = o.vals + base + lane; const int* cp = o.cols + base + lane;
#pragma unroll
			for (int u = 0; u < kBatchUB; ++u)
				if (k0 + u < len) { a[u] = __ldg(vp + (int64_t)(k0 + u) * 32); c[u] = __ldg(cp + (int64_t)(k0 + u) * 32); }
		}
#pragma unroll
		for (int u = 0; u < kBatchUB; ++u)
			if (k0 + u < len) x[u] = w[c[u]];
#pragma unroll
		for (int u = 0; u < kBatchUB; ++u) {
			if (k0 + u < len) {
				// beta * a is exact for beta = +-1, i.e. identical to the specialised kernels
				const double t = (beta * a[u]) * x[u];
				if (assign && k0 + u == 0) acc = t;
				else acc = acc + t;
			}
		}
	}
	if (fuse == FUSE_JACOBI) {
		o.dest[row] = acc;
		if (o.flags & UG4B200_SMOOTH_ADD_IN) scv = scv + own;
		if (o.flags & UG4B200_SMOOTH_JACOBI) {
			const double st = dinv * acc;
			o.st_out[row] = st;
			if (o.flags & UG4B200_SMOOTH_ADD_OUT) scv = scv + st;
		}
		if (o.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) o.sc[row] = scv;
	} else if (fuse == FUSE_RESTRICT_JACOBI) {
		double dv = own;
		if (len > 0) { o.dest[row] = acc; dv = acc; }
		o.st_out[row] = dinv * dv;
	} else {
		if (mode != MODE_ASSIGN_SKIP_EMPTY || len > 0) o.dest[row] = acc;
	}
}

__device__ void lu_solve_cta(const UgBatchParams& P, const int opi, double* sx)
{
	// same sequence as lu_apply_kernel (smoothers.cu): lu = o.vals, piv = o.cols, x = o.dest, b = o.w
	const int n = (int)o.n;
	const double* lu = o.vals; const int* piv = o.cols;
	for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = o.w[i];
	__syncthreads();
	if (threadIdx.x == 0)
		for (int i = 0; i < n; ++i) if (i < piv[i]) { const double t = sx[i]; sx[i] = sx[piv[i]]; sx[piv[i]] = t; }
	__syncthreads();
	for (int k = 0; k < n - 1; ++k) {
		const double xk = sx[k];
		for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) sx[i] = sx[i] - lu[(size_t)i * n + k] * xk;
		__syncthreads();
	}
	// (systems of more than kLuExactMax unknowns are not recorded: ug4b200_lu_apply launches lu_apply_large_kernel)
	if (threadIdx.x == 0) {
		for (int i = n - 1; i >= 0; --i) {
			double s = sx[i];
			for (int k = i + 1; k < n; ++k) s = s - lu[(size_t)i * n + k] * sx[k];
			sx[i] = s / lu[(size_t)i * n + i];
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < n; i += blockDim.x) o.dest[i] = sx[i];
}

__global__ void __launch_bounds__(kBatchThreads, 1)
batch_kernel(const __grid_constant__ UgBatchParams P, const int* guard)
{
	if (ug_guarded(guard)) return;
	__shared__ double s_lu[kLuMax];
	const int64_t tid = (int64_t)cluster_ctarank() * blockDim.x + threadIdx.x;
	const int64_t nthr = (int64_t)cluster_nctarank() * blockDim.x;
	for (int opi = 0; opi < P.nops; ++opi) {
		if (opi > 0) cluster_sync_all();
		switch (o.kind) {
			case UG_OP_SPMV:
				for (int64_t row = tid; row < o.n; row += nthr) spmv_row(P, opi, row);
				break;
			case UG_OP_JACOBI:
				for (int64_t r = tid; r < o.n; r += nthr) {
					const double st = o.diaginv[r] * o.w[r];   // MatMult(c[i], 1.0, diagInv[i], d[i])
					o.dest[r] = st;
					if (o.sub) o.sc[r] = o.sc[r] + st;
				}
				break;
			case UG_OP_EW:
				switch (o.sub) {
					case UG_EW_SET: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.alpha; break;
					case UG_EW_COPY: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.v[r]; break;
					case UG_EW_ADD: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.dest[r] + o.v[r]; break;
					case UG_EW_SUB: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.dest[r] - o.v[r]; break;
					case UG_EW_SCALE: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.dest[r] * o.alpha; break;
					case UG_EW_GATHER: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.v[o.cols[r]]; break;
					case UG_EW_SCALE_ADD2: for (int64_t r = tid; r < o.n; r += nthr) o.dest[r] = o.alpha * o.v[r] + o.beta * o.w[r]; break;
				}
				break;
			case UG_OP_LU:
				if (cluster_ctarank() == 0) lu_solve_cta(P, opi, s_lu);
				break;
		}
	}
}

#undef o

int batch_cluster_size(ug4b200_ctx* ctx)
{
	if (ctx->batch_cluster >= 0) return ctx->batch_cluster;
	ctx->batch_cluster = 0;
	if (cudaFuncSetAttribute(batch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
	const char* env = getenv("UG4B200_BATCH_CLUSTER");
	const int want = env ? atoi(env) : 16;
	for (int c = 16; c >= 1; c >>= 1) {
		if (c > want) continue;
		cudaLaunchConfig_t cfg{};
		cfg.gridDim = dim3(c); cfg.blockDim = dim3(kBatchThreads); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeClusterDimension;
		at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		int n = 0;
		if (cudaOccupancyMaxActiveClusters(&n, batch_kernel, &cfg) == cudaSuccess && n >= 1) { ctx->batch_cluster = c; break; }
		cudaGetLastError();
	}
	return ctx->batch_cluster;
}

} // namespace

int ug_batch_push(ug4b200_ctx* ctx, const UgBatchOp& op)
{
	if ((int)ctx->pending.size() >= kBatchMaxOps) { const int rc = ug_batch_flush(ctx); if (rc) return rc; }
	ctx->pending.push_back(op);
	return UG4B200_OK;
}

int ug_batch_flush(ug4b200_ctx* ctx)
{
	if (ctx->pending.empty()) return UG4B200_OK;
	const int c = batch_cluster_size(ctx);
	if (c <= 0) { ctx->pending.clear(); return ug4b200_fail(ctx, UG4B200_ERR_STATE, "batched operations recorded but no cluster launch available"); }
	static thread_local UgBatchParams P;
	P.nops = (int)ctx->pending.size(); P.pad_ = 0;
	for (int i = 0; i < P.nops; ++i) P.op[i] = ctx->pending[i];
	ctx->batched_ops += P.nops;
	ctx->pending.clear();   // before the launch: UG_LAUNCH-style helpers must not recurse into this flush
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(c); cfg.blockDim = dim3(kBatchThreads); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
	cudaLaunchAttribute at[2];
	at[0].id = cudaLaunchAttributeClusterDimension;
	at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	if (ctx->pdl) {
		at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[1].val.programmaticStreamSerializationAllowed = 1;
		cfg.numAttrs = 2;
	}
	cudaError_t e = cudaLaunchKernelEx(&cfg, batch_kernel, P, ctx->guard);
	ctx->launches++;
	if (e == cudaSuccess) e = cudaGetLastError(); else cudaGetLastError();
	if (e != cudaSuccess) return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string("batch_kernel: ") + cudaGetErrorString(e));
	return UG4B200_OK;
}

extern "C" {

int ug4b200_batch_flush(ug4b200_ctx* ctx) { return ug_batch_flush(ctx); }
int ug4b200_batch_enable(ug4b200_ctx* ctx, int on, int64_t max_rows)
{
	const int rc = ug_batch_flush(ctx);
	if (rc) return rc;
	ctx->batch = on != 0;
	if (max_rows >= 0) ctx->batch_max_rows = max_rows;
	if (ctx->batch && batch_cluster_size(ctx) <= 0) ctx->batch = false;   // no cluster launch on this device
	return UG4B200_OK;
}
int ug4b200_batch_stats(const ug4b200_ctx* ctx, int64_t* batched_ops, int* cluster_size)
{
	if (batched_ops) *batched_ops = ctx->batched_ops;
	if (cluster_size) *cluster_size = ctx->batch_cluster;
	return UG4B200_OK;
}

} // extern "C"
