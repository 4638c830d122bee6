// blas1.cu — Vector<T> arithmetic and fused Krylov vector kernels (fp64, HBM-bound).
//
// Reference semantics: ugbase/lib_algebra/common/operations_vec.h:49-175 (VecScaleAdd,
// element-wise, evaluated left to right), ugbase/lib_algebra/cpu_algebra/vector_impl.h
// :72-79 (dotprod), :323-329 (norm).  Element-wise kernels are bit-identical to the CPU
// loops (compiled with -fmad=false); reductions use a fixed two-stage tree.
//
// Roofline: pure streaming, 8 B per vector read/write; 128-bit accesses, grid sized
// to the SM count, no shared-memory staging needed (no reuse).
#include "../common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(const ug4b200_ctx* ctx, int64_t n, int per_thread = 4)
{
	int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
	const int64_t cap = (int64_t)ctx->num_sms * 8;
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (int)b;
}

// Generic element-wise driver: 2 doubles per thread per step (16-byte accesses) when
// all pointers are 16-byte aligned, scalar otherwise.
template <class F>
__global__ void __launch_bounds__(kThreads) ew_kernel(int64_t n, F f, const int* guard, bool vec2)
{
	if (ug_guarded(guard)) return;
	f.prepare();
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	if (vec2) {
		const int64_t n2 = n >> 1;
		for (int64_t i = tid; i < n2; i += stride) f.two(i);
		if (tid == 0 && (n & 1)) f.one(n - 1);
	} else {
		for (int64_t i = tid; i < n; i += stride) f.one(i);
	}
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

struct FSet {
	double* x; double v;
	__device__ void prepare() {}
	__device__ void one(int64_t i) { x[i] = v; }
	__device__ void two(int64_t i) { reinterpret_cast<double2*>(x)[i] = make_double2(v, v); }
};
struct FCopy {
	double* d; const double* s;
	__device__ void prepare() {}
	__device__ void one(int64_t i) { d[i] = s[i]; }
	__device__ void two(int64_t i) { reinterpret_cast<double2*>(d)[i] = reinterpret_cast<const double2*>(s)[i]; }
};
struct FScale {
	double* x; double a;
	__device__ void prepare() {}
	__device__ void one(int64_t i) { x[i] = x[i] * a; }
	__device__ void two(int64_t i)
	{ double2 v = reinterpret_cast<double2*>(x)[i]; v.x = v.x * a; v.y = v.y * a; reinterpret_cast<double2*>(x)[i] = v; }
};
template <int SIGN> struct FAdd {
	double* d; const double* s;
	__device__ void prepare() {}
	__device__ void one(int64_t i) { d[i] = SIGN > 0 ? d[i] + s[i] : d[i] - s[i]; }
	__device__ void two(int64_t i)
	{
		double2 a = reinterpret_cast<double2*>(d)[i]; const double2 b = reinterpret_cast<const double2*>(s)[i];
		a.x = SIGN > 0 ? a.x + b.x : a.x - b.x; a.y = SIGN > 0 ? a.y + b.y : a.y - b.y;
		reinterpret_cast<double2*>(d)[i] = a;
	}
};
struct FScaleAdd2 {
	double* d; ug4b200_coef c1; const double* v1; ug4b200_coef c2; const double* v2;
	double a1, a2;
	__device__ void prepare() { a1 = ug_coef(c1); a2 = ug_coef(c2); }
	__device__ void one(int64_t i) { d[i] = a1 * v1[i] + a2 * v2[i]; }
	__device__ void two(int64_t i)
	{
		const double2 x = reinterpret_cast<const double2*>(v1)[i], y = reinterpret_cast<const double2*>(v2)[i];
		reinterpret_cast<double2*>(d)[i] = make_double2(a1 * x.x + a2 * y.x, a1 * x.y + a2 * y.y);
	}
};
struct FScaleAdd3 {
	double* d; ug4b200_coef c1; const double* v1; ug4b200_coef c2; const double* v2; ug4b200_coef c3; const double* v3;
	double a1, a2, a3;
	__device__ void prepare() { a1 = ug_coef(c1); a2 = ug_coef(c2); a3 = ug_coef(c3); }
	__device__ void one(int64_t i) { d[i] = a1 * v1[i] + a2 * v2[i] + a3 * v3[i]; }
	__device__ void two(int64_t i)
	{
		const double2 x = reinterpret_cast<const double2*>(v1)[i], y = reinterpret_cast<const double2*>(v2)[i],
		              z = reinterpret_cast<const double2*>(v3)[i];
		reinterpret_cast<double2*>(d)[i] = make_double2(a1 * x.x + a2 * y.x + a3 * z.x, a1 * x.y + a2 * y.y + a3 * z.y);
	}
};

inline int push_ew(ug4b200_ctx* ctx, int sub, int64_t n, double* dest, const double* v, const double* w, double alpha, double beta,
                   const int* idx = nullptr)
{
	UgBatchOp o{};
	o.kind = UG_OP_EW; o.sub = sub; o.n = n; o.dest = dest; o.v = v; o.w = w; o.alpha = alpha; o.beta = beta; o.cols = idx;
	return ug_batch_push(ctx, o);
}

template <class F> int launch_ew(ug4b200_ctx* ctx, int64_t n, F f, bool aligned)
{
	if (n <= 0) return UG4B200_OK;
	UG_LAUNCH(ctx, ew_kernel<F>, grid_for(ctx, n), kThreads, 0, n, f, ctx->guard, aligned);
	return UG4B200_OK;
}

// ---- reductions -----------------------------------------------------------------
// MODE 0: sum a[i]*b[i];
// MODE 1: dest = a1*v1 + a2*v2, sum dest^2;
// MODE 2: CG update x += alpha p; r -= alpha q; sum r^2
template <int MODE>
__global__ void __launch_bounds__(kReduceThreads)
reduce_kernel(int64_t n, const double* a, const double* b, double* dest,
              ug4b200_coef c1, ug4b200_coef c2, double* x, const double* p, const double* alpha_dev,
              double* partials, unsigned int* counter, ug4b200_fin fin, const int* guard)
{
	if (ug_guarded(guard)) return;
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	double acc = 0.0;
	if (MODE == 0) {
		for (int64_t i = tid; i < n; i += stride) acc += a[i] * b[i];
	} else if (MODE == 1) {
		const double a1 = ug_coef(c1), a2 = ug_coef(c2);
		for (int64_t i = tid; i < n; i += stride) {
			const double v = a1 * a[i] + a2 * b[i];
			dest[i] = v;
			acc += v * v;
		}
	} else {
		const double al = *alpha_dev, nal = -al;
		for (int64_t i = tid; i < n; i += stride) {
			// VecScaleAdd(x, 1.0, x, alpha, p); VecScaleAdd(r, 1.0, r, -alpha, q)  (cg.h:187-193)
			x[i] = 1.0 * x[i] + al * p[i];
			const double v = 1.0 * dest[i] + nal * b[i];
			dest[i] = v;
			acc += v * v;
		}
	}
	ug_block_reduce_fin(acc, partials, counter, fin);
}

inline int reduce_grid(const ug4b200_ctx* ctx, int64_t n)
{
	int64_t b = (n + kReduceThreads * 8 - 1) / (kReduceThreads * 8);
	int64_t cap = (int64_t)ctx->num_sms * 8;
	if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (int)b;
}

__global__ void scalar_ratio_kernel(double* out, const double* a, const double* b, const double* c, const double* d,
                                    ug4b200_conv_state* conv, const int* guard)
{
	if (ug_guarded(guard)) return;
	// BiCGStab: beta = (rho/rhoOld) * (alpha/omega)   (bicgstab.h:227)
	const double n1 = a ? *a : 1.0, d1 = b ? *b : 1.0, n2 = c ? *c : 1.0, d2 = d ? *d : 1.0;
	// breakdown (bicgstab.h: "rhoOld == 0" / "omega == 0" -> return false): end the iteration before anything is updated
	if (conv && ((b && d1 == 0.0) || (d && d2 == 0.0))) { conv->status = 4; conv->done = 1; return; }
	double r = n1;
	if (b) r = n1 / d1;
	if (c || d) { double s = n2; if (d) s = n2 / d2; r = r * s; }
	*out = r;
}
__global__ void scalar_fin_kernel(const double* r, ug4b200_fin fin, const int* guard)
{
	if (ug_guarded(guard)) return;
	ug_apply_fin(*r, fin);
}
__global__ void conv_init_kernel(ug4b200_conv_state* s, int max_steps, double min_defect, double rel_reduction,
                                 double* history, int cap)
{
	ug_pdl_sync();
	s->initial_defect = 0; s->current_defect = 0; s->last_defect = 0;
	s->min_defect = min_defect; s->rel_reduction = rel_reduction;
	s->step = 0; s->max_steps = max_steps; s->done = 0; s->status = 0;
	s->history_cap = cap; s->pad_ = 0; s->history = history;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
gather_scatter_kernel(int64_t nidx, int block, double* dst, const double* src, const int* __restrict__ idx, const int* guard)
{
	if (ug_guarded(guard)) return;
	const int64_t total = nidx * block;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = t / block; const int q = (int)(t - i * block);
		const int64_t j = (int64_t)idx[i] * block + q;
		if (MODE == 0) dst[t] = src[j];          // gather
		else if (MODE == 1) dst[j] = src[t];     // scatter
		else dst[j] = dst[j] + src[t];           // scatter-add (indices unique)
	}
}

} // namespace

extern "C" {

int ug4b200_vec_set(ug4b200_ctx* ctx, int64_t n, double* x, double value)
{ if (n > 0 && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_SET, n, x, nullptr, nullptr, value, 0.0); return launch_ew(ctx, n, FSet{x, value}, al16(x)); }
int ug4b200_vec_copy(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src)
{ if (n > 0 && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_COPY, n, dst, src, nullptr, 0.0, 0.0); return launch_ew(ctx, n, FCopy{dst, src}, al16(dst) && al16(src)); }
int ug4b200_vec_scale(ug4b200_ctx* ctx, int64_t n, double* x, double alpha)
{ if (n > 0 && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_SCALE, n, x, nullptr, nullptr, alpha, 0.0); return launch_ew(ctx, n, FScale{x, alpha}, al16(x)); }
int ug4b200_vec_add(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src)
{ if (n > 0 && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_ADD, n, dst, src, nullptr, 0.0, 0.0); return launch_ew(ctx, n, FAdd<1>{dst, src}, al16(dst) && al16(src)); }
int ug4b200_vec_sub(ug4b200_ctx* ctx, int64_t n, double* dst, const double* src)
{ if (n > 0 && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_SUB, n, dst, src, nullptr, 0.0, 0.0); return launch_ew(ctx, n, FAdd<-1>{dst, src}, al16(dst) && al16(src)); }

int ug4b200_vec_scale_add2_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                              ug4b200_coef a2, const double* v2)
{
	if (n > 0 && !a1.dev && !a2.dev && ug_batchable(ctx, n)) return push_ew(ctx, UG_EW_SCALE_ADD2, n, dest, v1, v2, a1.host, a2.host);
	FScaleAdd2 f{dest, a1, v1, a2, v2, 0, 0};
	return launch_ew(ctx, n, f, al16(dest) && al16(v1) && al16(v2));
}
int ug4b200_vec_scale_add3_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                              ug4b200_coef a2, const double* v2, ug4b200_coef a3, const double* v3)
{
	FScaleAdd3 f{dest, a1, v1, a2, v2, a3, v3, 0, 0, 0};
	return launch_ew(ctx, n, f, al16(dest) && al16(v1) && al16(v2) && al16(v3));
}
int ug4b200_vec_scale_add2(ug4b200_ctx* ctx, int64_t n, double* dest, double a1, const double* v1, double a2,
                           const double* v2)
{ return ug4b200_vec_scale_add2_ds(ctx, n, dest, ug4b200_coef{nullptr, a1}, v1, ug4b200_coef{nullptr, a2}, v2); }
int ug4b200_vec_scale_add3(ug4b200_ctx* ctx, int64_t n, double* dest, double a1, const double* v1, double a2,
                           const double* v2, double a3, const double* v3)
{
	return ug4b200_vec_scale_add3_ds(ctx, n, dest, ug4b200_coef{nullptr, a1}, v1, ug4b200_coef{nullptr, a2}, v2,
	                                 ug4b200_coef{nullptr, a3}, v3);
}

int ug4b200_vec_dot_ds(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, ug4b200_fin fin)
{
	UG_LAUNCH(ctx, reduce_kernel<0>, reduce_grid(ctx, n), kReduceThreads, 0, n, a, b, nullptr, ug4b200_coef{nullptr, 0},
	          ug4b200_coef{nullptr, 0}, nullptr, nullptr, nullptr, ctx->partials, ctx->counter, fin, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_vec_scale_add2_norm_ds(ug4b200_ctx* ctx, int64_t n, double* dest, ug4b200_coef a1, const double* v1,
                                   ug4b200_coef a2, const double* v2, ug4b200_fin fin)
{
	UG_LAUNCH(ctx, reduce_kernel<1>, reduce_grid(ctx, n), kReduceThreads, 0, n, v1, v2, dest, a1, a2, nullptr, nullptr,
	          nullptr, ctx->partials, ctx->counter, fin, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_cg_update_ds(ug4b200_ctx* ctx, int64_t n, double* x, const double* p, double* r, const double* q,
                         const double* alpha_dev, ug4b200_fin fin)
{
	UG_LAUNCH(ctx, reduce_kernel<2>, reduce_grid(ctx, n), kReduceThreads, 0, n, nullptr, q, r, ug4b200_coef{nullptr, 0},
	          ug4b200_coef{nullptr, 0}, x, p, alpha_dev, ctx->partials, ctx->counter, fin, ctx->guard);
	return UG4B200_OK;
}

static int host_reduce(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, int op, double* host_out)
{
	const int* g = ctx->guard; ctx->guard = nullptr; // host-returning reductions are never skipped
	ug4b200_fin fin{op, ctx->dev_scalar, nullptr, nullptr, nullptr};
	int rc = ug4b200_vec_dot_ds(ctx, n, a, b, fin);
	ctx->guard = g;
	if (rc) return rc;
	UG_CUDA(ctx, cudaMemcpyAsync(ctx->host_scalar, ctx->dev_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	*host_out = ctx->host_scalar[0];
	return UG4B200_OK;
}
int ug4b200_vec_dot(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, double* host_out)
{ return host_reduce(ctx, n, a, b, UG4B200_FIN_STORE, host_out); }
int ug4b200_vec_norm(ug4b200_ctx* ctx, int64_t n, const double* a, double* host_out)
{ return host_reduce(ctx, n, a, a, UG4B200_FIN_SQRT, host_out); }

int ug4b200_scalar_ratio_ds(ug4b200_ctx* ctx, double* out, const double* a, const double* b, const double* c,
                            const double* d)
{
	UG_LAUNCH(ctx, scalar_ratio_kernel, 1, 1, 0, out, a, b, c, d, (ug4b200_conv_state*)nullptr, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_scalar_ratio_conv_ds(ug4b200_ctx* ctx, double* out, const double* a, const double* b, const double* c,
                                 const double* d, ug4b200_conv_state* conv)
{
	UG_LAUNCH(ctx, scalar_ratio_kernel, 1, 1, 0, out, a, b, c, d, conv, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_scalar_fin_ds(ug4b200_ctx* ctx, const double* r_dev, ug4b200_fin fin)
{
	UG_LAUNCH(ctx, scalar_fin_kernel, 1, 1, 0, r_dev, fin, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_conv_init(ug4b200_ctx* ctx, ug4b200_conv_state* dev_state, int max_steps, double min_defect,
                      double rel_reduction, double* dev_history, int history_cap)
{
	UG_LAUNCH(ctx, conv_init_kernel, 1, 1, 0, dev_state, max_steps, min_defect, rel_reduction, dev_history, history_cap);
	return UG4B200_OK;
}

int ug4b200_vec_gather(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const double* src, const int* idx)
{
	if (nidx <= 0) return UG4B200_OK;
	if (block == 1 && ug_batchable(ctx, nidx)) return push_ew(ctx, UG_EW_GATHER, nidx, dst, src, nullptr, 0.0, 0.0, idx);
	UG_LAUNCH(ctx, gather_scatter_kernel<0>, grid_for(ctx, nidx * block, 1), kThreads, 0, nidx, block, dst, src, idx, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_vec_scatter(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const int* idx, const double* src)
{
	if (nidx <= 0) return UG4B200_OK;
	UG_LAUNCH(ctx, gather_scatter_kernel<1>, grid_for(ctx, nidx * block, 1), kThreads, 0, nidx, block, dst, src, idx, ctx->guard);
	return UG4B200_OK;
}
int ug4b200_vec_scatter_add(ug4b200_ctx* ctx, int64_t nidx, int block, double* dst, const int* idx, const double* src)
{
	if (nidx <= 0) return UG4B200_OK;
	UG_LAUNCH(ctx, gather_scatter_kernel<2>, grid_for(ctx, nidx * block, 1), kThreads, 0, nidx, block, dst, src, idx, ctx->guard);
	return UG4B200_OK;
}

} // extern "C"
