// ctx.cu — context, memory and stream plumbing of the C ABI (include/ug4b200.h).
// Replaces CUDAManager (ugbase/lib_algebra/gpu_algebra/cuda/cuda_manager.{h,cpp}).
#include "common.cuh"
#include <cstdlib>

thread_local std::string g_ug4b200_err;

extern "C" {

int ug4b200_ctx_create(int device, void* stream, ug4b200_ctx** out)
{
	if (!out) return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_ctx_create: out is NULL");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return ug4b200_fail(nullptr, UG4B200_ERR_CUDA,
		                    std::string("ug4b200_ctx_create: no usable CUDA device (") +
		                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
		                    "); this library has no CPU fallback");
	if (device < 0 || device >= ndev)
		return ug4b200_fail(nullptr, UG4B200_ERR_ARG, "ug4b200_ctx_create: device ordinal out of range");
	ug4b200_ctx* ctx = new ug4b200_ctx;
	ctx->device = device;
	UG_CUDA(ctx, cudaSetDevice(device));
	cudaDeviceProp prop;
	UG_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
	ctx->num_sms = prop.multiProcessorCount;
	{ const char* e = getenv("UG4B200_NO_TMA"); ctx->no_tma = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_NO_COMPRESS"); ctx->no_comp = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_NO_XSTAGE"); ctx->no_xs = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_XSTAGE"); ctx->force_xs = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_PDL"); ctx->pdl = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_TMA_ALL"); ctx->tma_all = e && e[0] == '1'; }
	{ const char* e = getenv("UG4B200_TMA_MIN_SLICES"); if (e) ctx->tma_min_slices_per_warp = atoi(e); }
	{ const char* e = getenv("UG4B200_BATCH"); ctx->batch = !(e && e[0] == '0'); }
	{ const char* e = getenv("UG4B200_BATCH_MAX_ROWS"); if (e) ctx->batch_max_rows = atoll(e); }
	if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
	else { UG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
	UG_CUDA(ctx, cudaMalloc(&ctx->partials, sizeof(double) * kMaxReduceBlocks));
	UG_CUDA(ctx, cudaMalloc(&ctx->counter, sizeof(unsigned int)));
	UG_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
	UG_CUDA(ctx, cudaMalloc(&ctx->dev_scalar, sizeof(double) * 8));
	UG_CUDA(ctx, cudaMallocHost(&ctx->host_scalar, sizeof(double) * 8));
	{ const int rc = ug4b200_batch_enable(ctx, ctx->batch ? 1 : 0, -1); if (rc) return rc; }   // resolves the cluster size
	*out = ctx;
	return UG4B200_OK;
}

int ug4b200_ctx_destroy(ug4b200_ctx* ctx)
{
	if (!ctx) return UG4B200_OK;
	cudaSetDevice(ctx->device);
	ug_batch_flush(ctx);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->nccl) ug4b200_comm_destroy(ctx);
	if (ctx->p2p) ug4b200_p2p_window_destroy(ctx);
	cudaFree(ctx->partials); cudaFree(ctx->counter); cudaFree(ctx->dev_scalar);
	cudaFreeHost(ctx->host_scalar);
	if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
	return UG4B200_OK;
}

const char* ug4b200_last_error(const ug4b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_ug4b200_err.c_str(); }

int ug4b200_sync(ug4b200_ctx* ctx)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ug4b200_p2p_check(ctx);
}
void* ug4b200_stream(ug4b200_ctx* ctx) { ug_batch_flush(ctx); return (void*)ctx->stream; }
int ug4b200_launch_count(const ug4b200_ctx* ctx, int64_t* n) { *n = ctx->launches; return UG4B200_OK; }
int ug4b200_set_guard(ug4b200_ctx* ctx, const int* dev_flag) { UG_FLUSH(ctx); ctx->guard = dev_flag; return UG4B200_OK; }

struct ug4b200_graph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; int64_t kernels = 0; };

int ug4b200_graph_begin(ug4b200_ctx* ctx)
{
	if (ctx->capturing) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "graph capture already active");
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
	ctx->capturing = true; ctx->capture_start = ctx->launches;
	return UG4B200_OK;
}
int ug4b200_graph_end(ug4b200_ctx* ctx, ug4b200_graph** out)
{
	*out = nullptr;
	if (!ctx->capturing) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "no graph capture active");
	const int rcf = ug_batch_flush(ctx);   // recorded small operations belong to the graph
	ctx->capturing = false;
	ug4b200_graph* g = new ug4b200_graph;
	cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
	g->kernels = ctx->launches - ctx->capture_start;
	ctx->launches = ctx->capture_start; // captured launches did not run
	if (e != cudaSuccess) { delete g; return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
	if (rcf) { cudaGraphDestroy(g->graph); delete g; return rcf; }
	e = cudaGraphInstantiate(&g->exec, g->graph, 0);
	if (e != cudaSuccess) { cudaGraphDestroy(g->graph); delete g; return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
	*out = g;
	return UG4B200_OK;
}
int ug4b200_graph_launch(ug4b200_ctx* ctx, ug4b200_graph* g)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
	ctx->launches += g->kernels;
	return UG4B200_OK;
}
int ug4b200_graph_destroy(ug4b200_ctx* ctx, ug4b200_graph* g)
{
	if (!g) return UG4B200_OK;
	if (ctx) { ug_batch_flush(ctx); cudaStreamSynchronize(ctx->stream); }
	if (g->exec) cudaGraphExecDestroy(g->exec);
	if (g->graph) cudaGraphDestroy(g->graph);
	delete g;
	return UG4B200_OK;
}

int ug4b200_event_create(ug4b200_ctx* ctx, void** ev)
{
	cudaEvent_t e;
	UG_CUDA(ctx, cudaEventCreate(&e));
	*ev = (void*)e;
	return UG4B200_OK;
}
int ug4b200_event_record(ug4b200_ctx* ctx, void* ev) { UG_FLUSH(ctx); UG_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->stream)); return UG4B200_OK; }
int ug4b200_event_sync(ug4b200_ctx* ctx, void* ev) { UG_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)ev)); return UG4B200_OK; }
int ug4b200_event_elapsed_ms(ug4b200_ctx* ctx, void* a, void* b, float* ms)
{ UG_CUDA(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b)); return UG4B200_OK; }
int ug4b200_event_destroy(ug4b200_ctx* ctx, void* ev) { if (ev) UG_CUDA(ctx, cudaEventDestroy((cudaEvent_t)ev)); return UG4B200_OK; }

int ug4b200_alloc(ug4b200_ctx* ctx, size_t bytes, void** dptr)
{
	*dptr = nullptr;
	if (bytes == 0) bytes = 8;
	// every allocation is readable up to the next 16-byte boundary: the x-staging bulk copies of the SpMV kernels move
	// whole 16-byte units and may read the 8 bytes behind a vector of odd length
	bytes = (bytes + 15) & ~(size_t)15;
	cudaError_t e = cudaMalloc(dptr, bytes);
	if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "ug4b200_alloc: out of device memory"); }
	UG_CUDA(ctx, e);
	return UG4B200_OK;
}
int ug4b200_free(ug4b200_ctx* ctx, void* dptr)
{
	if (!dptr) return UG4B200_OK;
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	UG_CUDA(ctx, cudaFree(dptr));
	return UG4B200_OK;
}
int ug4b200_h2d(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return UG4B200_OK;
}
int ug4b200_d2h(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return UG4B200_OK;
}
int ug4b200_d2h_async(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return UG4B200_OK;
}
int ug4b200_d2d(ug4b200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return UG4B200_OK;
}
int ug4b200_memset(ug4b200_ctx* ctx, void* dst, int byte, size_t bytes)
{
	UG_FLUSH(ctx);
	UG_CUDA(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
	return UG4B200_OK;
}
int ug4b200_host_alloc(ug4b200_ctx* ctx, size_t bytes, void** hptr)
{
	UG_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 8));
	return UG4B200_OK;
}
int ug4b200_host_free(ug4b200_ctx* ctx, void* hptr)
{
	if (hptr) UG_CUDA(ctx, cudaFreeHost(hptr));
	return UG4B200_OK;
}

} // extern "C"
