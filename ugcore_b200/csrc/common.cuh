// common.cuh — context, error handling and launch helpers shared by the kernel TUs.
#pragma once
#include "../../include/ug4b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

// ---- batched small operations (batch.cu) ---------------------------------------------------
// Levels of a few thousand rows are launch- and latency-bound: a sweep takes ~1 us of dependent
// L2 round trips but costs 4-6 us as a kernel of its own.  Operations on such vectors/matrices are
// therefore not launched but RECORDED (in call order); the record is flushed as ONE kernel — a
// single thread-block cluster that executes the operations one after another with a cluster
// barrier in between — as soon as anything else is enqueued on the stream (every launch, copy,
// event, synchronisation and graph boundary flushes first, so stream order is preserved).
// The operation records travel as kernel parameters (no staging buffer: capture-safe).
enum { UG_OP_SPMV = 1, UG_OP_EW = 2, UG_OP_JACOBI = 3, UG_OP_LU = 4 };
enum { UG_EW_SET = 0, UG_EW_COPY, UG_EW_ADD, UG_EW_SUB, UG_EW_SCALE, UG_EW_GATHER, UG_EW_SCALE_ADD2 };
struct UgBatchOp {
	int kind, sub;                 // SPMV: sub = mode | fuse << 4 ; EW: sub = UG_EW_* ; JACOBI: sub = 1 adds into sc
	int flags, comp, vshift, pad_;
	int64_t n;                     // rows (SPMV, JACOBI, LU) / length (EW)
	const int64_t* slice_ptr; const int* rowlen; const int* cols; const double* vals;
	const unsigned int* vc; const int* colbase; const double* dict;
	double* dest; const double* v; const double* w; const double* diaginv; double* st_out; double* sc;
	double alpha, beta;
};
constexpr int kBatchMaxOps = 56;
/// dense LU base solve: up to this many unknowns the backward substitution keeps SolveLU's (serial) operation order and is
/// bit-identical to it; larger systems use the parallel column-oriented order (smoothers.cu: lu_apply_kernel)
constexpr int kLuExactMax = 128;
struct UgBatchParams { int nops; int pad_; UgBatchOp op[kBatchMaxOps]; };

struct ug4b200_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	const int* guard = nullptr;   // device flag: kernels return early when *guard != 0
	int64_t launches = 0;
	bool capturing = false;
	int64_t capture_start = 0;
	std::string err;
	int num_sms = 148;
	int tma_min_slices_per_warp = 2; // UG4B200_TMA_MIN_SLICES=0 forces the bulk-copy kernel (tests)
	bool tma_all = false;         // UG4B200_TMA_ALL=1: bulk-copy kernel also for unfused sweeps
	bool no_tma = false;          // UG4B200_NO_TMA=1: register-staged SpMV everywhere (A/B measurements)
	bool no_comp = false;         // UG4B200_NO_COMPRESS=1: never build / use the value-indexed entry stream
	bool no_xs = false;           // UG4B200_NO_XSTAGE=1: never build / use the x-staged stream (A/B measurements)
	bool force_xs = false;        // UG4B200_XSTAGE=1: build and use it also where the value-indexed stream applies (tests)
	bool pdl = false;             // UG4B200_PDL=1: programmatic dependent launch (next kernel's launch overlaps this one's tail)
	// batched small operations: UG4B200_BATCH=0 disables, UG4B200_BATCH_MAX_ROWS sets the size limit
	bool batch = true;
	int64_t batch_max_rows = 16384;
	int batch_cluster = -1;       // cluster size the device accepted (16, 8, ... ; 0: unavailable)
	int64_t batched_ops = 0;      // operations that ran inside batch kernels (statistics)
	std::vector<UgBatchOp> pending;
	// fused interface push (comm.cu): the next kernel that produces `armed_vec` as its smoother output
	// also stores the interface rows into the neighbours' peer windows
	struct ug4b200_interface* armed_iface = nullptr;
	const double* armed_vec = nullptr;
	// reduction workspace (stream-ordered reuse)
	double* partials = nullptr;   // [kMaxReduceBlocks]
	unsigned int* counter = nullptr;
	double* dev_scalar = nullptr; // scratch result slot for host-returning reductions
	double* host_scalar = nullptr; // pinned
	// NCCL (comm.cu)
	void* nccl = nullptr;         // ncclComm_t
	int nranks = 1, rank = 0;
	// peer windows (comm.cu): direct NVLink stores into the other ranks' memory
	struct ug4b200_p2p* p2p = nullptr;
};

// ---- peer windows: layout of every rank's window (identical on all ranks) ----------------
//   [0, 64 KiB)        table   Entry[kP2PMaxIfaces][kP2PMaxRanks]: where rank r writes for interface k
//   [64 KiB, +512)     all-reduce flags, one uint64 per source rank (vector all-reduce, p2p_allreduce_kernel)
//   [+512, +1024)      scalar all-reduce: tagged 16-byte slots [2 parities][kP2PMaxRanks] (ug_warp_allreduce)
//   [+2048, +256 KiB)  all-reduce slots  double[2 parities][kP2PMaxRanks][kP2PArMax]
//   [kP2PHeapOff, ..)  per-interface flags + receive regions (bump allocated)
constexpr int kP2PMaxRanks = 16;
constexpr int kP2PMaxIfaces = 64;
constexpr int kP2PArMax = 1024;
constexpr size_t kP2PArFlagOff = 65536;
constexpr size_t kP2PArLLOff = 65536 + 512;
constexpr size_t kP2PArDataOff = 65536 + 2048;
constexpr size_t kP2PHeapOff = kP2PArDataOff + (size_t)2 * kP2PMaxRanks * kP2PArMax * 8;
constexpr unsigned long long kP2PTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct ug4b200_p2p_entry { unsigned long long tag, recv_off, total, ptr_me, flag_off, pad_[3]; }; // 64 bytes

struct ug4b200_p2p {
	bool ipc = false;             // peers were opened with cudaIpcOpenMemHandle
	int nranks = 1, rank = 0;
	char* local = nullptr;
	size_t bytes = 0, bump = 0;
	char* peer[kP2PMaxRanks] = {};
	char** d_peer = nullptr;      // device copy of peer[]
	unsigned long long* d_epoch = nullptr; // all-reduce epoch (device, local)
	int* err_host = nullptr;      // mapped pinned: set by a kernel whose wait timed out
	int* err_dev = nullptr;
	cudaStream_t aux = nullptr;   // table publication / lookups, independent of the compute stream
	int next_iface = 0, live_ifaces = 0;
};

// what a reduction kernel needs to finish with an all-reduce over the peer windows
struct UgAr {
	char* const* peer; char* local; unsigned long long* epoch; int* err; int nranks, rank;
};
inline UgAr ug_ar_none() { return UgAr{nullptr, nullptr, nullptr, nullptr, 0, 0}; }
inline UgAr ug_ar_of(const ug4b200_ctx* ctx)
{
	const ug4b200_p2p* p = ctx->p2p;
	if (!p || p->nranks <= 1) return ug_ar_none();
	return UgAr{p->d_peer, p->local, p->d_epoch, p->err_dev, p->nranks, p->rank};
}

constexpr int kMaxReduceBlocks = 32768;
constexpr int kReduceThreads = 256;

extern thread_local std::string g_ug4b200_err;

inline int ug4b200_fail(ug4b200_ctx* ctx, int code, const std::string& msg)
{
	g_ug4b200_err = msg;
	if (ctx) ctx->err = msg;
	return code;
}

#define UG_CUDA(ctx, call)                                                                        \
	do {                                                                                          \
		cudaError_t e_ = (call);                                                                  \
		if (e_ != cudaSuccess)                                                                    \
			return ug4b200_fail(ctx, UG4B200_ERR_CUDA,                                            \
			                    std::string(#call) + ": " + cudaGetErrorString(e_));              \
	} while (0)

#define UG_ARG(ctx, cond, msg)                                                                    \
	do { if (!(cond)) return ug4b200_fail(ctx, UG4B200_ERR_ARG, std::string(__func__) + ": " + msg); } while (0)

// Every kernel launch goes through this so launches are counted and checked.  With ctx->pdl the
// launch carries the programmatic-stream-serialization attribute: the kernel may be scheduled
// while its predecessor drains; every kernel of this library therefore starts with ug_pdl_sync()
// (griddepcontrol.wait: returns once the predecessor has completed and flushed), which is a no-op
// for ordinary launches.
template <typename... KArgs, typename... Args>
inline cudaError_t ug_launch_ex(ug4b200_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
	cudaLaunchAttribute at[1];
	if (ctx->pdl) {
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
	}
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
int ug_batch_flush(ug4b200_ctx* ctx);   // batch.cu: launch the recorded small operations (no-op if none)
inline bool ug_batchable(const ug4b200_ctx* ctx, int64_t n) { return ctx->batch && ctx->batch_cluster != 0 && n <= ctx->batch_max_rows; }
int ug_batch_push(ug4b200_ctx* ctx, const UgBatchOp& op);
#define UG_FLUSH(ctx)                                                                             \
	do { if (!(ctx)->pending.empty()) { const int rcf_ = ug_batch_flush(ctx); if (rcf_) return rcf_; } } while (0)

#define UG_LAUNCH(ctx, kernel, grid, block, smem, ...)                                            \
	do {                                                                                          \
		UG_FLUSH(ctx);                                                                            \
		cudaError_t e_ = ug_launch_ex(ctx, kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__); \
		(ctx)->launches++;                                                                        \
		if (e_ == cudaSuccess) e_ = cudaGetLastError(); else cudaGetLastError();                  \
		if (e_ != cudaSuccess)                                                                    \
			return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string(#kernel) + ": " + cudaGetErrorString(e_)); \
	} while (0)

__device__ __forceinline__ void ug_pdl_sync()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// first statement of every kernel: dependency wait, then the "iterations queued past convergence" guard
__device__ __forceinline__ bool ug_guarded(const int* guard) { ug_pdl_sync(); return guard != nullptr && *guard != 0; }

__device__ __forceinline__ double ug_coef(const ug4b200_coef& c) { return c.dev ? c.host * (*c.dev) : c.host; }

// ---- device-side StdConvCheck (convergence_check_impl.h:139-169, 246-257) ----
__device__ __forceinline__ bool ug_valid_number(double v)
{
	if (v == 0.0) return true;
	return v >= 2.2250738585072014e-308 && v <= 1.7976931348623157e308 && v == v && v >= 0.0;
}
__device__ inline void ug_conv_check(ug4b200_conv_state* c)
{
	const double d = c->current_defect;
	const double red = d / c->initial_defect;
	int done = 0, status = 0;
	if (!ug_valid_number(d)) { done = 1; status = 3; }
	else if (c->step >= c->max_steps) { done = 1; status = 2; }
	else if (d < c->min_defect) { done = 1; }
	else if (red < c->rel_reduction) { done = 1; }
	if (done && status != 3) {
		// post(): success iff one of the two criteria holds
		if (d < c->min_defect || red < c->rel_reduction) status = 1;
		else status = 2;
	}
	if (done) { c->status = status; c->done = 1; }
}
__device__ inline void ug_apply_fin(double r, const ug4b200_fin& f)
{
	switch (f.op) {
		case UG4B200_FIN_STORE: *f.out = r; break;
		case UG4B200_FIN_A_DIV_R:
			if (f.out) *f.out = r;
			if (r == 0.0 && f.conv) { f.conv->status = 4; f.conv->done = 1; }
			*f.out2 = *f.a / r;
			break;
		case UG4B200_FIN_R_DIV_A: {
			const double av = *f.a;
			// with a convergence state attached a zero divisor is a breakdown (BiCGStab "tt == 0", bicgstab.h:344-352)
			if (av == 0.0 && f.conv) { f.conv->status = 4; f.conv->done = 1; }
			*f.out2 = r / av;
			if (f.out) *f.out = r;
			break;
		}
		case UG4B200_FIN_SQRT: *f.out = sqrt(r); break;
		case UG4B200_FIN_CONV_START: {
			ug4b200_conv_state* c = f.conv;
			const double d = sqrt(r);
			c->initial_defect = d; c->current_defect = d; c->last_defect = 0.0; c->step = 0;
			c->done = 0; c->status = 0;
			if (c->history && c->history_cap > 0) c->history[0] = d;
			if (f.out) *f.out = d;
			ug_conv_check(c);
			break;
		}
		case UG4B200_FIN_CONV_UPDATE: {
			ug4b200_conv_state* c = f.conv;
			const double d = sqrt(r);
			c->last_defect = c->current_defect; c->current_defect = d; c->step++;
			if (c->history && c->step < c->history_cap) c->history[c->step] = d;
			if (f.out) *f.out = d;
			ug_conv_check(c);
			break;
		}
	}
}

// ---- system-scope flags over NVLink ----------------------------------------------
__device__ __forceinline__ unsigned long long ug_ld_acquire_sys(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ug_st_release_sys(unsigned long long* p, unsigned long long v)
{ asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ double ug_ld_relaxed_sys(const double* p)
{ double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ug_st_relaxed_sys(double* p, double v)
{ asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ug_globaltimer()
{ unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// spin until *flag >= e (flags are monotone epochs); gives up after kP2PTimeoutNs and raises *err
__device__ inline bool ug_wait_flag(const unsigned long long* flag, unsigned long long e, int* err)
{
	unsigned long long t0 = 0; unsigned int spins = 0;
	while (ug_ld_acquire_sys(flag) < e) {
		if ((++spins & 0xfffu) == 0u) {
			const unsigned long long t = ug_globaltimer();
			if (t0 == 0) t0 = t;
			else if (t - t0 > kP2PTimeoutNs) { if (err) *(volatile int*)err = 1; return false; }
		}
	}
	return true;
}
// ---- flag-in-data slots ("LL"): every double travels as two 8-byte words, each carrying 32 data
// bits and the 32-bit tag of the exchange epoch.  An 8-byte store is single-copy atomic, so the
// receiver simply polls the slot until both tags match: no fence, no separate flag, one one-way
// NVLink trip instead of (store, ack, flag) — measured 8 us -> see DESIGN.md for an empty exchange.
__device__ __forceinline__ unsigned int ug_ll_tag(unsigned long long e) { return (unsigned int)(e % 0xffffffffull) + 1u; }   // never 0 (= cleared slot)
__device__ __forceinline__ void ug_ll_store(unsigned long long* slot, double v, unsigned int tag)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
	asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ inline double ug_ll_load(const unsigned long long* slot, unsigned int tag, int* err)
{
	unsigned long long w0, w1, t0 = 0; unsigned int spins = 0;
	for (;;) {
		asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
		if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag) break;
		if ((++spins & 0xfffu) == 0u) {
			const unsigned long long t = ug_globaltimer();
			if (t0 == 0) t0 = t;
			else if (t - t0 > kP2PTimeoutNs) { if (err) *(volatile int*)err = 1; break; }
		}
	}
	return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// ---- interface push fused into a producing kernel (comm.cu holds the host side) ------------
// The kernel that computes an additive vector whose interface rows must become consistent
// (AdditiveToConsistent, parallelization_util.h:159-191) stores those rows straight into the
// neighbours' peer windows (tagged slots, see above) while it streams the matrix.  The exchange that
// follows is then only "poll the neighbours' slots and add the copies".
struct UgPushNb {
	double* rbase;               // neighbour's receive region (in ITS window): 16-byte tagged slots
	int64_t rpar_stride;         // slots between its two parity buffers
	int64_t rptr;                // my first entry inside its region
	unsigned long long* rflag;   // (flag of the flag protocol; unused by the tagged-slot exchange)
};
struct UgPushDev {
	int nneigh, pad_;
	const unsigned int* rowmask;   // [ceil(nlocal/32)] bit l: row 32 s + l is an interface row
	const int* rowprefix;          // [ceil(nlocal/32)] interface rows before slice s
	const int* sptr;               // [nu + 1] sends of interface row u ...
	const int* scode;              // ... (position inside the neighbour's list << 5) | neighbour slot
	const int* scode1;             // [nu] the code of rows with exactly one send (faces), -1 otherwise: one look-up instead of three
	const UgPushNb* nb;
	unsigned long long* epoch; unsigned int* arrive;
};
__device__ __forceinline__ void ug_push_row(const UgPushDev* P, unsigned long long e, int64_t slice, int lane, unsigned int mask, double val)
{
	const int u = P->rowprefix[slice] + __popc(mask & ((1u << lane) - 1u));
	const int64_t par = (int64_t)(e & 1ull);
	const int c1 = P->scode1[u];
	if (c1 >= 0) {
		const UgPushNb nb = P->nb[c1 & 31];
		ug_ll_store(reinterpret_cast<unsigned long long*>(nb.rbase) + 2 * (par * nb.rpar_stride + nb.rptr + (c1 >> 5)), val, ug_ll_tag(e));
		return;
	}
	for (int p = P->sptr[u]; p < P->sptr[u + 1]; ++p) {
		const int code = P->scode[p];
		const UgPushNb nb = P->nb[code & 31];
		ug_ll_store(reinterpret_cast<unsigned long long*>(nb.rbase) + 2 * (par * nb.rpar_stride + nb.rptr + (code >> 5)), val, ug_ll_tag(e));
	}
}
// comm.cu: if `I` can take a fused push for `vec` return its device descriptor and remember that the
// next AdditiveToConsistent of `vec` only has to wait and add; nullptr otherwise
const UgPushDev* ug_iface_push_begin(ug4b200_ctx* ctx, struct ug4b200_interface* I, const double* vec);

// Sum of one double over all ranks, executed by warp 0 of ONE block per rank: every rank stores its
// value as a tagged slot [parity][rank] into every window, polls the slots of all ranks in its own
// window and adds them in ascending rank order, so the result is bitwise identical on every rank.
// `a` is taken from lane 0, the result is valid in lane 0.
__device__ inline double ug_warp_allreduce(double a, const UgAr& ar)
{
	const int lane = threadIdx.x & 31;
	a = __shfl_sync(0xffffffffu, a, 0);
	const unsigned long long e = *(volatile unsigned long long*)ar.epoch + 1ull;
	const int par = (int)(e & 1ull);
	const unsigned int tag = ug_ll_tag(e);
	__syncwarp();
	double x = 0.0;
	if (lane < ar.nranks) {
		// tagged slots: the value is its own arrival notice (one one-way NVLink trip, no flag, no fence)
		ug_ll_store(reinterpret_cast<unsigned long long*>(ar.peer[lane] + kP2PArLLOff) + 2 * (par * kP2PMaxRanks + ar.rank), a, tag);
		x = ug_ll_load(reinterpret_cast<const unsigned long long*>(ar.local + kP2PArLLOff) + 2 * (par * kP2PMaxRanks + lane), tag, ar.err);
	}
	double s = __shfl_sync(0xffffffffu, x, 0);
	for (int p = 1; p < ar.nranks; ++p) s = s + __shfl_sync(0xffffffffu, x, p);   // ascending rank order on every rank
	if (lane == 0) *(volatile unsigned long long*)ar.epoch = e;
	return s;
}

// ---- deterministic block reduction + last-block finalisation --------------------
// Every block reduces its value (fixed shuffle tree), writes partials[blockIdx.x];
// the last block to arrive sums the partials in a fixed order and applies `fin`.
// Result is independent of block scheduling: bit-reproducible for a given grid size.
__device__ __forceinline__ double ug_warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}
// must be called by all threads of the block; blockDim.x multiple of 32, <= 1024
__device__ inline void ug_block_reduce_fin(double v, double* partials, unsigned int* counter, const ug4b200_fin& fin,
                                           const UgAr& ar = UgAr{nullptr, nullptr, nullptr, nullptr, 0, 0})
{
	__shared__ double s_w[32];
	__shared__ bool s_last;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	v = ug_warp_sum(v);
	if (lane == 0) s_w[wid] = v;
	__syncthreads();
	if (wid == 0) {
		v = lane < nw ? s_w[lane] : 0.0;
		v = ug_warp_sum(v);
		if (lane == 0) {
			partials[blockIdx.x] = v;
			__threadfence();
			const unsigned int t = atomicAdd(counter, 1u);
			s_last = (t == gridDim.x - 1);
		}
	}
	__syncthreads();
	if (s_last) {
		__threadfence();
		double a = 0.0;
		for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += ((volatile double*)partials)[i];
		a = ug_warp_sum(a);
		if (lane == 0) s_w[wid] = a;
		__syncthreads();
		if (wid == 0) {
			a = lane < nw ? s_w[lane] : 0.0;
			a = ug_warp_sum(a);
			if (ar.nranks > 1) a = ug_warp_allreduce(a, ar); // local sum -> sum over ranks, same kernel
			if (lane == 0) { *counter = 0u; ug_apply_fin(a, fin); }
		}
	}
}

// streaming loads for data touched exactly once (matrix values / columns)
__device__ __forceinline__ double ug_ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ug_ld_stream(const int* p) { return __ldcs(p); }

struct ug4b200_matrix {
	int block = 1;
	int64_t nrows = 0, ncols = 0, nnz = 0, padded_nnz = 0, num_slices = 0;
	int max_row_len = 0;
	int64_t* slice_ptr = nullptr; // [num_slices+1] entry offsets (multiples of 32)
	int* rowlen = nullptr;        // [num_slices*32]
	int* diagpos = nullptr;       // [num_slices*32] position of the diagonal inside the row, -1 if absent
	int* cols = nullptr;          // [padded_nnz]
	double* vals = nullptr;       // [padded_nnz*block*block]; entry e, component q at (e/32*BB + q)*32 + e%32
	bool has_all_diag = false;
	size_t device_bytes = 0;
	// value-indexed copy of the entry stream (scalar matrices whose values repeat, e.g. every
	// uniformly refined level): one 32-bit word per entry = column offset from the slice's smallest
	// column in the high half, dictionary index << vshift in the low half; 4 B instead of 12 B per
	// entry, lossless (dict holds the exact fp64 bits).  vshift = 3 for dictionaries of <= 1024 values
	// (the low half is then the BYTE offset into the dictionary, and word >> 13 the byte offset of the
	// column: one instruction each in the kernel), 0 otherwise.  Padding words are 0 (safe to load).
	bool comp = false;
	int ndict = 0;
	int vshift = 0;
	unsigned int* vc = nullptr;       // [padded_nnz]
	int* colbase = nullptr;           // [num_slices]
	double* dict = nullptr;           // [ndict]
	// x-staged copy of the value-indexed stream (spmv_tma.cuh: spmv1_xs_kernel): the words address x by its position
	// in a per-slice staging buffer instead of by column; per slice a header {entry offset / 32, width, staged bytes,
	// runs} and xs_rmax run slots {first column (even), doubles (even) | staging position << 16}.  Built when the
	// dictionary has <= 256 values, rows have <= 27 entries and every slice's columns form <= xs_rmax runs of
	// <= 384 doubles in total (any banded ordering of a structured grid); the kernel then reads x from shared memory.
	bool xs = false;
	int xs_rmax = 0;
	int64_t xs_doubles = 0;
	unsigned int* xw = nullptr;       // [padded_nnz]
	int4* xs_hdr = nullptr;           // [num_slices]
	int2* xs_runs = nullptr;          // [num_slices * xs_rmax]
};
