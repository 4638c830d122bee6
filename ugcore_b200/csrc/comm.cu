// comm.cu — horizontal-interface exchange and scalar all-reduce over NCCL (NVLink 5 /
// NVSwitch), one rank per GPU.
//
// Reference semantics (paths relative to /root/reference/ugbase):
//   pcl/pcl_interface_communicator_impl.hpp:408-739   pack -> MPI_Isend/Irecv per neighbour -> unpack
//   pcl/pcl_process_communicator.cpp:311-326          allreduce
//   lib_algebra/parallelization/parallelization_util.h:159-191  AdditiveToConsistent
//   lib_algebra/parallelization/parallelization_util.h:260-280  AdditiveToUnique
//   lib_algebra/parallelization/communication_policies.h:86-191 ComPol_VecAdd / VecCopy buffers
// Here: pack = gather kernel into one staging buffer, transport = grouped
// ncclSend/ncclRecv on the compute stream (no host sync), unpack = one kernel that sums
// every copy of an interface DoF in ascending rank order, so all copies end bitwise
// identical (ugcore's master/slave two-phase exchange gives the same sum up to order).
//
// NCCL is bound at run time (dlopen) so the library loads on hosts without NCCL and
// shares the libnccl.so.2 a host process (e.g. torch) already loaded.
//
// Peer windows (preferred transport inside one NVSwitch box): every rank owns a window of
// device memory that all other ranks map through CUDA IPC.  An interface exchange is then
// ONE kernel per rank: it gathers the interface values and stores them straight into the
// neighbours' windows over NVLink, raises a monotone epoch flag there (st.release.sys),
// waits for the neighbours' flags in its own window (ld.acquire.sys) and sums all copies
// in ascending rank order.  Receive regions are double buffered by epoch parity: a rank
// can be at most one epoch ahead of a neighbour, because finishing epoch e needs the
// neighbour's flag e, which it raises only after it has consumed epoch e-1.  The scalar
// all-reduce of the dot products uses the same mechanism inside the reduction kernel's
// last block (common.cuh: ug_warp_allreduce).  No host round trip, graph-capturable.
#include "common.cuh"
#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

typedef void* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
	void* h = nullptr;
	int (*GetUniqueId)(ncclUniqueId*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	bool ok = false;
	std::string why;
};
constexpr int kNcclDouble = 8, kNcclSum = 0;

NcclApi& nccl()
{
	static NcclApi api;
	static bool tried = false;
	if (tried) return api;
	tried = true;
	const char* names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char* n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.h) break; }
	if (!api.h) for (const char* n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
	if (!api.h) { api.why = "cannot dlopen libnccl.so.2"; return api; }
#define SYM(f) *(void**)(&api.f) = dlsym(api.h, "nccl" #f); if (!api.f) { api.why = "missing symbol nccl" #f; return api; }
	SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd)
	SYM(GetErrorString)
#undef SYM
	api.ok = true;
	return api;
}

#define UG_NCCL(ctx, call)                                                                         \
	do {                                                                                           \
		int r_ = (call);                                                                           \
		if (r_ != 0)                                                                               \
			return ug4b200_fail(ctx, UG4B200_ERR_NCCL, std::string(#call) + ": " + nccl().GetErrorString(r_)); \
	} while (0)

__global__ void pack_kernel(int64_t total, int block, const int* __restrict__ idx, const double* v, double* buf,
                            const int* guard)
{
	if (ug_guarded(guard)) return;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total * block; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t e = t / block; const int q = (int)(t - e * block);
		buf[t] = v[(int64_t)idx[e] * block + q];
	}
}
// u_idx[u]: local index; sources u_ptr[u]..u_ptr[u+1]: recv-buffer entry or -1 (= own value), ascending rank
__global__ void unpack_sum_kernel(int64_t nu, int block, const int* __restrict__ u_idx, const int* __restrict__ u_ptr,
                                  const int* __restrict__ u_src, const double* __restrict__ recv, double* v,
                                  const int* guard)
{
	if (ug_guarded(guard)) return;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nu * block; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t u = t / block; const int q = (int)(t - u * block);
		const int64_t li = (int64_t)u_idx[u] * block + q;
		double s = 0.0;
		for (int p = u_ptr[u]; p < u_ptr[u + 1]; ++p) {
			const int src = u_src[p];
			const double x = src < 0 ? v[li] : recv[(int64_t)src * block + q];
			s = (p == u_ptr[u]) ? x : s + x;
		}
		v[li] = s;
	}
}
__global__ void zero_idx_kernel(int64_t n, int block, const int* __restrict__ idx, double* v, const int* guard)
{
	if (ug_guarded(guard)) return;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * block; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t e = t / block; const int q = (int)(t - e * block);
		v[(int64_t)idx[e] * block + q] = 0.0;
	}
}
__global__ void __launch_bounds__(kReduceThreads)
dot_unique_kernel(int64_t nblocks, int block, const unsigned char* __restrict__ owned, const double* a, const double* b,
                  double* partials, unsigned int* counter, ug4b200_fin fin, const int* guard)
{
	if (ug_guarded(guard)) return;
	double acc = 0.0;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nblocks * block; t += (int64_t)gridDim.x * blockDim.x)
		if (owned[t / block]) acc += a[t] * b[t];
	ug_block_reduce_fin(acc, partials, counter, fin);
}


// ---- peer-window exchange ------------------------------------------------------------
typedef UgPushNb P2PNb;   // common.cuh: shared with the producing kernels that push interface rows themselves
struct P2PIfaceDev {
	int nneigh; int ell_w; int64_t total, nu;
	const int* idx;                    // [total] local index of send entry
	const int* ent_code;               // [total] (position inside the neighbour's list << 5) | neighbour slot
	const P2PNb* nb;
	const unsigned long long* lflag;   // [nneigh] in my window (flag protocol; unused by the tagged-slot exchange)
	const double* lrecv;               // my receive region (2 parities x total x 9 tagged 16-byte slots)
	int64_t lpar_stride;               // slots between the parity buffers
	const int* uidx;                   // [nu] local index of every interface DoF
	const int* uell;                   // [nu][ell_w] copies in ascending rank order: recv entry, -1 own value, -2 none
	const int* sptr; const int* scode; // [nu + 1] / [total] sends of interface DoF u: (position in the neighbour's list << 5) | slot
	const unsigned char* owned;        // [nlocal] 1 where this rank is the h-master
	unsigned long long* epoch; unsigned int* arrive; unsigned int* depart; int* err;
};

// One thread per interface DoF component: it sends its own value to every rank that shares the DoF
// (tagged slots), then polls the slots of the other copies and adds all copies in ascending rank
// order.  A DoF is read, sent and overwritten by the same thread, so no barrier is needed anywhere.
template <int W>
__device__ __forceinline__ void p2p_send_sum(const P2PIfaceDev& d, const P2PNb* s_nb, const unsigned long long* recv, int64_t par,
                                             unsigned int tag, double* v, int block, int unique, int pushed, int64_t tid, int64_t stride)
{
	for (int64_t t = tid; t < d.nu * block; t += stride) {
		const int64_t u = t / block; const int q = (int)(t - u * block);
		const int lidx = d.uidx[u];
		const int64_t li = (int64_t)lidx * block + q;
		int src[W]; double x[W];
#pragma unroll
		for (int k = 0; k < W; ++k) src[k] = d.uell[u * d.ell_w + k];
		const double own = v[li];
		if (!pushed) {
			for (int p = d.sptr[u]; p < d.sptr[u + 1]; ++p) {
				const int code = d.scode[p];
				const P2PNb& nb = s_nb[code & 31];
				ug_ll_store(reinterpret_cast<unsigned long long*>(nb.rbase) + 2 * (par * nb.rpar_stride + (nb.rptr + (code >> 5)) * block + q), own, tag);
			}
		}
		const bool keep = !unique || d.owned[lidx];
#pragma unroll
		for (int k = 0; k < W; ++k) x[k] = src[k] >= 0 ? ug_ll_load(recv + 2 * ((int64_t)src[k] * block + q), tag, d.err) : own;
		double s = x[0];
#pragma unroll
		for (int k = 1; k < W; ++k) if (src[k] != -2) s = s + x[k];
		// AdditiveToUnique (parallelization_util.h:260-280): only the h-master keeps the sum
		v[li] = keep ? s : 0.0;
	}
}

// AdditiveToConsistent / AdditiveToUnique in one kernel.  Grid <= #SMs so that all CTAs are
// co-resident: CTAs that poll for a neighbour's values must not keep CTAs that still have to send
// off the SMs (every CTA sends its share before it polls).
__global__ void __launch_bounds__(1024)
p2p_exchange_sum_kernel(P2PIfaceDev d, double* v, int block, int unique, int pushed, const int* guard)
{
	if (ug_guarded(guard)) return;
	__shared__ P2PNb s_nb[32];
	if (threadIdx.x < d.nneigh) s_nb[threadIdx.x] = d.nb[threadIdx.x];
	const unsigned long long e = *(volatile unsigned long long*)d.epoch + 1ull;
	const int64_t par = (int64_t)(e & 1ull);
	const unsigned int tag = ug_ll_tag(e);
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	__syncthreads();
	// send + poll + sum per interface DoF (pushed: the kernel that produced v has already sent, ug_push_row)
	const unsigned long long* recv = reinterpret_cast<const unsigned long long*>(d.lrecv) + 2 * par * d.lpar_stride;
	if (d.ell_w <= 2) p2p_send_sum<2>(d, s_nb, recv, par, tag, v, block, unique, pushed, tid, stride);
	else if (d.ell_w <= 4) p2p_send_sum<4>(d, s_nb, recv, par, tag, v, block, unique, pushed, tid, stride);
	else p2p_send_sum<8>(d, s_nb, recv, par, tag, v, block, unique, pushed, tid, stride);
	// the last CTA to leave publishes the epoch (every CTA read it on entry)
	__syncthreads();
	if (threadIdx.x == 0) {
		if (gridDim.x == 1) *(volatile unsigned long long*)d.epoch = e;
		else {
			__threadfence();
			if (atomicAdd(d.depart, 1u) == gridDim.x - 1) { *d.depart = 0u; *(volatile unsigned long long*)d.epoch = e; }
		}
	}
}

// in-place sum over ranks of n <= kP2PArMax doubles, one CTA per rank
__global__ void __launch_bounds__(1024)
p2p_allreduce_kernel(UgAr ar, double* x, int n, const int* guard)
{
	if (ug_guarded(guard)) return;
	const unsigned long long e = *(volatile unsigned long long*)ar.epoch + 1ull;
	const int par = (int)(e & 1ull);
	__syncthreads(); // every thread has read the epoch
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const double a = x[i];
		for (int p = 0; p < ar.nranks; ++p) {
			double* slot = reinterpret_cast<double*>(ar.peer[p] + kP2PArDataOff) + (size_t)(par * kP2PMaxRanks + ar.rank) * kP2PArMax + i;
			ug_st_relaxed_sys(slot, a);
		}
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < ar.nranks) {
		ug_st_release_sys(reinterpret_cast<unsigned long long*>(ar.peer[threadIdx.x] + kP2PArFlagOff) + ar.rank, e);
		ug_wait_flag(reinterpret_cast<const unsigned long long*>(ar.local + kP2PArFlagOff) + threadIdx.x, e, ar.err);
	}
	__syncthreads();
	const double* base = reinterpret_cast<const double*>(ar.local + kP2PArDataOff) + (size_t)par * kP2PMaxRanks * kP2PArMax;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		double s = 0.0;
		for (int p = 0; p < ar.nranks; ++p) {
			const double v = ug_ld_relaxed_sys(base + (size_t)p * kP2PArMax + i);
			s = (p == 0) ? v : s + v;
		}
		x[i] = s;
	}
	__syncthreads();
	if (threadIdx.x == 0) *(volatile unsigned long long*)ar.epoch = e;
}

__global__ void __launch_bounds__(kReduceThreads)
dot_allreduce_kernel(int64_t n, const double* a, const double* b, double* partials, unsigned int* counter,
                     ug4b200_fin fin, UgAr ar, const int* guard)
{
	if (ug_guarded(guard)) return;
	double acc = 0.0;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
		acc += a[t] * b[t];
	ug_block_reduce_fin(acc, partials, counter, fin, ar);
}


// ---- gathered level: local additive vectors -> global vector summed over all ranks ----------
struct GatherDst { double* rbase; unsigned long long* rflag; };   // my slot / my flag in rank r's window
struct GatherDev {
	int nranks, rank, block; int64_t nlocal, nglobal;               // in blocks
	const int* l2g; const GatherDst* dst;
	const unsigned long long* lflag; const double* lrecv;           // [nranks] / [2][nranks][nglobal*block]
	unsigned long long* epoch; unsigned int* arrive; unsigned int* depart; int* err;
};
// Every rank stores its local entries at their GLOBAL position of its own slot in every window
// (positions outside a rank's box are never written and stay 0 from creation), then each rank adds
// the nranks slots in ascending rank order: the gathered, consistent global vector of
// mg_solver_impl.hpp:2008-2068 on every rank, identical bits everywhere.
__global__ void __launch_bounds__(1024)
p2p_gather_sum_kernel(GatherDev d, double* gout, const double* lin, const int* guard)
{
	if (ug_guarded(guard)) return;
	__shared__ bool s_last;
	__shared__ GatherDst s_dst[kP2PMaxRanks];
	if (threadIdx.x < d.nranks) s_dst[threadIdx.x] = d.dst[threadIdx.x];
	const unsigned long long e = *(volatile unsigned long long*)d.epoch + 1ull;
	const int64_t par = (int64_t)(e & 1ull);
	const int64_t tot = d.nglobal * d.block, pstride = tot * d.nranks;
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
	__syncthreads();
	for (int64_t t = tid; t < d.nlocal * d.block; t += stride) {
		const int64_t i = t / d.block; const int q = (int)(t - i * d.block);
		const double val = lin[t];
		const int64_t g = (int64_t)d.l2g[i] * d.block + q;
		for (int r = 0; r < d.nranks; ++r) ug_st_relaxed_sys(s_dst[r].rbase + par * pstride + g, val);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		s_last = true;
		if (gridDim.x > 1) {
			__threadfence_system();
			s_last = (atomicAdd(d.arrive, 1u) == gridDim.x - 1);
			if (s_last) { *d.arrive = 0u; __threadfence_system(); }
		}
	}
	__syncthreads();
	if (s_last && threadIdx.x < d.nranks) ug_st_release_sys(s_dst[threadIdx.x].rflag, e);
	if (threadIdx.x < d.nranks) ug_wait_flag(d.lflag + threadIdx.x, e, d.err);
	__syncthreads();
	const double* base = d.lrecv + par * pstride;
	for (int64_t t = tid; t < tot; t += stride) {
		double s = 0.0;
		for (int r = 0; r < d.nranks; ++r) {
			const double x = ug_ld_relaxed_sys(base + (int64_t)r * tot + t);
			s = (r == 0) ? x : s + x;
		}
		gout[t] = s;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (gridDim.x == 1) *(volatile unsigned long long*)d.epoch = e;
		else {
			__threadfence();
			if (atomicAdd(d.depart, 1u) == gridDim.x - 1) { *d.depart = 0u; *(volatile unsigned long long*)d.epoch = e; }
		}
	}
}
__global__ void gather_scatter_local_kernel(int64_t nlocal, int block, const int* __restrict__ l2g, double* gout, const double* lin,
                                            const int* guard)
{
	if (ug_guarded(guard)) return;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nlocal * block; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = t / block; const int q = (int)(t - i * block);
		gout[(int64_t)l2g[i] * block + q] = lin[t];
	}
}

} // namespace

struct ug4b200_interface {
	int nneigh = 0;
	std::vector<int> rank;
	std::vector<int64_t> ptr;     // [nneigh+1]
	int64_t total = 0, nlocal = 0, nu = 0, nslave = 0;
	int* d_idx = nullptr;         // [total] local indices, neighbour-major
	int* d_uidx = nullptr; int* d_uptr = nullptr; int* d_usrc = nullptr;
	int* d_slave = nullptr;       // local indices whose h-master is another rank
	unsigned char* d_owned = nullptr; // [nlocal]
	double* sendbuf = nullptr; double* recvbuf = nullptr; // total*9 doubles (NCCL transport)
	// peer-window transport
	bool p2p = false, committed = false;
	int id = -1;                  // creation counter, identical on all ranks
	size_t win_off = 0;           // flags at win_off, receive region behind them
	size_t recv_off = 0;
	int* d_ent_code = nullptr; int* d_uell = nullptr; int ell_w = 2; P2PNb* d_nb = nullptr;
	unsigned long long* d_epoch = nullptr; unsigned int* d_counters = nullptr;
	// fused push (block size 1): row bitmap + prefix, sends grouped by interface row, device descriptor
	unsigned int* d_rowmask = nullptr; int* d_rowprefix = nullptr; int* d_sptr = nullptr; int* d_scode = nullptr; int* d_scode1 = nullptr;
	unsigned int* d_push_arrive = nullptr; UgPushDev* d_push = nullptr;
	const double* pushed_vec = nullptr;   // vector whose interface rows a producing kernel has pushed for the next exchange
};


// Allocate this interface's flags + receive region in the local window and publish, for every
// neighbour r, where r has to write (table entry [id][r]); neighbours look it up in commit.
template <class Up>
static int p2p_interface_setup(ug4b200_ctx* ctx, ug4b200_interface* I, Up& up, const std::vector<int>& uptr,
                               const std::vector<int>& usrc)
{
	ug4b200_p2p* P = ctx->p2p;
	// the one-kernel exchange keeps the neighbour table in shared memory and unrolls over <= 8 copies
	// (a 3-D box partition has <= 26 neighbours and <= 8 copies); anything else stays on NCCL
	int w = 1;
	for (size_t u = 0; u + 1 < uptr.size(); ++u) w = std::max(w, uptr[u + 1] - uptr[u]);
	int64_t maxrel = 0;
	for (int p = 0; p < I->nneigh; ++p) maxrel = std::max<int64_t>(maxrel, I->ptr[p + 1] - I->ptr[p]);
	if (I->nneigh > 32 || w > 8 || maxrel >= (1ll << 26)) return UG4B200_OK;
	I->ell_w = w <= 2 ? 2 : (w <= 4 ? 4 : 8);
	const size_t flag_bytes = (((size_t)I->nneigh * 8) + 255) / 256 * 256;
	const size_t recv_bytes = (((size_t)2 * I->total * 9 * 16) + 255) / 256 * 256;   // 2 parities x tagged 16-byte slots
	if (P->bump + flag_bytes + recv_bytes > P->bytes)
		return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "interface: peer window exhausted (UG4B200_P2P_WINDOW_MB)");
	I->id = P->next_iface++;
	I->win_off = P->bump; I->recv_off = P->bump + flag_bytes;
	P->bump += flag_bytes + recv_bytes; P->live_ifaces++;
	I->p2p = true;
	std::vector<int> ent_code(I->total);
	for (int p = 0; p < I->nneigh; ++p)
		for (int64_t e = I->ptr[p]; e < I->ptr[p + 1]; ++e) ent_code[e] = (int)(((e - I->ptr[p]) << 5) | p);
	const size_t nu = uptr.size() - 1;
	std::vector<int> uell(nu * I->ell_w > 0 ? nu * I->ell_w : 1, -2);
	for (size_t u = 0; u < nu; ++u)
		for (int k = uptr[u]; k < uptr[u + 1]; ++k) uell[u * I->ell_w + (k - uptr[u])] = usrc[k];
	int rc = 0;
	if (!rc) rc = up((void**)&I->d_ent_code, ent_code.data(), sizeof(int) * ent_code.size());
	if (!rc) rc = up((void**)&I->d_uell, uell.data(), sizeof(int) * uell.size());
	if (!rc) rc = up((void**)&I->d_nb, nullptr, sizeof(P2PNb) * I->nneigh);
	if (!rc) rc = up((void**)&I->d_epoch, nullptr, 8);
	if (!rc) rc = up((void**)&I->d_counters, nullptr, 8);
	if (rc) return rc;
	UG_CUDA(ctx, cudaMemsetAsync(I->d_epoch, 0, 8, ctx->stream));
	UG_CUDA(ctx, cudaMemsetAsync(I->d_counters, 0, 8, ctx->stream));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	// slots start untagged (tag 0; the region may be recycled); then publish
	UG_CUDA(ctx, cudaMemsetAsync(P->local + I->win_off, 0, flag_bytes + recv_bytes, P->aux));
	UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
	for (int p = 0; p < I->nneigh; ++p) {
		ug4b200_p2p_entry ent{};
		ent.tag = (unsigned long long)I->id + 1ull;
		ent.recv_off = I->recv_off; ent.total = (unsigned long long)I->total; ent.ptr_me = (unsigned long long)I->ptr[p];
		ent.flag_off = I->win_off + (size_t)p * 8;
		char* dst = P->local + sizeof(ug4b200_p2p_entry) * ((size_t)(I->id % kP2PMaxIfaces) * kP2PMaxRanks + I->rank[p]);
		UG_CUDA(ctx, cudaMemcpyAsync(dst, &ent, sizeof(ent), cudaMemcpyHostToDevice, P->aux));
		UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
	}
	return UG4B200_OK;
}

static int p2p_finish_open(ug4b200_ctx* ctx, ug4b200_p2p* P)
{
	UG_CUDA(ctx, cudaMalloc(&P->d_peer, sizeof(char*) * kP2PMaxRanks));
	UG_CUDA(ctx, cudaMemcpy(P->d_peer, P->peer, sizeof(char*) * kP2PMaxRanks, cudaMemcpyHostToDevice));
	UG_CUDA(ctx, cudaMalloc(&P->d_epoch, 8));
	UG_CUDA(ctx, cudaMemset(P->d_epoch, 0, 8));
	ctx->nranks = P->nranks; ctx->rank = P->rank;
	return UG4B200_OK;
}

struct ug4b200_gather {
	int64_t nglobal = 0, nlocal = 0; int block = 1;
	int* d_l2g = nullptr;
	bool p2p = false, committed = false;
	int id = -1;
	size_t win_off = 0, recv_off = 0;
	GatherDst* d_dst = nullptr;
	unsigned long long* d_epoch = nullptr; unsigned int* d_counters = nullptr;
};

const UgPushDev* ug_iface_push_begin(ug4b200_ctx* ctx, ug4b200_interface* I, const double* vec)
{
	if (!I || !I->p2p || !I->committed || !I->d_push ) return nullptr;
	I->pushed_vec = vec;
	return I->d_push;
}

extern "C" {

int ug4b200_comm_unique_id(unsigned char id[UG4B200_NCCL_ID_BYTES])
{
	NcclApi& N = nccl();
	if (!N.ok) return ug4b200_fail(nullptr, UG4B200_ERR_NCCL, "NCCL unavailable: " + N.why);
	ncclUniqueId u;
	UG_NCCL(nullptr, N.GetUniqueId(&u));
	std::memcpy(id, u.internal, UG4B200_NCCL_ID_BYTES);
	return UG4B200_OK;
}

int ug4b200_comm_init(ug4b200_ctx* ctx, int nranks, int rank, const unsigned char id[UG4B200_NCCL_ID_BYTES])
{
	NcclApi& N = nccl();
	if (!N.ok) return ug4b200_fail(ctx, UG4B200_ERR_NCCL, "NCCL unavailable: " + N.why);
	UG_ARG(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
	if (ctx->nccl) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "communicator already initialised");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	ncclUniqueId u;
	std::memcpy(u.internal, id, UG4B200_NCCL_ID_BYTES);
	ncclComm_t c = nullptr;
	UG_NCCL(ctx, N.CommInitRank(&c, nranks, u, rank));
	ctx->nccl = c; ctx->nranks = nranks; ctx->rank = rank;
	return UG4B200_OK;
}

int ug4b200_comm_destroy(ug4b200_ctx* ctx)
{
	if (ctx->nccl) { ug_batch_flush(ctx); cudaStreamSynchronize(ctx->stream); nccl().CommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
	if (!(ctx->p2p && ctx->p2p->nranks > 1)) { ctx->nranks = 1; ctx->rank = 0; }
	return UG4B200_OK;
}

int ug4b200_allreduce_sum(ug4b200_ctx* ctx, double* dev, int n)
{
	if (ctx->nranks <= 1) return UG4B200_OK;
	if (ctx->p2p && n <= kP2PArMax) {
		if (n <= 0) return UG4B200_OK;
		int threads = ((n + 31) / 32) * 32; if (threads < 32) threads = 32; if (threads > 1024) threads = 1024;
		UG_LAUNCH(ctx, p2p_allreduce_kernel, 1, threads, 0, ug_ar_of(ctx), dev, n, ctx->guard);
		return UG4B200_OK;
	}
	if (!ctx->nccl) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "communicator not initialised");
	UG_FLUSH(ctx);
	UG_NCCL(ctx, nccl().AllReduce(dev, dev, (size_t)n, kNcclDouble, kNcclSum, (ncclComm_t)ctx->nccl, ctx->stream));
	return UG4B200_OK;
}

int ug4b200_interface_create(ug4b200_ctx* ctx, int nneigh, const int* neigh_rank, const int64_t* neigh_ptr,
                             const int* indices, int64_t nlocal, ug4b200_interface** out)
{
	UG_ARG(ctx, out && nneigh >= 0 && nlocal >= 0, "bad argument");
	*out = nullptr;
	ug4b200_interface* I = new ug4b200_interface;
	I->nneigh = nneigh; I->nlocal = nlocal;
	I->rank.assign(neigh_rank, neigh_rank + nneigh);
	I->ptr.assign(neigh_ptr, neigh_ptr + nneigh + 1);
	I->total = nneigh ? neigh_ptr[nneigh] : 0;
	for (int p = 0; p < nneigh; ++p) {
		if (neigh_rank[p] == ctx->rank || neigh_rank[p] < 0 || neigh_rank[p] >= ctx->nranks) {
			delete I; return ug4b200_fail(ctx, UG4B200_ERR_ARG, "interface: bad neighbour rank");
		}
		if (p > 0 && neigh_rank[p] <= neigh_rank[p - 1]) {
			delete I; return ug4b200_fail(ctx, UG4B200_ERR_ARG, "interface: neighbours must be sorted by rank");
		}
	}
	// per local index: sharers in ascending rank order (neighbour lists are rank-sorted)
	std::vector<std::vector<int> > src(nlocal);
	std::vector<char> own_in(nlocal > 0 ? nlocal : 1, 0);
	std::vector<int> minrank(nlocal > 0 ? nlocal : 1, ctx->rank);
	for (int p = 0; p < nneigh; ++p)
		for (int64_t e = neigh_ptr[p]; e < neigh_ptr[p + 1]; ++e) {
			const int li = indices[e];
			if (li < 0 || li >= nlocal) { delete I; return ug4b200_fail(ctx, UG4B200_ERR_ARG, "interface: index out of range"); }
			if (neigh_rank[p] > ctx->rank && !own_in[li]) { src[li].push_back(-1); own_in[li] = 1; }
			src[li].push_back((int)e);
			minrank[li] = std::min(minrank[li], neigh_rank[p]);
		}
	std::vector<int> uidx, uptr(1, 0), usrc, slave;
	std::vector<unsigned char> owned(nlocal > 0 ? nlocal : 1, 1);
	for (int64_t li = 0; li < nlocal; ++li) {
		if (src[li].empty()) continue;
		if (!own_in[li]) src[li].push_back(-1); // own rank is the largest sharer
		uidx.push_back((int)li);
		for (int s : src[li]) usrc.push_back(s);
		uptr.push_back((int)usrc.size());
		if (minrank[li] < ctx->rank) { slave.push_back((int)li); owned[li] = 0; }
	}
	I->nu = (int64_t)uidx.size(); I->nslave = (int64_t)slave.size();
	auto up = [&](void** d, const void* h, size_t bytes) -> int {
		if (bytes == 0) bytes = 8;
		if (cudaMalloc(d, bytes) != cudaSuccess) { cudaGetLastError(); return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "interface: out of device memory"); }
		if (h && cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
			return ug4b200_fail(ctx, UG4B200_ERR_CUDA, "interface: upload failed");
		return 0;
	};
	int rc = 0;
	if (!rc) rc = up((void**)&I->d_idx, indices, sizeof(int) * I->total);
	if (!rc) rc = up((void**)&I->d_uidx, uidx.data(), sizeof(int) * uidx.size());
	if (!rc) rc = up((void**)&I->d_uptr, uptr.data(), sizeof(int) * uptr.size());
	if (!rc) rc = up((void**)&I->d_usrc, usrc.data(), sizeof(int) * usrc.size());
	if (!rc) rc = up((void**)&I->d_slave, slave.data(), sizeof(int) * slave.size());
	if (!rc) rc = up((void**)&I->d_owned, owned.data(), owned.size());
	if (!rc && ctx->p2p && ctx->p2p->nranks > 1 && I->total > 0) rc = p2p_interface_setup(ctx, I, up, uptr, usrc);
	if (!rc && I->p2p) {
		// fused push: which rows are interface rows (bitmap + running count per 32-row slice) and, per
		// interface row (ascending local index = order of uidx), where its value goes
		const int64_t ns = (nlocal + 31) / 32;
		std::vector<unsigned int> mask(ns > 0 ? ns : 1, 0u); std::vector<int> prefix(ns > 0 ? ns : 1, 0);
		for (int li : uidx) mask[li >> 5] |= 1u << (li & 31);
		for (int64_t sidx = 1; sidx < ns; ++sidx) prefix[sidx] = prefix[sidx - 1] + __builtin_popcount(mask[sidx - 1]);
		std::vector<int> upos(nlocal > 0 ? nlocal : 1, -1);
		for (size_t u = 0; u < uidx.size(); ++u) upos[uidx[u]] = (int)u;
		std::vector<int> sptr(uidx.size() + 1, 0), scode(I->total > 0 ? I->total : 1, 0);
		for (int64_t e = 0; e < I->total; ++e) sptr[upos[indices[e]] + 1]++;
		for (size_t u = 0; u < uidx.size(); ++u) sptr[u + 1] += sptr[u];
		std::vector<int> fill(sptr.begin(), sptr.end() - 1);
		for (int p = 0; p < nneigh; ++p)
			for (int64_t e = neigh_ptr[p]; e < neigh_ptr[p + 1]; ++e) scode[fill[upos[indices[e]]]++] = (int)(((e - neigh_ptr[p]) << 5) | p);
		if (!rc) rc = up((void**)&I->d_rowmask, mask.data(), sizeof(unsigned int) * mask.size());
		if (!rc) rc = up((void**)&I->d_rowprefix, prefix.data(), sizeof(int) * prefix.size());
		if (!rc) rc = up((void**)&I->d_sptr, sptr.data(), sizeof(int) * sptr.size());
		if (!rc) rc = up((void**)&I->d_scode, scode.data(), sizeof(int) * scode.size());
		std::vector<int> scode1(uidx.size() > 0 ? uidx.size() : 1, -1);
		for (size_t u = 0; u < uidx.size(); ++u) if (sptr[u + 1] - sptr[u] == 1) scode1[u] = scode[sptr[u]];
		if (!rc) rc = up((void**)&I->d_scode1, scode1.data(), sizeof(int) * scode1.size());
		if (!rc) rc = up((void**)&I->d_push_arrive, nullptr, 8);
		if (!rc) rc = up((void**)&I->d_push, nullptr, sizeof(UgPushDev));
		if (!rc && cudaMemsetAsync(I->d_push_arrive, 0, 8, ctx->stream) != cudaSuccess) rc = ug4b200_fail(ctx, UG4B200_ERR_CUDA, "interface: memset failed");
	}
	if (!rc && !I->p2p) {
		rc = up((void**)&I->sendbuf, nullptr, sizeof(double) * 9 * I->total);
		if (!rc) rc = up((void**)&I->recvbuf, nullptr, sizeof(double) * 9 * I->total);
	}
	if (rc) { ug4b200_interface_destroy(ctx, I); return rc; }
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	*out = I;
	return UG4B200_OK;
}

int ug4b200_interface_destroy(ug4b200_ctx* ctx, ug4b200_interface* I)
{
	if (!I) return UG4B200_OK;
	if (ctx) { ug_batch_flush(ctx); cudaStreamSynchronize(ctx->stream); }
	cudaFree(I->d_idx); cudaFree(I->d_uidx); cudaFree(I->d_uptr); cudaFree(I->d_usrc); cudaFree(I->d_slave);
	cudaFree(I->d_owned); cudaFree(I->sendbuf); cudaFree(I->recvbuf);
	cudaFree(I->d_ent_code); cudaFree(I->d_uell); cudaFree(I->d_nb); cudaFree(I->d_epoch); cudaFree(I->d_counters);
	cudaFree(I->d_rowmask); cudaFree(I->d_rowprefix); cudaFree(I->d_sptr); cudaFree(I->d_scode); cudaFree(I->d_scode1); cudaFree(I->d_push_arrive); cudaFree(I->d_push);
	if (ctx && ctx->armed_iface == I) { ctx->armed_iface = nullptr; ctx->armed_vec = nullptr; }
	if (I->p2p && ctx && ctx->p2p) {
		// window space is recycled once no interface is alive (ids keep counting: all ranks create and
		// destroy interfaces in the same order)
		if (--ctx->p2p->live_ifaces == 0) ctx->p2p->bump = kP2PHeapOff;
	}
	delete I;
	return UG4B200_OK;
}

int ug4b200_interface_commit(ug4b200_ctx* ctx, ug4b200_interface* I)
{
	UG_ARG(ctx, I != nullptr, "interface is NULL");
	if (!I->p2p || I->committed) return UG4B200_OK;
	if (ctx->capturing) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "interface_commit during graph capture: commit the interface first");
	ug4b200_p2p* P = ctx->p2p;
	std::vector<P2PNb> nb(I->nneigh);
	for (int p = 0; p < I->nneigh; ++p) {
		const int r = I->rank[p];
		const char* src = P->peer[r] + sizeof(ug4b200_p2p_entry) * ((size_t)(I->id % kP2PMaxIfaces) * kP2PMaxRanks + P->rank);
		ug4b200_p2p_entry ent{};
		// the neighbour publishes this entry when IT creates interface number I->id
		for (int tries = 0;; ++tries) {
			UG_CUDA(ctx, cudaMemcpyAsync(&ent, src, sizeof(ent), cudaMemcpyDefault, P->aux));
			UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
			if (ent.tag == (unsigned long long)I->id + 1ull) break;
			if (tries > 150000) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "interface_commit: neighbour rank never published its window entry");
			usleep(200);
		}
		if ((int64_t)ent.total <= 0 || (int64_t)(ent.ptr_me + (unsigned long long)(I->ptr[p + 1] - I->ptr[p])) > (int64_t)ent.total)
			return ug4b200_fail(ctx, UG4B200_ERR_STATE, "interface_commit: neighbour's interface does not match mine");
		nb[p].rbase = reinterpret_cast<double*>(P->peer[r] + ent.recv_off);
		nb[p].rpar_stride = (int64_t)ent.total * 9;
		nb[p].rptr = (int64_t)ent.ptr_me;
		nb[p].rflag = reinterpret_cast<unsigned long long*>(P->peer[r] + ent.flag_off);
	}
	UG_CUDA(ctx, cudaMemcpyAsync(I->d_nb, nb.data(), sizeof(P2PNb) * nb.size(), cudaMemcpyHostToDevice, P->aux));
	if (I->d_push) {
		UgPushDev pd{};
		pd.nneigh = I->nneigh; pd.rowmask = I->d_rowmask; pd.rowprefix = I->d_rowprefix; pd.sptr = I->d_sptr; pd.scode = I->d_scode; pd.scode1 = I->d_scode1;
		pd.nb = I->d_nb; pd.epoch = I->d_epoch; pd.arrive = I->d_push_arrive;   // (arrive: unused by the tagged-slot push)
		UG_CUDA(ctx, cudaMemcpyAsync(I->d_push, &pd, sizeof(pd), cudaMemcpyHostToDevice, P->aux));
	}
	UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
	I->committed = true;
	return UG4B200_OK;
}

static int exchange_sum(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block, int unique)
{
	UG_ARG(ctx, I && v && block >= 1 && block <= 9, "bad argument");
	ctx->armed_iface = nullptr; ctx->armed_vec = nullptr;   // an arm that no producer consumed is void
	if (I->total == 0) return UG4B200_OK;
	if (I->p2p) {
		if (!I->committed) { const int rc = ug4b200_interface_commit(ctx, I); if (rc) return rc; }
		P2PIfaceDev d{};
		d.nneigh = I->nneigh; d.ell_w = I->ell_w; d.total = I->total; d.nu = I->nu;
		d.idx = I->d_idx; d.ent_code = I->d_ent_code; d.nb = I->d_nb;
		d.lflag = reinterpret_cast<const unsigned long long*>(ctx->p2p->local + I->win_off);
		d.lrecv = reinterpret_cast<const double*>(ctx->p2p->local + I->recv_off);
		d.lpar_stride = I->total * 9;
		d.uidx = I->d_uidx; d.uell = I->d_uell; d.sptr = I->d_sptr; d.scode = I->d_scode;
		d.owned = I->d_owned;
		d.epoch = I->d_epoch; d.arrive = I->d_counters; d.depart = I->d_counters + 1; d.err = ctx->p2p->err_dev;
		// One CTA for small interfaces (no inter-CTA hand-shake); large ones need many SMs because one
		// SM sustains only a few GB/s of remote stores (measured: 16641 values from one CTA 42 us, from
		// 17 CTAs 18 us).  Never more CTAs than SMs (co-residency, see kernel).
		if (I->pushed_vec && (I->pushed_vec != v || block != 1 || unique)) {
			I->pushed_vec = nullptr;
			return ug4b200_fail(ctx, UG4B200_ERR_STATE, "interface: a vector whose interface rows were pushed by its producing kernel must be made consistent next");
		}
		const int pushed = I->pushed_vec == v ? 1 : 0;
		I->pushed_vec = nullptr;
		const int64_t work = I->nu * block;
		static const int cta_work = getenv("UG4B200_P2P_CTA_WORK") ? atoi(getenv("UG4B200_P2P_CTA_WORK")) : 128;
		int64_t g = work <= 2048 ? 1 : (work + cta_work - 1) / cta_work;
		if (g > ctx->num_sms) g = ctx->num_sms;
		const int threads = g > 1 ? 256 : (work >= 1024 ? 1024 : (int)std::max<int64_t>(64, (work + 31) / 32 * 32));
		UG_LAUNCH(ctx, p2p_exchange_sum_kernel, (int)g, threads, 0, d, v, block, unique, pushed, ctx->guard);
		return UG4B200_OK;
	}
	if (!ctx->nccl) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "communicator not initialised");
	NcclApi& N = nccl();
	int grid = (int)((I->total * block + 255) / 256); if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
	UG_LAUNCH(ctx, pack_kernel, grid, 256, 0, I->total, block, I->d_idx, v, I->sendbuf, ctx->guard);
	UG_NCCL(ctx, N.GroupStart());
	for (int p = 0; p < I->nneigh; ++p) {
		const size_t cnt = (size_t)(I->ptr[p + 1] - I->ptr[p]) * block;
		if (!cnt) continue;
		UG_NCCL(ctx, N.Send(I->sendbuf + I->ptr[p] * block, cnt, kNcclDouble, I->rank[p], (ncclComm_t)ctx->nccl, ctx->stream));
		UG_NCCL(ctx, N.Recv(I->recvbuf + I->ptr[p] * block, cnt, kNcclDouble, I->rank[p], (ncclComm_t)ctx->nccl, ctx->stream));
	}
	UG_NCCL(ctx, N.GroupEnd());
	grid = (int)((I->nu * block + 255) / 256); if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
	UG_LAUNCH(ctx, unpack_sum_kernel, grid, 256, 0, I->nu, block, I->d_uidx, I->d_uptr, I->d_usrc, I->recvbuf, v, ctx->guard);
	if (unique) return ug4b200_set_slaves_zero(ctx, I, v, block);
	return UG4B200_OK;
}

int ug4b200_additive_to_consistent(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block)
{ return exchange_sum(ctx, I, v, block, 0); }
int ug4b200_interface_arm(ug4b200_ctx* ctx, ug4b200_interface* I, const double* vec)
{
	ctx->armed_iface = nullptr; ctx->armed_vec = nullptr;
	// opt-in (UG4B200_FUSED_PUSH=1): measured at N = 2, 129^3 per GPU the fused push is SLOWER than the one-kernel
	// exchange (8.15 vs 7.76 ms per solve): the system-scope fence at the end of all 444 CTAs and the
	// look-ups in the epilogue of every fourth slice cost more than the push phase they take out of
	// the exchange kernel
	static const bool on = getenv("UG4B200_FUSED_PUSH") && getenv("UG4B200_FUSED_PUSH")[0] == '1';
	if (!on || !I || !vec || !I->p2p || !I->d_push || I->total == 0) return UG4B200_OK;
	if (!I->committed) { const int rc = ug4b200_interface_commit(ctx, I); if (rc) return rc; }
	ctx->armed_iface = I; ctx->armed_vec = vec;
	return UG4B200_OK;
}
int ug4b200_additive_to_unique(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block)
{ return exchange_sum(ctx, I, v, block, 1); }

/* ---- peer windows ---- */

int ug4b200_p2p_window_create(ug4b200_ctx* ctx, size_t bytes, unsigned char handle[UG4B200_IPC_HANDLE_BYTES], void** base)
{
	UG_ARG(ctx, ctx != nullptr, "ctx is NULL");
	if (ctx->p2p) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "peer window already created");
	if (bytes == 0) {
		const char* e = getenv("UG4B200_P2P_WINDOW_MB");
		bytes = (size_t)(e ? atoi(e) : 128) << 20;
	}
	if (bytes < kP2PHeapOff + (1u << 20)) bytes = kP2PHeapOff + (1u << 20);
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	ug4b200_p2p* P = new ug4b200_p2p;
	cudaError_t e = cudaMalloc(&P->local, bytes);
	if (e != cudaSuccess) { cudaGetLastError(); delete P; return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "p2p window: out of device memory"); }
	P->bytes = bytes; P->bump = kP2PHeapOff;
	cudaMemset(P->local, 0, bytes);
	cudaStreamCreateWithFlags(&P->aux, cudaStreamNonBlocking);
	cudaHostAlloc(&P->err_host, sizeof(int), cudaHostAllocMapped);
	*P->err_host = 0;
	cudaHostGetDevicePointer(&P->err_dev, P->err_host, 0);
	if (handle) {
		cudaIpcMemHandle_t h;
		e = cudaIpcGetMemHandle(&h, P->local);
		if (e != cudaSuccess) {
			cudaGetLastError(); cudaFree(P->local); cudaStreamDestroy(P->aux); cudaFreeHost(P->err_host); delete P;
			return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
		}
		static_assert(sizeof(h) == UG4B200_IPC_HANDLE_BYTES, "IPC handle size");
		std::memcpy(handle, &h, sizeof(h));
	}
	if (base) *base = P->local;
	ctx->p2p = P; // not usable until window_open / window_attach (nranks == 1)
	return UG4B200_OK;
}

int ug4b200_p2p_window_open(ug4b200_ctx* ctx, int nranks, int rank, const unsigned char* handles)
{
	UG_ARG(ctx, ctx->p2p && handles, "create the local window first");
	UG_ARG(ctx, nranks >= 1 && nranks <= kP2PMaxRanks && rank >= 0 && rank < nranks, "bad rank / nranks");
	ug4b200_p2p* P = ctx->p2p;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	for (int r = 0; r < nranks; ++r) {
		if (r == rank) { P->peer[r] = P->local; continue; }
		cudaIpcMemHandle_t h;
		std::memcpy(&h, handles + (size_t)r * UG4B200_IPC_HANDLE_BYTES, sizeof(h));
		void* ptr = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			cudaGetLastError();
			for (int q = 0; q < r; ++q) if (q != rank && P->peer[q]) { cudaIpcCloseMemHandle(P->peer[q]); P->peer[q] = nullptr; }
			return ug4b200_fail(ctx, UG4B200_ERR_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
		}
		P->peer[r] = (char*)ptr;
	}
	P->ipc = true; P->nranks = nranks; P->rank = rank;
	return p2p_finish_open(ctx, P);
}

int ug4b200_p2p_window_attach(ug4b200_ctx* ctx, int nranks, int rank, void* const* bases)
{
	UG_ARG(ctx, ctx->p2p && bases, "create the local window first");
	UG_ARG(ctx, nranks >= 1 && nranks <= kP2PMaxRanks && rank >= 0 && rank < nranks, "bad rank / nranks");
	ug4b200_p2p* P = ctx->p2p;
	UG_ARG(ctx, bases[rank] == (void*)P->local, "bases[rank] must be this context's own window");
	for (int r = 0; r < nranks; ++r) P->peer[r] = (char*)bases[r];
	P->ipc = false; P->nranks = nranks; P->rank = rank;
	return p2p_finish_open(ctx, P);
}

int ug4b200_p2p_window_destroy(ug4b200_ctx* ctx)
{
	ug4b200_p2p* P = ctx ? ctx->p2p : nullptr;
	if (!P) return UG4B200_OK;
	cudaSetDevice(ctx->device);
	ug_batch_flush(ctx);
	cudaStreamSynchronize(ctx->stream);
	if (P->ipc) for (int r = 0; r < P->nranks; ++r) if (r != P->rank && P->peer[r]) cudaIpcCloseMemHandle(P->peer[r]);
	cudaFree(P->d_peer); cudaFree(P->d_epoch); cudaFree(P->local);
	if (P->aux) cudaStreamDestroy(P->aux);
	if (P->err_host) cudaFreeHost(P->err_host);
	delete P;
	ctx->p2p = nullptr;
	if (!ctx->nccl) { ctx->nranks = 1; ctx->rank = 0; }
	return UG4B200_OK;
}

int ug4b200_p2p_enabled(const ug4b200_ctx* ctx) { return ctx && ctx->p2p && ctx->p2p->nranks > 1 ? 1 : 0; }

int ug4b200_p2p_check(ug4b200_ctx* ctx)
{
	if (ctx->p2p && ctx->p2p->err_host && *(volatile int*)ctx->p2p->err_host)
		return ug4b200_fail(ctx, UG4B200_ERR_NCCL, "peer-window exchange timed out waiting for a neighbour rank");
	return UG4B200_OK;
}

int ug4b200_vec_dot_allreduce_ds(ug4b200_ctx* ctx, int64_t n, const double* a, const double* b, ug4b200_fin fin,
                                 double* scratch_dev)
{
	UG_ARG(ctx, a && b, "NULL argument");
	if (ctx->nranks <= 1) return ug4b200_vec_dot_ds(ctx, n, a, b, fin);
	if (ctx->p2p && ctx->p2p->nranks > 1) {
		int64_t g = (n + kReduceThreads * 8 - 1) / (kReduceThreads * 8);
		const int64_t cap = std::min<int64_t>((int64_t)ctx->num_sms * 8, kMaxReduceBlocks);
		if (g > cap) g = cap; if (g < 1) g = 1;
		UG_LAUNCH(ctx, dot_allreduce_kernel, (int)g, kReduceThreads, 0, n, a, b, ctx->partials, ctx->counter, fin, ug_ar_of(ctx), ctx->guard);
		return UG4B200_OK;
	}
	// NCCL transport: local dot -> all-reduce -> finaliser
	UG_ARG(ctx, scratch_dev != nullptr, "scratch_dev needed for the NCCL transport");
	ug4b200_fin st{UG4B200_FIN_STORE, scratch_dev, nullptr, nullptr, nullptr};
	int rc = ug4b200_vec_dot_ds(ctx, n, a, b, st);
	if (!rc) rc = ug4b200_allreduce_sum(ctx, scratch_dev, 1);
	if (!rc) rc = ug4b200_scalar_fin_ds(ctx, scratch_dev, fin);
	return rc;
}

/* ---- gathered level ---- */

int ug4b200_gather_create(ug4b200_ctx* ctx, int64_t nglobal, int64_t nlocal, const int* local_to_global, int block,
                          ug4b200_gather** out)
{
	UG_ARG(ctx, out && nglobal >= 0 && nlocal >= 0 && block >= 1 && block <= 9 && (local_to_global || nlocal == 0), "bad argument");
	*out = nullptr;
	for (int64_t i = 0; i < nlocal; ++i)
		if (local_to_global[i] < 0 || local_to_global[i] >= nglobal) return ug4b200_fail(ctx, UG4B200_ERR_ARG, "gather: global index out of range");
	ug4b200_gather* G = new ug4b200_gather;
	G->nglobal = nglobal; G->nlocal = nlocal; G->block = block;
	if (cudaMalloc(&G->d_l2g, sizeof(int) * (nlocal > 0 ? nlocal : 1)) != cudaSuccess) { cudaGetLastError(); delete G; return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "gather: out of device memory"); }
	if (nlocal) cudaMemcpyAsync(G->d_l2g, local_to_global, sizeof(int) * nlocal, cudaMemcpyHostToDevice, ctx->stream);
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ug4b200_p2p* P = ctx->p2p;
	if (P && P->nranks > 1 && nglobal > 0) {
		const size_t flag_bytes = 256;
		const size_t recv_bytes = (((size_t)2 * P->nranks * nglobal * block * 8) + 255) / 256 * 256;
		if (P->bump + flag_bytes + recv_bytes <= P->bytes) {
			G->id = P->next_iface++;
			G->win_off = P->bump; G->recv_off = P->bump + flag_bytes;
			P->bump += flag_bytes + recv_bytes; P->live_ifaces++;
			G->p2p = true;
			cudaMalloc(&G->d_dst, sizeof(GatherDst) * kP2PMaxRanks);
			cudaMalloc(&G->d_epoch, 8); cudaMalloc(&G->d_counters, 8);
			cudaMemset(G->d_epoch, 0, 8); cudaMemset(G->d_counters, 0, 8);
			// slots must read 0 wherever a rank never writes; flags start at epoch 0
			UG_CUDA(ctx, cudaMemsetAsync(P->local + G->win_off, 0, flag_bytes + recv_bytes, P->aux));
			UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
			for (int r = 0; r < P->nranks; ++r) {
				if (r == P->rank) continue;
				ug4b200_p2p_entry ent{};
				ent.tag = (unsigned long long)G->id + 1ull;
				ent.recv_off = G->recv_off + (size_t)r * nglobal * block * 8;   // slot of source rank r (parity 0)
				ent.total = (unsigned long long)(nglobal * block);
				ent.flag_off = G->win_off + (size_t)r * 8;
				char* dst = P->local + sizeof(ug4b200_p2p_entry) * ((size_t)(G->id % kP2PMaxIfaces) * kP2PMaxRanks + r);
				UG_CUDA(ctx, cudaMemcpyAsync(dst, &ent, sizeof(ent), cudaMemcpyHostToDevice, P->aux));
				UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
			}
		}
	}
	*out = G;
	return UG4B200_OK;
}

int ug4b200_gather_commit(ug4b200_ctx* ctx, ug4b200_gather* G)
{
	UG_ARG(ctx, G != nullptr, "gather is NULL");
	if (!G->p2p || G->committed) return UG4B200_OK;
	if (ctx->capturing) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "gather_commit during graph capture");
	ug4b200_p2p* P = ctx->p2p;
	std::vector<GatherDst> dst(kP2PMaxRanks);
	for (int r = 0; r < P->nranks; ++r) {
		if (r == P->rank) {
			dst[r].rbase = reinterpret_cast<double*>(P->local + G->recv_off) + (size_t)r * G->nglobal * G->block;
			dst[r].rflag = reinterpret_cast<unsigned long long*>(P->local + G->win_off) + r;
			continue;
		}
		const char* src = P->peer[r] + sizeof(ug4b200_p2p_entry) * ((size_t)(G->id % kP2PMaxIfaces) * kP2PMaxRanks + P->rank);
		ug4b200_p2p_entry ent{};
		for (int tries = 0;; ++tries) {
			UG_CUDA(ctx, cudaMemcpyAsync(&ent, src, sizeof(ent), cudaMemcpyDefault, P->aux));
			UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
			if (ent.tag == (unsigned long long)G->id + 1ull) break;
			if (tries > 150000) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "gather_commit: a rank never published its window entry");
			usleep(200);
		}
		if ((int64_t)ent.total != G->nglobal * G->block) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "gather_commit: global size differs between ranks");
		dst[r].rbase = reinterpret_cast<double*>(P->peer[r] + ent.recv_off);
		dst[r].rflag = reinterpret_cast<unsigned long long*>(P->peer[r] + ent.flag_off);
	}
	UG_CUDA(ctx, cudaMemcpyAsync(G->d_dst, dst.data(), sizeof(GatherDst) * kP2PMaxRanks, cudaMemcpyHostToDevice, P->aux));
	UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
	G->committed = true;
	return UG4B200_OK;
}

int ug4b200_gather_sum(ug4b200_ctx* ctx, ug4b200_gather* G, double* global_out, const double* local_in)
{
	UG_ARG(ctx, G && global_out && (local_in || G->nlocal == 0), "bad argument");
	const int64_t tot = G->nglobal * G->block;
	if (tot == 0) return UG4B200_OK;
	if (G->p2p) {
		if (!G->committed) { const int rc = ug4b200_gather_commit(ctx, G); if (rc) return rc; }
		ug4b200_p2p* P = ctx->p2p;
		GatherDev d{};
		d.nranks = P->nranks; d.rank = P->rank; d.block = G->block; d.nlocal = G->nlocal; d.nglobal = G->nglobal;
		d.l2g = G->d_l2g; d.dst = G->d_dst;
		d.lflag = reinterpret_cast<const unsigned long long*>(P->local + G->win_off);
		d.lrecv = reinterpret_cast<const double*>(P->local + G->recv_off);
		d.epoch = G->d_epoch; d.arrive = G->d_counters; d.depart = G->d_counters + 1; d.err = P->err_dev;
		const int64_t work = std::max<int64_t>(tot, G->nlocal * G->block) * P->nranks;
		int64_t g = work <= 2048 ? 1 : (work + 1023) / 1024;
		if (g > ctx->num_sms) g = ctx->num_sms;
		const int threads = g > 1 ? 256 : (work >= 1024 ? 1024 : (int)std::max<int64_t>(64, (work + 31) / 32 * 32));
		UG_LAUNCH(ctx, p2p_gather_sum_kernel, (int)g, threads, 0, d, global_out, local_in, ctx->guard);
		return UG4B200_OK;
	}
	// single rank or NCCL transport: global = 0 ; global[l2g] = local ; all-reduce
	int rc = ug4b200_vec_set(ctx, tot, global_out, 0.0);
	if (rc) return rc;
	if (G->nlocal) {
		int grid = (int)((G->nlocal * G->block + 255) / 256); if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
		UG_LAUNCH(ctx, gather_scatter_local_kernel, grid, 256, 0, G->nlocal, G->block, G->d_l2g, global_out, local_in, ctx->guard);
	}
	if (ctx->nranks <= 1) return UG4B200_OK;
	if (tot > 2147483647LL) return ug4b200_fail(ctx, UG4B200_ERR_ARG, "gather: vector too long");
	return ug4b200_allreduce_sum(ctx, global_out, (int)tot);
}

int ug4b200_gather_destroy(ug4b200_ctx* ctx, ug4b200_gather* G)
{
	if (!G) return UG4B200_OK;
	if (ctx) { ug_batch_flush(ctx); cudaStreamSynchronize(ctx->stream); }
	cudaFree(G->d_l2g); cudaFree(G->d_dst); cudaFree(G->d_epoch); cudaFree(G->d_counters);
	if (G->p2p && ctx && ctx->p2p) { if (--ctx->p2p->live_ifaces == 0) ctx->p2p->bump = kP2PHeapOff; }
	delete G;
	return UG4B200_OK;
}

int ug4b200_set_slaves_zero(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block)
{
	UG_ARG(ctx, I && v, "bad argument");
	if (I->nslave == 0) return UG4B200_OK;
	int grid = (int)((I->nslave * block + 255) / 256); if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
	UG_LAUNCH(ctx, zero_idx_kernel, grid, 256, 0, I->nslave, block, I->d_slave, v, ctx->guard);
	return UG4B200_OK;
}

int ug4b200_vec_dot_unique_ds(ug4b200_ctx* ctx, ug4b200_interface* I, int64_t n, int block, const double* a,
                              const double* b, double* out_dev)
{
	UG_ARG(ctx, I && a && b && out_dev && n == I->nlocal * block, "bad argument");
	int64_t g = (n + kReduceThreads * 8 - 1) / (kReduceThreads * 8);
	int64_t cap = std::min<int64_t>((int64_t)ctx->num_sms * 8, kMaxReduceBlocks);
	if (g > cap) g = cap; if (g < 1) g = 1;
	ug4b200_fin fin{UG4B200_FIN_STORE, out_dev, nullptr, nullptr, nullptr};
	UG_LAUNCH(ctx, dot_unique_kernel, (int)g, kReduceThreads, 0, I->nlocal, block, I->d_owned, a, b, ctx->partials,
	          ctx->counter, fin, ctx->guard);
	return UG4B200_OK;
}

} // extern "C"
