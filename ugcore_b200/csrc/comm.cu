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
struct P2PNb {
	double* rbase;               // neighbour's receive region (in ITS window)
	int64_t rpar_stride;         // doubles between its two parity buffers
	int64_t rptr;                // my first entry inside its region
	unsigned long long* rflag;   // flag in its window that I raise
};
struct P2PIfaceDev {
	int nneigh; int64_t total, nu;
	const int* idx; const unsigned char* ent_nb; const int64_t* ptr;
	const P2PNb* nb;
	const unsigned long long* lflag;   // [nneigh] in my window, raised by the neighbours
	const double* lrecv;               // my receive region (2 parities x total x 9 doubles)
	int64_t lpar_stride;
	const int* uidx; const int* uptr; const int* usrc;
	unsigned long long* epoch; unsigned int* arrive; unsigned int* depart; int* err;
};

// AdditiveToConsistent in one kernel (grid <= #SMs so that all CTAs are co-resident:
// CTAs that wait for a neighbour must not keep CTAs that still have to send off the SMs)
__global__ void __launch_bounds__(256)
p2p_exchange_sum_kernel(P2PIfaceDev d, double* v, int block, const int* guard)
{
	if (ug_guarded(guard)) return;
	__shared__ bool s_last;
	const unsigned long long e = *(volatile unsigned long long*)d.epoch + 1ull;
	const int64_t par = (int64_t)(e & 1ull);
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	// 1. push my interface values into the neighbours' windows
	for (int64_t t = tid; t < d.total * block; t += stride) {
		const int64_t en = t / block; const int q = (int)(t - en * block);
		const int n = d.ent_nb[en];
		const P2PNb nb = d.nb[n];
		double* dst = nb.rbase + par * nb.rpar_stride + (nb.rptr + (en - d.ptr[n])) * block + q;
		ug_st_relaxed_sys(dst, v[(int64_t)d.idx[en] * block + q]);
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		s_last = true;
		if (gridDim.x > 1) {
			__threadfence();
			s_last = (atomicAdd(d.arrive, 1u) == gridDim.x - 1);
			if (s_last) { *d.arrive = 0u; __threadfence_system(); }
		}
	}
	__syncthreads();
	// 2. everything of this rank is on its way: raise my flag at every neighbour
	if (s_last && threadIdx.x < d.nneigh) ug_st_release_sys(d.nb[threadIdx.x].rflag, e);
	// 3. wait for the neighbours
	if (threadIdx.x < d.nneigh) ug_wait_flag(d.lflag + threadIdx.x, e, d.err);
	__syncthreads();
	// 4. sum all copies in ascending rank order (own value where usrc < 0)
	const double* recv = d.lrecv + par * d.lpar_stride;
	for (int64_t t = tid; t < d.nu * block; t += stride) {
		const int64_t u = t / block; const int q = (int)(t - u * block);
		const int64_t li = (int64_t)d.uidx[u] * block + q;
		double s = 0.0;
		for (int p = d.uptr[u]; p < d.uptr[u + 1]; ++p) {
			const int src = d.usrc[p];
			const double x = src < 0 ? v[li] : ug_ld_relaxed_sys(recv + (int64_t)src * block + q);
			s = (p == d.uptr[u]) ? x : s + x;
		}
		v[li] = s;
	}
	// 5. the last CTA to leave publishes the epoch (every CTA read it on entry)
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
	unsigned char* d_ent_nb = nullptr; int64_t* d_ptr = nullptr; P2PNb* d_nb = nullptr;
	unsigned long long* d_epoch = nullptr; unsigned int* d_counters = nullptr;
};


// Allocate this interface's flags + receive region in the local window and publish, for every
// neighbour r, where r has to write (table entry [id][r]); neighbours look it up in commit.
template <class Up>
static int p2p_interface_setup(ug4b200_ctx* ctx, ug4b200_interface* I, Up& up)
{
	ug4b200_p2p* P = ctx->p2p;
	if (I->nneigh > 255) return ug4b200_fail(ctx, UG4B200_ERR_ARG, "interface: too many neighbours for the peer-window transport");
	const size_t flag_bytes = (((size_t)I->nneigh * 8) + 255) / 256 * 256;
	const size_t recv_bytes = (((size_t)2 * I->total * 9 * 8) + 255) / 256 * 256;
	if (P->bump + flag_bytes + recv_bytes > P->bytes)
		return ug4b200_fail(ctx, UG4B200_ERR_NOMEM, "interface: peer window exhausted (UG4B200_P2P_WINDOW_MB)");
	I->id = P->next_iface++;
	I->win_off = P->bump; I->recv_off = P->bump + flag_bytes;
	P->bump += flag_bytes + recv_bytes; P->live_ifaces++;
	I->p2p = true;
	std::vector<unsigned char> ent_nb(I->total);
	for (int p = 0; p < I->nneigh; ++p) for (int64_t e = I->ptr[p]; e < I->ptr[p + 1]; ++e) ent_nb[e] = (unsigned char)p;
	int rc = 0;
	if (!rc) rc = up((void**)&I->d_ent_nb, ent_nb.data(), ent_nb.size());
	if (!rc) rc = up((void**)&I->d_ptr, I->ptr.data(), sizeof(int64_t) * I->ptr.size());
	if (!rc) rc = up((void**)&I->d_nb, nullptr, sizeof(P2PNb) * I->nneigh);
	if (!rc) rc = up((void**)&I->d_epoch, nullptr, 8);
	if (!rc) rc = up((void**)&I->d_counters, nullptr, 8);
	if (rc) return rc;
	UG_CUDA(ctx, cudaMemsetAsync(I->d_epoch, 0, 8, ctx->stream));
	UG_CUDA(ctx, cudaMemsetAsync(I->d_counters, 0, 8, ctx->stream));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	// flags start at epoch 0 (the region may be recycled); then publish
	UG_CUDA(ctx, cudaMemsetAsync(P->local + I->win_off, 0, flag_bytes, P->aux));
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
	if (ctx->nccl) { cudaStreamSynchronize(ctx->stream); nccl().CommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
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
	if (!rc && ctx->p2p && ctx->p2p->nranks > 1 && I->total > 0) rc = p2p_interface_setup(ctx, I, up);
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
	if (ctx) cudaStreamSynchronize(ctx->stream);
	cudaFree(I->d_idx); cudaFree(I->d_uidx); cudaFree(I->d_uptr); cudaFree(I->d_usrc); cudaFree(I->d_slave);
	cudaFree(I->d_owned); cudaFree(I->sendbuf); cudaFree(I->recvbuf);
	cudaFree(I->d_ent_nb); cudaFree(I->d_ptr); cudaFree(I->d_nb); cudaFree(I->d_epoch); cudaFree(I->d_counters);
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
	UG_CUDA(ctx, cudaStreamSynchronize(P->aux));
	I->committed = true;
	return UG4B200_OK;
}

int ug4b200_additive_to_consistent(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block)
{
	UG_ARG(ctx, I && v && block >= 1 && block <= 9, "bad argument");
	if (I->total == 0) return UG4B200_OK;
	if (I->p2p) {
		if (!I->committed) { const int rc = ug4b200_interface_commit(ctx, I); if (rc) return rc; }
		P2PIfaceDev d{};
		d.nneigh = I->nneigh; d.total = I->total; d.nu = I->nu;
		d.idx = I->d_idx; d.ent_nb = I->d_ent_nb; d.ptr = I->d_ptr; d.nb = I->d_nb;
		d.lflag = reinterpret_cast<const unsigned long long*>(ctx->p2p->local + I->win_off);
		d.lrecv = reinterpret_cast<const double*>(ctx->p2p->local + I->recv_off);
		d.lpar_stride = I->total * 9;
		d.uidx = I->d_uidx; d.uptr = I->d_uptr; d.usrc = I->d_usrc;
		d.epoch = I->d_epoch; d.arrive = I->d_counters; d.depart = I->d_counters + 1; d.err = ctx->p2p->err_dev;
		// 4 values per thread; never more CTAs than SMs (co-residency, see kernel)
		int64_t g = (I->total * block + 1023) / 1024;
		if (g > ctx->num_sms) g = ctx->num_sms; if (g < 1) g = 1;
		UG_LAUNCH(ctx, p2p_exchange_sum_kernel, (int)g, 256, 0, d, v, block, ctx->guard);
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
	return UG4B200_OK;
}

/* ---- peer windows ---- */

int ug4b200_p2p_window_create(ug4b200_ctx* ctx, size_t bytes, unsigned char handle[UG4B200_IPC_HANDLE_BYTES], void** base)
{
	UG_ARG(ctx, ctx != nullptr, "ctx is NULL");
	if (ctx->p2p) return ug4b200_fail(ctx, UG4B200_ERR_STATE, "peer window already created");
	if (bytes == 0) {
		const char* e = getenv("UG4B200_P2P_WINDOW_MB");
		bytes = (size_t)(e ? atoi(e) : 64) << 20;
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
