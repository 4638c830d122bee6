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
#include "common.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
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
	double* sendbuf = nullptr; double* recvbuf = nullptr; // total*3 doubles
};

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
	ctx->nranks = 1; ctx->rank = 0;
	return UG4B200_OK;
}

int ug4b200_allreduce_sum(ug4b200_ctx* ctx, double* dev, int n)
{
	if (ctx->nranks <= 1) return UG4B200_OK;
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
	if (!rc) rc = up((void**)&I->sendbuf, nullptr, sizeof(double) * 9 * I->total);
	if (!rc) rc = up((void**)&I->recvbuf, nullptr, sizeof(double) * 9 * I->total);
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
	delete I;
	return UG4B200_OK;
}

int ug4b200_additive_to_consistent(ug4b200_ctx* ctx, ug4b200_interface* I, double* v, int block)
{
	UG_ARG(ctx, I && v && block >= 1 && block <= 9, "bad argument");
	if (I->total == 0) return UG4B200_OK;
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
