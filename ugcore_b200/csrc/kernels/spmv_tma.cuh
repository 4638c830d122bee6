// spmv_tma.cuh — scalar SELL-32 SpMV with the matrix stream staged through shared memory
// by bulk asynchronous copies (cp.async.bulk, the 1-D form of TMA; SASS UBLKCP) completing
// on mbarriers.
//
// Why: the register-staged kernel needs ~80 registers per thread to keep enough matrix
// bytes in flight, which caps occupancy at 24 warps/SM and left the fused smoothing step at
// ~5.1 TB/s (profiles/r01a_kernel_variants.jsonl).  Here the values/columns of a slice never
// pass through registers on their way from HBM: every warp owns a ring of NST chunk buffers
// (KC entry columns = KC*32 values + KC*32 column indices = 6 KB each) that one elected lane
// keeps filled NST chunks ahead, across slice boundaries.  In-flight bytes per SM are
// 16 warps * 2 stages * 6 KB = 192 KB, independent of register pressure; the only
// register-staged loads left are the x-gather (L1/L2 resident) and the per-row vectors.
// Arithmetic is unchanged: one thread per row, ascending column order, no FMA.
//
// Grid: persistent, 2 CTAs (8 warps each, 96 KB of shared memory) per SM; warp g handles
// slices g, g + W, g + 2W, ...
#pragma once

namespace tma {

#ifndef UG_TMA_KC
#define UG_TMA_KC 16
#endif
#ifndef UG_TMA_NST
#define UG_TMA_NST 2
#endif
#ifndef UG_TMA_WPB
#define UG_TMA_WPB 8
#endif
#ifndef UG_TMA_MINCTA
#define UG_TMA_MINCTA 2
#endif
constexpr int KC = UG_TMA_KC;     // entry columns per chunk
constexpr int NST = UG_TMA_NST;   // ring depth per warp
constexpr int WPB = UG_TMA_WPB;   // warps per CTA
constexpr int CHUNK_VALS = KC * 32;                       // doubles
constexpr int WARP_BYTES = NST * CHUNK_VALS * 12 + NST * 8; // values + columns + barriers
constexpr int SMEM_BYTES = WPB * WARP_BYTES + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t.reg .pred P1;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
	    "@P1 bra.uni WAIT_DONE;\n\t"
	    "bra.uni WAIT_LOOP;\n\t"
	    "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// position of a warp in its chunk sequence
struct Cursor {
	int64_t s;      // slice
	int64_t base;   // entry offset of the slice
	int width;      // entry columns of the slice
	int k0;         // first entry column of the current chunk
};

template <int BETAK, int MODE, int FUSE>
__global__ void __launch_bounds__(WPB * 32, UG_TMA_MINCTA)
spmv1_tma_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
                 Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned char* wbase = smem_raw + (size_t)wid * WARP_BYTES;
	double* vals_s = reinterpret_cast<double*>(wbase);                                  // [NST][KC*32]
	int* cols_s = reinterpret_cast<int*>(wbase + NST * CHUNK_VALS * 8);                 // [NST][KC*32]
	uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + NST * CHUNK_VALS * 12);        // [NST]
	const int64_t gwarp = (int64_t)blockIdx.x * WPB + wid;
	const int64_t nwarps = (int64_t)gridDim.x * WPB;

	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < NST; ++i) mbar_init(&bars[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	auto load_slice = [&](Cursor& c) {
		if (c.s < A.num_slices) {
			c.base = A.slice_ptr[c.s];
			c.width = (int)((A.slice_ptr[c.s + 1] - c.base) >> 5);
		} else { c.base = 0; c.width = 0; }
		c.k0 = 0;
	};
	auto advance = [&](Cursor& c) {
		c.k0 += KC;
		if (c.k0 >= c.width) { c.s += nwarps; load_slice(c); }
	};
	// issue the chunk at cursor c into ring stage st (every slice has >= 1 chunk, possibly empty)
	auto issue = [&](const Cursor& c, int st) {
		if (lane == 0) {
			const int nk = min(KC, c.width - c.k0);
			if (nk > 0) {
				mbar_expect_tx(&bars[st], (uint32_t)nk * 32 * 12);
				bulk_g2s(vals_s + st * CHUNK_VALS, A.vals + c.base + (int64_t)c.k0 * 32, (uint32_t)nk * 256, &bars[st]);
				bulk_g2s(cols_s + st * CHUNK_VALS, A.cols + c.base + (int64_t)c.k0 * 32, (uint32_t)nk * 128, &bars[st]);
			} else mbar_arrive(&bars[st]);
		}
	};

	Cursor prod; prod.s = gwarp; load_slice(prod);
	Cursor cons = prod;
	// prologue: fill the ring
#pragma unroll
	for (int i = 0; i < NST; ++i) {
		if (prod.s < A.num_slices) { issue(prod, i); advance(prod); }
	}
	int stage = 0; uint32_t phase = 0;
	double dot = 0.0;
	while (cons.s < A.num_slices) {
		const int64_t row = cons.s * 32 + lane;
		const bool live = row < A.nrows;
		const int len = A.rowlen[row];
		const int width = cons.width;
		// per-row streams first: their latency overlaps the whole slice
		double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
		if (MODE == MODE_INPLACE) { if (live) acc = dest[row]; }
		else if (MODE == MODE_GENERAL) { if (live) acc = alpha * v[row]; }
		if (FUSE == FUSE_DOT) { if (live) own = w[row]; }
		if (FUSE == FUSE_JACOBI && live) {
			if (fz.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
			if ((fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) scv = fz.sc[row];
			if (fz.flags & UG4B200_SMOOTH_JACOBI) dinv = fz.diaginv[row];
		}
		const int64_t my_slice = cons.s;
		do {
			const int k0 = cons.k0;
			const int nk = min(KC, width - k0);
			mbar_wait(&bars[stage], phase);
			const double* vs = vals_s + stage * CHUNK_VALS + lane;
			const int* cs = cols_s + stage * CHUNK_VALS + lane;
			double a[KC]; int c[KC]; double x[KC];
#pragma unroll
			for (int u = 0; u < KC; ++u)
				if (u < nk) { a[u] = vs[u * 32]; c[u] = cs[u * 32]; }
#pragma unroll
			for (int u = 0; u < KC; ++u)
				if (k0 + u < len) x[u] = __ldg(w + c[u]);
#pragma unroll
			for (int u = 0; u < KC; ++u) {
				if (k0 + u < len) {
					const double t = mulbeta<BETAK>(a[u], beta) * x[u];
					if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && k0 + u == 0) acc = t;
					else acc = acc + t;
				}
			}
			__syncwarp();                       // every lane has consumed this stage: refill it
			if (prod.s < A.num_slices) { issue(prod, stage); advance(prod); }
			if (++stage == NST) { stage = 0; phase ^= 1u; }
			advance(cons);
		} while (cons.s == my_slice);
		if (FUSE == FUSE_JACOBI) {
			if (live) {
				dest[row] = acc;
				if (fz.flags & UG4B200_SMOOTH_ADD_IN) scv = scv + own;
				if (fz.flags & UG4B200_SMOOTH_JACOBI) {
					const double st = dinv * acc;
					fz.st_out[row] = st;
					if (fz.flags & UG4B200_SMOOTH_ADD_OUT) scv = scv + st;
				}
				if (fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) fz.sc[row] = scv;
			}
		} else {
			if (live && (MODE != MODE_ASSIGN_SKIP_EMPTY || len > 0)) dest[row] = acc;
			if (FUSE == FUSE_DOT && live) dot += acc * own;
		}
	}
	if (FUSE == FUSE_DOT) ug_block_reduce_fin(dot, fz.partials, fz.counter, fz.fin, fz.ar);
}

} // namespace tma
