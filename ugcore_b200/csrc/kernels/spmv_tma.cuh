// spmv_tma.cuh — scalar SELL-32 SpMV with the matrix stream staged through shared memory
// by bulk asynchronous copies (cp.async.bulk, the 1-D form of TMA; SASS UBLKCP) completing
// on mbarriers.
//
// Why: the register-staged kernel needs ~80 registers per thread to keep enough matrix
// bytes in flight, which caps occupancy at 24 warps/SM and left the fused smoothing step at
// ~5.1 TB/s (profiles/r01a_kernel_variants.jsonl).  Here the values/columns of a slice never
// pass through registers on their way from HBM: every warp owns a ring of NST chunk buffers
// (KC entry columns each) that one elected lane keeps filled NST chunks ahead, across slice
// boundaries.  In-flight bytes per SM are 16 warps * 12 KB = 192 KB, independent of register
// pressure; the only register-staged loads left are the x-gather (L1/L2 resident) and the
// per-row vectors.  Arithmetic is unchanged: one thread per row, ascending column order, no FMA.
//
// Two entry-stream formats (template parameter COMP):
//   plain   8-byte value + 4-byte column per entry; KC = 16 columns per chunk (6 KB), ring of 2
//   COMP    value-indexed: u16 dictionary index + u16 column offset from the slice's smallest
//           column (4 B per entry instead of 12, lossless); KC = 32 (4 KB), ring of 3.  The
//           dictionary (a few dozen doubles on a uniformly refined level) is read through L1.
// A chunk is consumed in register batches of UB = 16 entries.
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
#ifndef UG_TMA_KCC
#define UG_TMA_KCC 32
#endif
#ifndef UG_TMA_NSTC
#define UG_TMA_NSTC 3
#endif
#ifndef UG_TMA_WPB
#define UG_TMA_WPB 8
#endif
#ifndef UG_TMA_MINCTA
#define UG_TMA_MINCTA 2
#endif
constexpr int WPB = UG_TMA_WPB;   // warps per CTA
constexpr int UB = 16;            // entries per register batch

template <bool COMP> struct Cfg;
template <> struct Cfg<false> {
	static constexpr int KC = UG_TMA_KC, NST = UG_TMA_NST, VB = 8, CB = 4;
	static constexpr int WARP_BYTES = NST * KC * 32 * (VB + CB) + 64;
	static constexpr int SMEM_BYTES = WPB * WARP_BYTES + 128;
};
template <> struct Cfg<true> {
	static constexpr int KC = UG_TMA_KCC, NST = UG_TMA_NSTC, VB = 2, CB = 2;
	static constexpr int WARP_BYTES = NST * KC * 32 * (VB + CB) + 64;
	static constexpr int SMEM_BYTES = WPB * WARP_BYTES + 128;
};

// Shared addresses are handled as 32-bit integers derived ONCE from an opaque (volatile asm)
// conversion: with the __cvta intrinsic the compiler rematerialises the CTA's shared window
// (MOV 0x400 ; S2R SR_CgaCtaId ; LEA) at every use under register pressure, and S2R runs on the XU
// pipe — measured at 107 % utilisation, the limiter of the value-indexed kernel.
__device__ __forceinline__ uint32_t smem_base_opaque(const void* p)
{
	uint32_t a;
	asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(a) : "l"(p));
	return a;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t.reg .pred P1;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
	    "@P1 bra.uni WAIT_DONE;\n\t"
	    "bra.uni WAIT_LOOP;\n\t"
	    "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory loads by 32-bit shared address: a generic pointer into dynamic shared memory makes
// the compiler re-derive the CTA's shared window (S2R SR_CgaCtaId, XU pipe) for every access —
// measured as the top limiter of the value-indexed kernel (XU pipe at 107 %)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u16(uint32_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (int)v; }

// position of a warp in its chunk sequence
struct Cursor {
	int64_t s;      // slice
	int64_t base;   // entry offset of the slice
	int width;      // entry columns of the slice
	int k0;         // first entry column of the current chunk
};

constexpr int SDICT_MAX = 1024;   // dictionaries up to this size are staged in shared memory (8 KB)

template <int BETAK, int MODE, int FUSE, bool COMP, bool SDICT = false>
__global__ void __launch_bounds__(WPB * 32, UG_TMA_MINCTA)
spmv1_tma_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
                 Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	typedef Cfg<COMP> C;
	constexpr int KC = C::KC, NST = C::NST, VB = C::VB, CB = C::CB;
	constexpr int CHUNK = KC * 32;                     // entries per chunk
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t smem0 = smem_base_opaque(smem_raw);
	const uint32_t wbase = smem0 + (uint32_t)wid * C::WARP_BYTES;
	const uint32_t vbuf_s = wbase;                                   // [NST][CHUNK] values / value indices
	const uint32_t cbuf_s = wbase + NST * CHUNK * VB;                // [NST][CHUNK] columns / column offsets
	const uint32_t bars = wbase + NST * CHUNK * (VB + CB);           // [NST] mbarriers (8 bytes each)
	const int64_t gwarp = (int64_t)blockIdx.x * WPB + wid;
	const int64_t nwarps = (int64_t)gridDim.x * WPB;
	// dictionary of the value-indexed stream: behind the rings in shared memory when it is small
	double* sdict = reinterpret_cast<double*>(smem_raw + (size_t)WPB * C::WARP_BYTES + 128);
	if (COMP && SDICT) {
		for (int i = threadIdx.x; i < A.ndict; i += blockDim.x) sdict[i] = A.dict[i];
		__syncthreads();
	}
	const uint32_t sdict_s = smem0 + (uint32_t)WPB * C::WARP_BYTES + 128u;
	auto dict_at = [&](int vi) -> double { return SDICT ? lds_f64(sdict_s + (uint32_t)vi * 8u) : __ldg(A.dict + vi); };

	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < NST; ++i) mbar_init(bars + i * 8, 1);
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
				const int64_t e0 = c.base + (int64_t)c.k0 * 32;
				const uint32_t bar = bars + st * 8;
				mbar_expect_tx(bar, (uint32_t)nk * 32 * (VB + CB));
				if (COMP) {
					bulk_g2s(vbuf_s + st * CHUNK * VB, A.vidx + e0, (uint32_t)nk * 32 * VB, bar);
					bulk_g2s(cbuf_s + st * CHUNK * CB, A.cidx + e0, (uint32_t)nk * 32 * CB, bar);
				} else {
					bulk_g2s(vbuf_s + st * CHUNK * VB, A.vals + e0, (uint32_t)nk * 32 * VB, bar);
					bulk_g2s(cbuf_s + st * CHUNK * CB, A.cols + e0, (uint32_t)nk * 32 * CB, bar);
				}
			} else mbar_arrive(bars + st * 8);
		}
	};

	Cursor prod; prod.s = gwarp; load_slice(prod);
	// prologue: fill the ring
#pragma unroll
	for (int i = 0; i < NST; ++i) {
		if (prod.s < A.num_slices) { issue(prod, i); advance(prod); }
	}
	// Consumer side: the metadata of the NEXT slice (entry offset, width, this lane's row length, column
	// base) is requested while the current slice is processed, so that no slice starts with an exposed
	// HBM round trip (measured: 30 % of all stall samples before this was done).
	struct Meta { int64_t s; int width, len, cbase; double acc0; };
	auto fetch_meta = [&](int64_t sl) {
		Meta m; m.s = sl; m.width = 0; m.len = 0; m.cbase = 0; m.acc0 = 0.0;
		if (sl < A.num_slices) {
			m.width = (int)((A.slice_ptr[sl + 1] - A.slice_ptr[sl]) >> 5);
			m.len = A.rowlen[sl * 32 + lane];
			if (COMP) m.cbase = A.colbase[sl];
			// the row sum starts from dest / v: the first addition must not wait for HBM either
			if (MODE == MODE_INPLACE) { if (sl * 32 + lane < A.nrows) m.acc0 = dest[sl * 32 + lane]; }
			else if (MODE == MODE_GENERAL) { if (sl * 32 + lane < A.nrows) m.acc0 = v[sl * 32 + lane]; }
		}
		return m;
	};
	Meta cur = fetch_meta(gwarp);
	int stage = 0; uint32_t phase = 0;
	double dot = 0.0;
	while (cur.s < A.num_slices) {
		const Meta nxt = fetch_meta(cur.s + nwarps);
		const int64_t row = cur.s * 32 + lane;
		const bool live = row < A.nrows;
		const int len = cur.len;
		const int width = cur.width;
		const int cbase = cur.cbase;
		// per-row streams first: their latency overlaps the whole slice
		double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
		if (MODE == MODE_INPLACE) acc = cur.acc0;
		else if (MODE == MODE_GENERAL) acc = alpha * cur.acc0;
		if (FUSE == FUSE_DOT) { if (live) own = w[row]; }
		if (FUSE == FUSE_JACOBI && live) {
			if (fz.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
			if ((fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) scv = fz.sc[row];
			if (fz.flags & UG4B200_SMOOTH_JACOBI) dinv = fz.diaginv[row];
		}
		int k0 = 0;
		do {
			const int nk = min(KC, width - k0);
			mbar_wait(bars + stage * 8, phase);
			if constexpr (COMP) {
				// All x-gathers of the chunk are issued before any arithmetic (up to 32 independent loads in
				// flight per lane; the column offsets come from shared memory and are not kept), then the
				// row sum is built in column order with the values looked up through the dictionary.
				const uint32_t vs = vbuf_s + (uint32_t)(stage * CHUNK + lane) * 2u;
				const uint32_t cs = cbuf_s + (uint32_t)(stage * CHUNK + lane) * 2u;
				const double* wb = w + cbase;
				double x[KC];
				if (__all_sync(0xffffffffu, len >= k0 + nk)) {      // every lane owns all nk entries (interior slices)
#pragma unroll
					for (int u = 0; u < KC; ++u)
						if (u < nk) x[u] = __ldg(wb + lds_u16(cs + u * 64));
#pragma unroll
					for (int u = 0; u < KC; ++u) {
						if (u < nk) {
							const double t = mulbeta<BETAK>(dict_at(lds_u16(vs + u * 64)), beta) * x[u];
							if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && u == 0 && k0 == 0) acc = t;
							else acc = acc + t;
						}
					}
				} else {
#pragma unroll
					for (int u = 0; u < KC; ++u)
						if (k0 + u < len) x[u] = __ldg(wb + lds_u16(cs + u * 64));
#pragma unroll
					for (int u = 0; u < KC; ++u) {
						if (k0 + u < len) {
							const double t = mulbeta<BETAK>(dict_at(lds_u16(vs + u * 64)), beta) * x[u];
							if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && u == 0 && k0 == 0) acc = t;
							else acc = acc + t;
						}
					}
				}
			} else {
#pragma unroll
				for (int ub = 0; ub < KC; ub += UB) {
					if (ub < nk) {
						double a[UB]; int c[UB]; double x[UB];
						const uint32_t vs = vbuf_s + (uint32_t)(stage * CHUNK + ub * 32 + lane) * 8u;
						const uint32_t cs = cbuf_s + (uint32_t)(stage * CHUNK + ub * 32 + lane) * 4u;
#pragma unroll
						for (int u = 0; u < UB; ++u)
							if (ub + u < nk) { a[u] = lds_f64(vs + u * 256); c[u] = lds_s32(cs + u * 128); }
#pragma unroll
						for (int u = 0; u < UB; ++u)
							if (k0 + ub + u < len) x[u] = __ldg(w + c[u]);
#pragma unroll
						for (int u = 0; u < UB; ++u) {
							if (k0 + ub + u < len) {
								const double t = mulbeta<BETAK>(a[u], beta) * x[u];
								if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && k0 + ub + u == 0) acc = t;
								else acc = acc + t;
							}
						}
					}
				}
			}
			__syncwarp();                       // every lane has consumed this stage: refill it
			if (prod.s < A.num_slices) { issue(prod, stage); advance(prod); }
			if (++stage == NST) { stage = 0; phase ^= 1u; }
			k0 += KC;
		} while (k0 < width);
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
		cur = nxt;
	}
	if (FUSE == FUSE_DOT) ug_block_reduce_fin(dot, fz.partials, fz.counter, fz.fin, fz.ar);
}

} // namespace tma
