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
// Two kernels, one per entry-stream format:
//   spmv1_tma_kernel  plain stream: 8-byte value + 4-byte column per entry; KC = 16 entry columns per
//                     chunk (6 KB), ring of 2 per warp; consumed in register batches of UB = 16
//   spmv1_vi_kernel   value-indexed stream: ONE 32-bit word per entry (column offset from the slice's
//                     smallest column << 16 | dictionary byte offset), 4 B per entry instead of 12,
//                     lossless; see below
//
// Grid: persistent, one wave of (#SMs x resident CTAs); warp g handles slices g, g + W, g + 2W, ...
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
constexpr int WPB = UG_TMA_WPB;   // warps per CTA
constexpr int UB = 16;            // entries per register batch

struct Cfg {
	static constexpr int KC = UG_TMA_KC, NST = UG_TMA_NST, VB = 8, CB = 4;
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

// ---------------------------------------------------------------- plain stream
template <int BETAK, int MODE, int FUSE>
__global__ void __launch_bounds__(WPB * 32, UG_TMA_MINCTA)
spmv1_tma_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
                 Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	typedef Cfg C;
	constexpr int KC = C::KC, NST = C::NST, VB = C::VB, CB = C::CB;
	constexpr int CHUNK = KC * 32;                     // entries per chunk
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t smem0 = smem_base_opaque(smem_raw);
	const uint32_t wbase = smem0 + (uint32_t)wid * C::WARP_BYTES;
	const uint32_t vbuf_s = wbase;                                   // [NST][CHUNK] values
	const uint32_t cbuf_s = wbase + NST * CHUNK * VB;                // [NST][CHUNK] columns
	const uint32_t bars = wbase + NST * CHUNK * (VB + CB);           // [NST] mbarriers (8 bytes each)
	const int64_t gwarp = (int64_t)blockIdx.x * WPB + wid;
	const int64_t nwarps = (int64_t)gridDim.x * WPB;

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
				bulk_g2s(vbuf_s + st * CHUNK * VB, A.vals + e0, (uint32_t)nk * 32 * VB, bar);
				bulk_g2s(cbuf_s + st * CHUNK * CB, A.cols + e0, (uint32_t)nk * 32 * CB, bar);
			} else mbar_arrive(bars + st * 8);
		}
	};

	Cursor prod; prod.s = gwarp; load_slice(prod);
	// prologue: fill the ring
#pragma unroll
	for (int i = 0; i < NST; ++i) {
		if (prod.s < A.num_slices) { issue(prod, i); advance(prod); }
	}
	// Consumer side: the metadata of the NEXT slice (entry offset, width, this lane's row length) is
	// requested while the current slice is processed, so that no slice starts with an exposed HBM
	// round trip (measured: 30 % of all stall samples before this was done).
	struct Meta { int64_t s; int width, len; double acc0; };
	auto fetch_meta = [&](int64_t sl) {
		Meta m; m.s = sl; m.width = 0; m.len = 0; m.acc0 = 0.0;
		if (sl < A.num_slices) {
			m.width = (int)((A.slice_ptr[sl + 1] - A.slice_ptr[sl]) >> 5);
			m.len = A.rowlen[sl * 32 + lane];
			// the row sum starts from dest / v: the first addition must not wait for HBM either
			if (MODE == MODE_INPLACE) { if (sl * 32 + lane < A.nrows) m.acc0 = dest[sl * 32 + lane]; }
			else if (MODE == MODE_GENERAL) { if (sl * 32 + lane < A.nrows) m.acc0 = v[sl * 32 + lane]; }
		}
		return m;
	};
	Meta cur = fetch_meta(gwarp);
	int stage = 0; uint32_t phase = 0;
	double dot = 0.0;
	const UgPushDev* push = FUSE == FUSE_JACOBI ? fz.push : nullptr;
	const unsigned long long pe = push ? *(volatile unsigned long long*)push->epoch + 1ull : 0ull;
	while (cur.s < A.num_slices) {
		const Meta nxt = fetch_meta(cur.s + nwarps);
		const int64_t row = cur.s * 32 + lane;
		const bool live = row < A.nrows;
		const int len = cur.len;
		const int width = cur.width;
		const unsigned int pmask = push ? push->rowmask[cur.s] : 0u;
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
					if ((pmask >> lane) & 1u) ug_push_row(push, pe, cur.s, lane, pmask, st);
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

// ---------------------------------------------------------------- value-indexed stream
// One 32-bit word per entry: (column offset from the slice's smallest column) << 16 | (dictionary
// index << vshift).  With a dictionary of <= 1024 values (SDICT: staged in shared memory at an
// 8 KB-aligned shared address, vshift = 3) the two look-ups of an entry cost one integer
// instruction each:
//     x address (low word) = word >> 13 + low word of (w + colbase)     one IMAD.HI (mul.hi by 2^19 + add)
//     dictionary address   = (word & 0x1ff8) | dictionary base          one LOP3
// so an entry is LDS.32, IMAD.HI, LDG.64, LOP3, LDS.64, DMUL, DADD.  (The 64-bit x address is
// assembled from that low word and the unchanged high word; a slice whose 512 KB window would carry
// into the high word — once per 4 GB of address space — takes the generic path.)
// A chunk (one bulk copy) is a slice's whole entry block up to VKC = VNB * VUB columns; it is consumed
// in register batches of VUB entries, software-pipelined: the x-gathers of batch j + 1 are in flight
// while batch j is accumulated.  Batches in which every lane owns all VUB entries run unpredicated.
#ifndef UG_VI_UB
#define UG_VI_UB 9
#endif
#ifndef UG_VI_NB
#define UG_VI_NB 3
#endif
#ifndef UG_VI_NST
#define UG_VI_NST 2
#endif
#ifndef UG_VI_WPB
#define UG_VI_WPB 8
#endif
#ifndef UG_VI_MINCTA
#define UG_VI_MINCTA 3
#endif
#ifndef UG_VI_PIPE
#define UG_VI_PIPE 0   // 1: gathers of batch j + 1 in flight while batch j is accumulated (two register buffers; measured slower: spills at 80 registers)
#endif
constexpr int SDICT_MAX = 1024;   // dictionaries up to this size are staged in shared memory (8 KB)
struct VCfg {
	static constexpr int UB = UG_VI_UB, NB = UG_VI_NB, KC = UB * NB, NST = UG_VI_NST, WPB = UG_VI_WPB;
	static constexpr int STAGE_BYTES = KC * 128;
	static constexpr int WARP_BYTES = NST * STAGE_BYTES;
	static constexpr int RING_BYTES = WPB * WARP_BYTES;
	static constexpr int BAR_BYTES = ((WPB * NST * 8 + 127) / 128) * 128;
	// dictionary: 8 KB at the next 8 KB-aligned shared address behind rings and barriers
	static constexpr int SMEM_BYTES_SDICT = RING_BYTES + BAR_BYTES + 2 * SDICT_MAX * 8;
	static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ double ldg_nc_f64_lohi(uint32_t lo, uint32_t hi)
{
	double v;
	asm volatile("{\n\t.reg .u64 a;\n\tmov.b64 a, {%1, %2};\n\tld.global.nc.f64 %0, [a];\n\t}" : "=d"(v) : "r"(lo), "r"(hi));
	return v;
}

template <int BETAK, int MODE, int FUSE, bool SDICT>
__global__ void __launch_bounds__(VCfg::WPB * 32, UG_VI_MINCTA)
spmv1_vi_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
                Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	typedef VCfg C;
	constexpr int UB = C::UB, NB = C::NB, KC = C::KC, NST = C::NST, VWPB = C::WPB;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t smem0 = smem_base_opaque(smem_raw);
	const uint32_t ring_s = smem0 + (uint32_t)wid * C::WARP_BYTES;             // [NST][KC][32] words
	const uint32_t bars = smem0 + C::RING_BYTES + (uint32_t)wid * NST * 8;     // [NST] mbarriers
	const uint32_t sdict_s = (smem0 + C::RING_BYTES + C::BAR_BYTES + 8191u) & ~8191u;
	// slices and rows are < 2^31 (checked at upload): 32-bit bookkeeping keeps the kernel at 80 registers
	const int gwarp = (int)blockIdx.x * VWPB + wid;
	const int nwarps = (int)gridDim.x * VWPB;
	const int nslices = (int)A.num_slices, nrows = (int)A.nrows;
	if (SDICT) {
		for (int i = threadIdx.x; i < A.ndict; i += blockDim.x)
			asm volatile("st.shared.f64 [%0], %1;" ::"r"(sdict_s + (uint32_t)i * 8u), "d"(A.dict[i]) : "memory");
		__syncthreads();
	}
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < NST; ++i) mbar_init(bars + i * 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	// ---- producer (lane 0): one bulk copy per chunk
	struct VCursor { int s, width, k0; int64_t base; };
	auto load_slice = [&](VCursor& c) {
		if (c.s < nslices) {
			c.base = A.slice_ptr[c.s];
			c.width = (int)((A.slice_ptr[c.s + 1] - c.base) >> 5);
		} else { c.base = 0; c.width = 0; }
		c.k0 = 0;
	};
	auto advance = [&](VCursor& c) {
		c.k0 += KC;
		if (c.k0 >= c.width) { c.s += nwarps; load_slice(c); }
	};
	auto issue = [&](const VCursor& c, int st) {
		if (lane == 0) {
			const int nk = min(KC, c.width - c.k0);
			const uint32_t bar = bars + st * 8;
			if (nk > 0) {
				mbar_expect_tx(bar, (uint32_t)nk * 128u);
				bulk_g2s(ring_s + st * C::STAGE_BYTES, A.vc + c.base + (int64_t)c.k0 * 32, (uint32_t)nk * 128u, bar);
			} else mbar_arrive(bar);
		}
	};
	VCursor prod; prod.s = gwarp; load_slice(prod);
#pragma unroll
	for (int i = 0; i < NST; ++i) {
		if (prod.s < nslices) { issue(prod, i); advance(prod); }
	}

	// ---- consumer
	struct Meta { int s, width, len, cbase; double acc0; };
	auto fetch_meta = [&](int sl) {
		Meta m; m.s = sl; m.width = 0; m.len = 0; m.cbase = 0; m.acc0 = 0.0;
		if (sl < nslices) {
			m.width = (int)((A.slice_ptr[sl + 1] - A.slice_ptr[sl]) >> 5);
			m.len = A.rowlen[sl * 32 + lane];
			m.cbase = A.colbase[sl];
			if (MODE == MODE_INPLACE) { if (sl * 32 + lane < nrows) m.acc0 = dest[sl * 32 + lane]; }
			else if (MODE == MODE_GENERAL) { if (sl * 32 + lane < nrows) m.acc0 = v[sl * 32 + lane]; }
		}
		return m;
	};
	Meta cur = fetch_meta(gwarp);
	int stage = 0; uint32_t phase = 0;
	double dot = 0.0;
	const UgPushDev* push = (FUSE == FUSE_JACOBI || FUSE == FUSE_RESTRICT_JACOBI) ? fz.push : nullptr;
	const unsigned long long pe = push ? *(volatile unsigned long long*)push->epoch + 1ull : 0ull;
	while (cur.s < nslices) {
		const Meta nxt = fetch_meta(cur.s + nwarps);
		const int row = cur.s * 32 + lane;
		const bool live = row < nrows;
		const int len = cur.len;
		const int width = cur.width;
		const unsigned int pmask = push ? push->rowmask[cur.s] : 0u;
		// all lanes own the first minlen entry columns of the slice: those batches run unpredicated
		const int minlen = __reduce_min_sync(0xffffffffu, len);
		const double* wb = w + cur.cbase;
		const uint32_t wlo = (uint32_t)(uintptr_t)wb, whi = (uint32_t)((uintptr_t)wb >> 32);
		const bool fast_addr = SDICT && wlo <= 0xffffffffu - (65535u << 3);
		double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
		if (MODE == MODE_INPLACE) acc = cur.acc0;
		else if (MODE == MODE_GENERAL) acc = alpha * cur.acc0;
		if (FUSE == FUSE_DOT) { if (live) own = w[row]; }
		if (FUSE == FUSE_RESTRICT_JACOBI && live) { dinv = fz.diaginv[row]; if (len == 0) own = dest[row]; }
		if (FUSE == FUSE_JACOBI && live) {
			if (fz.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
			if ((fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) scv = fz.sc[row];
			if (fz.flags & UG4B200_SMOOTH_JACOBI) dinv = fz.diaginv[row];
		}
		constexpr int NBUF = UG_VI_PIPE ? 2 : 1;
		uint32_t e[NBUF][UB]; double x[NBUF][UB];
		// entry words + x-gathers of the batch starting at chunk column kb (absolute column k0 + kb)
		auto load_batch = [&](uint32_t cs, int k0, int kb, int nk, uint32_t (&eb)[UB], double (&xb)[UB]) {
			const uint32_t a0 = cs + (uint32_t)kb * 128u;
			if (fast_addr && k0 + kb + UB <= minlen) {
#pragma unroll
				for (int u = 0; u < UB; ++u) {
					eb[u] = lds_u32(a0 + u * 128);
					xb[u] = ldg_nc_f64_lohi(__umulhi(eb[u], 1u << 19) + wlo, whi);
				}
			} else {
				// generic: words beyond the chunk's columns are not read; padding words (0) are safe to follow
#pragma unroll
				for (int u = 0; u < UB; ++u) {
					if (kb + u < nk) {
						eb[u] = lds_u32(a0 + u * 128);
						xb[u] = __ldg(wb + (eb[u] >> 16));
					}
				}
			}
		};
		auto dict_val = [&](uint32_t word) -> double {
			if (SDICT) return lds_f64((word & 0x1ff8u) | sdict_s);
			return __ldg(A.dict + ((word & 0xffffu) >> A.vshift));
		};
		auto arith_batch = [&](int k0, int kb, const uint32_t (&eb)[UB], const double (&xb)[UB]) {
			const int k = k0 + kb;
			if (k + UB <= minlen) {
#pragma unroll
				for (int u = 0; u < UB; ++u) {
					const double t = mulbeta<BETAK>(dict_val(eb[u]), beta) * xb[u];
					if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && u == 0) acc = (k == 0) ? t : acc + t;
					else acc = acc + t;
				}
			} else {
#pragma unroll
				for (int u = 0; u < UB; ++u) {
					if (k + u < len) {
						const double t = mulbeta<BETAK>(dict_val(eb[u]), beta) * xb[u];
						if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && u == 0) acc = (k == 0) ? t : acc + t;
						else acc = acc + t;
					}
				}
			}
		};
		int k0 = 0;
		do {
			const int nk = min(KC, width - k0);
			mbar_wait(bars + stage * 8, phase);
			const uint32_t cs = ring_s + (uint32_t)stage * C::STAGE_BYTES + (uint32_t)lane * 4u;
			if (UG_VI_PIPE) load_batch(cs, k0, 0, nk, e[0], x[0]);
#pragma unroll
			for (int j = 0; j < NB; ++j) {
				if (UG_VI_PIPE) {
					if (j + 1 < NB) { if ((j + 1) * UB < nk) load_batch(cs, k0, (j + 1) * UB, nk, e[(j + 1) % NBUF], x[(j + 1) % NBUF]); }
				} else {
					if (j == 0 || j * UB < nk) load_batch(cs, k0, j * UB, nk, e[0], x[0]);
				}
				if ((j == 0 || j * UB < nk) && (j + 1) * UB >= nk) {
					// the chunk's last words are in registers: every lane is done with the stage, refill it
					__syncwarp();
					if (prod.s < nslices) { issue(prod, stage); advance(prod); }
				}
				if (j * UB < nk) arith_batch(k0, j * UB, e[j % NBUF], x[j % NBUF]);
			}
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
					if ((pmask >> lane) & 1u) ug_push_row(push, pe, cur.s, lane, pmask, st);
					if (fz.flags & UG4B200_SMOOTH_ADD_OUT) scv = scv + st;
				}
				if (fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) fz.sc[row] = scv;
			}
		} else if (FUSE == FUSE_RESTRICT_JACOBI) {
			// coarse defect by restriction (rows without connections keep their value), then the first
			// Jacobi step of the coarse level on it: st = diagInv * sd
			if (live) {
				double dv = own;
				if (len > 0) { dest[row] = acc; dv = acc; }
				const double st = dinv * dv;
				fz.st_out[row] = st;
				if ((pmask >> lane) & 1u) ug_push_row(push, pe, cur.s, lane, pmask, st);
			}
		} else {
			if (live && (MODE != MODE_ASSIGN_SKIP_EMPTY || len > 0)) dest[row] = acc;
			if (FUSE == FUSE_DOT && live) dot += acc * own;
		}
		cur = nxt;
	}
	if (FUSE == FUSE_DOT) ug_block_reduce_fin(dot, fz.partials, fz.counter, fz.fin, fz.ar);
}


// ---------------------------------------------------------------- x-staged value-indexed stream
// The value-indexed words above carry a 16-bit column offset from the slice's smallest column: fine for 129^3
// (window of two grid planes = 33 282 columns), too narrow for 257^3 (132 098) — BASELINE configs[2] on one GPU fell
// back to the plain 12-byte stream (50.2 ms per solve, profiles/r02c).  This variant has no window: besides the entry
// words the producer also brings the slice's x operand into shared memory — the slice's distinct columns form a
// handful of runs of consecutive columns (27-point operator, banded numbering: 9 runs of 34), each one bulk copy
// (cp.async.bulk, >= 16 bytes, 16-byte aligned) landing on the same mbarrier as the words — and the words address x
// by its position in that staging buffer.  An entry is
//     LDS.32 word ; LEA.HI (staging address) ; LDS.64 x ; LOP3 (dictionary address) ; LDS.64 value ; DMUL ; DADD
// with shared-memory latencies only; the arithmetic per row is unchanged (ascending column order, separate multiply
// and add).  Per slice metadata: header {entry offset / 32, width, staged bytes, runs} and xs_rmax run slots
// {first column (even), doubles (even) | staging position << 16} (built at upload, spmv.cu).
//
// Measured at 129^3 where both streams apply (profiles/r02d, r02e): fused sweep 70.5 us vs 69.9 us for spmv1_vi_kernel —
// the x-gathers were NOT what bounds that kernel.  What does: with 8 / 10 / 16 warps per SM this kernel takes
// 110 / 92 / 70 us, time ~ 1 / warps — every warp spends ~2 us per slice on the per-row streams (defect, correction,
// inverse diagonal) it requests at the start of a slice and needs at its end.  Staging those through bulk copies too
// (one more variant, 15 copies of 128-3456 bytes per slice) was slower, 86-96 us: the copy engine, not HBM, became the
// limit.  Hence: value-indexed kernel where its window fits, this one where it does not (or UG4B200_XSTAGE=1).
#ifndef UG_XS_NST
#define UG_XS_NST 2
#endif
#ifndef UG_XS_WPB
#define UG_XS_WPB 8
#endif
#ifndef UG_XS_MINCTA
#define UG_XS_MINCTA 2
#endif
#ifndef UG_XS_XCAP
#define UG_XS_XCAP 384
#endif
struct XCfg {
	static constexpr int UB = 9, NB = 3, KC = UB * NB, NST = UG_XS_NST, WPB = UG_XS_WPB;
	static constexpr int XCAP = UG_XS_XCAP;              // doubles of x per slice
	static constexpr int RMAX = 32;                      // run slots per slice (one lane issues one run); lexicographic 27-point: 9, Cuthill-McKee: ~23
	static constexpr int DICT_MAX = 256;                 // dictionary entries (2 KB, at a 2 KB-aligned shared address)
	static constexpr int WORD_BYTES = KC * 128;
	static constexpr int STAGE_BYTES = WORD_BYTES + XCAP * 8;
	static constexpr int WARP_BYTES = NST * STAGE_BYTES;
	static constexpr int RING_BYTES = WPB * WARP_BYTES;
	static constexpr int BAR_BYTES = ((WPB * NST * 8 + 127) / 128) * 128;
	static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + 2 * DICT_MAX * 8;
};

template <int BETAK, int MODE, int FUSE>
__global__ void __launch_bounds__(XCfg::WPB * 32, UG_XS_MINCTA)
spmv1_xs_kernel(Sell A, double* dest, const double* v, double alpha, double beta, const double* __restrict__ w,
                Fuse fz, const int* guard)
{
	if (ug_guarded(guard)) return;
	typedef XCfg C;
	constexpr int UB = C::UB, NB = C::NB, NST = C::NST, XWPB = C::WPB;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t smem0 = smem_base_opaque(smem_raw);
	const uint32_t ring_s = smem0 + (uint32_t)wid * C::WARP_BYTES;             // [NST]{words [KC][32], x [XCAP]}
	const uint32_t bars = smem0 + C::RING_BYTES + (uint32_t)wid * NST * 8;     // [NST] mbarriers
	const uint32_t sdict_s = (smem0 + C::RING_BYTES + C::BAR_BYTES + 2047u) & ~2047u;
	const int gwarp = (int)blockIdx.x * XWPB + wid;
	const int nwarps = (int)gridDim.x * XWPB;
	const int nslices = (int)A.num_slices, nrows = (int)A.nrows;
	const int rmax = A.xs_rmax;
	for (int i = threadIdx.x; i < A.ndict; i += blockDim.x)
		asm volatile("st.shared.f64 [%0], %1;" ::"r"(sdict_s + (uint32_t)i * 8u), "d"(A.dict[i]) : "memory");
	__syncthreads();
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < NST; ++i) mbar_init(bars + i * 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	// ---- producer: the whole warp issues (lane 0: words, lane q < runs: run q); descriptors of the NEXT slice to
	// produce are requested while the current one is issued, so that no production starts with a global round trip
	struct Desc { int4 h; int2 r; };
	auto fetch_desc = [&](int sl) {
		Desc d; d.h = make_int4(0, 0, 0, 0); d.r = make_int2(0, 0);
		if (sl < nslices) {
			d.h = __ldg(A.xs_hdr + sl);
			if (lane < rmax) d.r = __ldg(A.xs_runs + (int64_t)sl * rmax + lane);
		}
		return d;
	};
	int ps = gwarp;
	Desc pd = fetch_desc(ps);
	// issue the slice described by pd into ring stage st; lane 0 posts the expected byte count BEFORE any copy is issued
	auto issue = [&](int st) {
		const uint32_t bar = bars + st * 8;
		const uint32_t dst = ring_s + (uint32_t)st * C::STAGE_BYTES;
		const int width = pd.h.y, xbytes = pd.h.z, nruns = pd.h.w;
		if (lane == 0) {
			if (width > 0) mbar_expect_tx(bar, (uint32_t)(width * 128 + xbytes));
			else mbar_arrive(bar);
		}
		__syncwarp();
		if (lane == 0 && width > 0) bulk_g2s(dst, A.xw + (int64_t)pd.h.x * 32, (uint32_t)width * 128u, bar);
		if (lane < nruns) {
			const int len = pd.r.y & 0xffff, pos = pd.r.y >> 16;
			bulk_g2s(dst + C::WORD_BYTES + (uint32_t)pos * 8u, w + pd.r.x, (uint32_t)len * 8u, bar);
		}
	};
	auto produce = [&](int st) {
		issue(st);
		ps += nwarps;
		pd = fetch_desc(ps);
	};
	{
		// prologue: the descriptors of the first NST slices are requested together (one round trip, not NST)
		Desc d1 = fetch_desc(gwarp + nwarps);
		if (ps < nslices) { issue(0); ps += nwarps; pd = d1; }
#pragma unroll
		for (int i = 1; i < NST; ++i) {
			if (ps < nslices) produce(i);
		}
	}

	// ---- consumer
	struct Meta { int s, width, len; double acc0; };
	auto fetch_meta = [&](int sl) {
		Meta m; m.s = sl; m.width = 0; m.len = 0; m.acc0 = 0.0;
		if (sl < nslices) {
			m.width = __ldg(&A.xs_hdr[sl].y);
			m.len = A.rowlen[sl * 32 + lane];
			if (MODE == MODE_INPLACE) { if (sl * 32 + lane < nrows) m.acc0 = dest[sl * 32 + lane]; }
			else if (MODE == MODE_GENERAL) { if (sl * 32 + lane < nrows) m.acc0 = v[sl * 32 + lane]; }
		}
		return m;
	};
	Meta cur = fetch_meta(gwarp);
	int stage = 0; uint32_t phase = 0;
	double dot = 0.0;
	const UgPushDev* push = (FUSE == FUSE_JACOBI || FUSE == FUSE_RESTRICT_JACOBI) ? fz.push : nullptr;
	const unsigned long long pe = push ? *(volatile unsigned long long*)push->epoch + 1ull : 0ull;
	while (cur.s < nslices) {
		const Meta nxt = fetch_meta(cur.s + nwarps);
		const int row = cur.s * 32 + lane;
		const bool live = row < nrows;
		const int len = cur.len;
		const int width = cur.width;
		const unsigned int pmask = push ? push->rowmask[cur.s] : 0u;
		const int minlen = __reduce_min_sync(0xffffffffu, len);
		double acc = 0.0, own = 0.0, scv = 0.0, dinv = 0.0;
		if (MODE == MODE_INPLACE) acc = cur.acc0;
		else if (MODE == MODE_GENERAL) acc = alpha * cur.acc0;
		// per-row streams: requested before the slice is consumed, used after it
		if (FUSE == FUSE_DOT) { if (live) own = w[row]; }
		if (FUSE == FUSE_RESTRICT_JACOBI && live) { dinv = fz.diaginv[row]; if (len == 0) own = dest[row]; }
		if (FUSE == FUSE_JACOBI && live) {
			if (fz.flags & UG4B200_SMOOTH_ADD_IN) own = w[row];
			if ((fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) && !(fz.flags & UG4B200_SMOOTH_SC_ZERO)) scv = fz.sc[row];
			if (fz.flags & UG4B200_SMOOTH_JACOBI) dinv = fz.diaginv[row];
		}
		if (width > 0) {
			mbar_wait(bars + stage * 8, phase);
			const uint32_t cs = ring_s + (uint32_t)stage * C::STAGE_BYTES + (uint32_t)lane * 4u;
			const uint32_t xs = ring_s + (uint32_t)stage * C::STAGE_BYTES + C::WORD_BYTES;
#pragma unroll
			for (int j = 0; j < NB; ++j) {
				const int kb = j * UB;
				if (kb < width) {
					if (kb + UB <= minlen) {
#pragma unroll
						for (int u = 0; u < UB; ++u) {
							const uint32_t e = lds_u32(cs + (uint32_t)(kb + u) * 128u);
							const double x = lds_f64(__umulhi(e, 1u << 19) + xs);
							const double t = mulbeta<BETAK>(lds_f64((e & 0x7f8u) | sdict_s), beta) * x;
							if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && j == 0 && u == 0) acc = t;
							else acc = acc + t;
						}
					} else {
#pragma unroll
						for (int u = 0; u < UB; ++u) {
							if (kb + u < len) {
								const uint32_t e = lds_u32(cs + (uint32_t)(kb + u) * 128u);
								const double x = lds_f64(__umulhi(e, 1u << 19) + xs);
								const double t = mulbeta<BETAK>(lds_f64((e & 0x7f8u) | sdict_s), beta) * x;
								if ((MODE == MODE_ASSIGN || MODE == MODE_ASSIGN_SKIP_EMPTY) && kb + u == 0) acc = t;
								else acc = acc + t;
							}
						}
					}
				}
			}
		} else {
			mbar_wait(bars + stage * 8, phase);   // empty slice: the producer arrived without a copy
		}
		// every lane is done with the stage: refill it
		__syncwarp();
		if (ps < nslices) produce(stage);
		if (++stage == NST) { stage = 0; phase ^= 1u; }
		if (FUSE == FUSE_JACOBI) {
			if (live) {
				dest[row] = acc;
				if (fz.flags & UG4B200_SMOOTH_ADD_IN) scv = scv + own;
				if (fz.flags & UG4B200_SMOOTH_JACOBI) {
					const double st = dinv * acc;
					fz.st_out[row] = st;
					if ((pmask >> lane) & 1u) ug_push_row(push, pe, cur.s, lane, pmask, st);
					if (fz.flags & UG4B200_SMOOTH_ADD_OUT) scv = scv + st;
				}
				if (fz.flags & (UG4B200_SMOOTH_ADD_IN | UG4B200_SMOOTH_ADD_OUT)) fz.sc[row] = scv;
			}
		} else if (FUSE == FUSE_RESTRICT_JACOBI) {
			if (live) {
				double dv = own;
				if (len > 0) { dest[row] = acc; dv = acc; }
				const double st = dinv * dv;
				fz.st_out[row] = st;
				if ((pmask >> lane) & 1u) ug_push_row(push, pe, cur.s, lane, pmask, st);
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

// limits of the x-staged format, for the builder in spmv.cu
inline int tma_xs_max_dict() { return tma::XCfg::DICT_MAX; }
inline int tma_xs_max_width() { return tma::XCfg::KC; }
inline int tma_xs_max_runs() { return tma::XCfg::RMAX; }
inline int tma_xs_max_doubles() { return tma::XCfg::XCAP; }
