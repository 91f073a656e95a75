// scan_common.cuh — device helpers shared by the scan and probe kernels (scan_kernel.cu, probe_dense_kernel.cu):
// PTX glue (mbarrier, bulk async copy, relaxed status words), 128-bit accumulation, the prefix
// warp's chain-free look-back, and the staged, position-ordered emission (stage_word / write_out /
// emit_span).  See scan_kernel.cu for the algorithm description and the reference citations.
#pragma once
#include "col_ref.cuh"
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {


// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t addr = smem_u32(bar);
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		             "selp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done)
		             : "r"(addr), "r"(parity)
		             : "memory");
	} while (!done);
}
// the same for the producer and prefix warps, whose waits are long and not latency-critical: a bare try_wait
// loop re-issues every few cycles and takes issue slots from the consumer warps of the same scheduler
// (16 % of all issued instructions at s = 0.1, profiles/r1_ncu_scan_s01.md)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, unsigned ns) {
	uint32_t addr = smem_u32(bar);
	uint32_t done;
	while (true) {
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		             "selp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done)
		             : "r"(addr), "r"(parity)
		             : "memory");
		if (done) {
			return;
		}
		__nanosleep(ns);
	}
}
// shared-window addresses computed ONCE per thread (the generic → shared conversion of a dynamic shared-memory
// pointer costs an S2R + LEA every time the compiler re-derives it)
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		             "selp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done)
		             : "r"(addr), "r"(parity)
		             : "memory");
	} while (!done);
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t addr) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
	uint64_t v;
	asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
	return v;
}
// 1-D bulk async copy global → shared, completion on an mbarrier (TMA engine; SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void consumer_bar_sync() {
	asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// look-back status word: [63:62] flag (0 = not published yet), [61:0] popcount of the segment
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;
constexpr uint32_t kNoTile = 0xffffffffu;

struct StageMeta {
	uint32_t tile; // kNoTile = no more work
	uint32_t d0;   // first delta entry of (stream, segment)
	uint32_t dcnt; // number of delta entries
	uint32_t pad;
};

// 128-bit signed accumulate of an int64 (AddToHugeint::AddValue, sum_helpers.hpp:92-113)
__device__ __forceinline__ void add128(unsigned long long &lo, long long &hi, long long v) {
	unsigned long long uv = (unsigned long long)v;
	lo += uv;
	hi += (long long)(lo < uv) + (v >> 63);
}
__device__ __forceinline__ void add128(unsigned long long &lo, long long &hi, unsigned long long lo2, long long hi2) {
	lo += lo2;
	hi += hi2 + (long long)(lo < lo2);
}

#ifndef CUBIT_PREFIX_DELAY_NS
#define CUBIT_PREFIX_DELAY_NS 0
#endif
#ifndef CUBIT_SMEM_ADDR
#define CUBIT_SMEM_ADDR 1
#endif
#ifndef CUBIT_STAGE_ASM
#define CUBIT_STAGE_ASM 1
#endif
#ifndef CUBIT_PRODUCER_SLEEP_NS
#define CUBIT_PRODUCER_SLEEP_NS 0
#endif
#ifndef CUBIT_PREFIX_RETRY_NS
#define CUBIT_PREFIX_RETRY_NS 64
#endif
#ifndef CUBIT_DEFER
#define CUBIT_DEFER 2
#endif
#ifndef CUBIT_REQ_SLOTS
#define CUBIT_REQ_SLOTS 4
#endif
constexpr int kReqSlots = CUBIT_REQ_SLOTS; // look-back requests in flight per CTA (> emission deferral depth)
constexpr int kDefer = CUBIT_DEFER;        // segments merged between a segment's merge and its emission
#ifndef CUBIT_PREFIX_LAG
#define CUBIT_PREFIX_LAG (CUBIT_DEFER - 1)
#endif
constexpr int kPrefixLag = CUBIT_PREFIX_LAG; // requests posted after request n before the prefix warp serves n
static_assert(kPrefixLag >= 0 && kPrefixLag < kDefer && kReqSlots > kDefer, "look-back pipeline depth");
#ifndef CUBIT_WAIT_BATCH
#define CUBIT_WAIT_BATCH 4
#endif
constexpr int kWaitBatch = CUBIT_WAIT_BATCH; // ring stages consumed per mbarrier round trip
constexpr int kDeltaStage = 32;    // delta words staged in shared memory per ring stage (the rest is read from L2)
constexpr int kSlotRows = 32 * 64; // rows one warp compacts at a time (32 lanes × one 64-bit word)
// per-warp staging buffer (uint16 units): [0, kSlotRows + 8) row numbers, + 32 per-lane dummy slots, then the
// pack-block headers of up to kMaxFusedCols bit-packed columns (16 × 16 bytes each, 16-byte aligned)
constexpr int kCompactHdrOff = kSlotRows + 8 + 32;

constexpr int kLookSlots = 16; // status words per prefix-warp lane per window → 512 segments per window

// ---- the prefix warp's look-back, off the consumers' path.
// status[t] = kFlagAgg | popcount(segment t), published once by the CTA that merged t.
// The warp remembers the last segment this CTA handled (t0) and the inclusive prefix through
// it (s0); the exclusive prefix of the next one is s0 + Σ status(t0+1 .. t-1).  It depends
// only on those segments having been MERGED (inherent) — never on another CTA's look-back,
// so there are no prefix chains, and the usual gap (≈ number of resident CTAs) is one window.
__device__ __forceinline__ unsigned long long sum_aggregates(const unsigned long long *status, int64_t lo, int64_t hi,
                                                             int lane) {
	unsigned long long acc = 0; // per-lane partial
	for (int64_t w0 = lo; w0 < hi; w0 += kLookSlots * 32) {
		unsigned pending = 0; // bit j: slot j of this lane still unpublished
#pragma unroll
		for (int j = 0; j < kLookSlots; j++) {
			if (w0 + j * 32 + lane < hi) {
				pending |= 1u << j;
			}
		}
		while (__any_sync(0xffffffffu, pending != 0)) {
			unsigned long long sv[kLookSlots];
#pragma unroll
			for (int j = 0; j < kLookSlots; j++) {
				sv[j] = (pending >> j) & 1u ? ld_relaxed_u64(&status[w0 + j * 32 + lane]) : 0ull;
			}
#pragma unroll
			for (int j = 0; j < kLookSlots; j++) {
				if (sv[j] >> 62) {
					acc += sv[j] & kValMask;
					pending &= ~(1u << j);
				}
			}
			if (__any_sync(0xffffffffu, pending != 0)) {
				__nanosleep(CUBIT_PREFIX_RETRY_NS); // some predecessor is still being merged: re-read only the holes
			}
		}
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		acc += __shfl_xor_sync(0xffffffffu, acc, d);
	}
	return acc;
}

// per-lane aggregate state of the fused / bit-driven probe
struct Agg {
	unsigned long long lo = 0; // 128-bit integer SUM, low limb
	long long hi = 0;          //                      high limb
	double f = 0.0;            // SUM over a DOUBLE column (CUBIT_AGG_SUM_F64)
	unsigned int overflow = 0;
};

// warp-reduce an Agg and add it to the result header: exact for the integer limbs (the carry out of the
// low limb is recovered from the value the atomic returns), plain atomicAdd for the double
__device__ __forceinline__ void agg_flush_warp(Agg &g, ResultHeader *hdr, int lane) {
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const unsigned long long olo = __shfl_xor_sync(0xffffffffu, g.lo, d);
		const long long ohi = __shfl_xor_sync(0xffffffffu, g.hi, d);
		add128(g.lo, g.hi, olo, ohi);
		g.f += __shfl_xor_sync(0xffffffffu, g.f, d);
		g.overflow |= __shfl_xor_sync(0xffffffffu, g.overflow, d);
	}
	if (lane == 0) {
		if (g.lo | (unsigned long long)g.hi) {
			const unsigned long long old = atomicAdd(&hdr->sum_lo, g.lo);
			const long long carry = (old + g.lo) < old ? 1 : 0;
			atomicAdd(reinterpret_cast<unsigned long long *>(&hdr->sum_hi), (unsigned long long)(g.hi + carry));
		}
		if (g.f != 0.0) {
			atomicAdd(&hdr->sum_f64, g.f);
		}
		if (g.overflow) {
			atomicOr(&hdr->overflow, 1u);
		}
	}
}

// one selected row: store its id / gathered values at output position `pos`, accumulate
template <int NL, bool POS>
__device__ __forceinline__ void consume_row(const ScanArgs &a, unsigned long long pos, int64_t rid,
                                            const long long (&v)[NL > 0 ? NL : 1], Agg &agg) {
	if (POS) {
		if (a.ids_out) {
			__stcs(a.ids_out + pos, (long long)rid);
		}
#pragma unroll
		for (int c = 0; c < NL; c++) {
			if (a.lout[c]) {
				__stcs(a.lout[c] + pos, v[c]);
			}
		}
	}
	if (NL > 0) {
		const long long x = (NL > 1 && a.agg_ia == 1) ? v[NL - 1] : v[0];
		if (a.agg_kind == 1) {
			add128(agg.lo, agg.hi, x);
		} else if (a.agg_kind == 2) {
			const long long y = (NL > 1 && a.agg_ib == 1) ? v[NL - 1] : v[0];
			const long long pr = x * y;
			if (__mul64hi(x, y) != (pr >> 63)) {
				agg.overflow = 1;
			}
			add128(agg.lo, agg.hi, pr);
		} else if (a.agg_kind == 3) {
			agg.f += __longlong_as_double(x);
		}
	}
}

// Position-ordered write-out of `count` staged rows (16-bit row numbers relative to
// row_origin, staged at cbuf[pad ..), pad = first output position & 1).  Per iteration the
// warp writes 128 consecutive results as two fully contiguous 512-byte stores (lane l:
// pairs l and l+32), gathers the fused-probe columns for them first (independent loads in
// flight) and accumulates the aggregates.
template <int NL, bool POS, bool PK = true>
__device__ __forceinline__ void write_out(const ScanArgs &a, const uint16_t *cbuf, uint32_t pad, uint32_t count,
                                          unsigned long long pos0, int64_t row_origin, int lane,
                                          Agg &agg) {
	const int64_t local0 = row_origin - a.row_base;
	const unsigned long long obase = pos0 - pad; // output position of staging index 0 (even)
	const uint32_t end = pad + count;
	const uint32_t *cb32 = reinterpret_cast<const uint32_t *>(cbuf);
	// pack-block headers of the probed columns, parked behind the staging area (kCompactHdrOff)
	uint4 *hs = reinterpret_cast<uint4 *>(const_cast<uint16_t *>(cbuf) + kCompactHdrOff);
	if (NL > 0 && PK) { // PK = false: every probed column is a raw array, nothing to stage
#pragma unroll
		for (int cc = 0; cc < NL; cc++) {
			stage_hdrs(a.lcol[cc], local0, lane, hs + cc * kHdrSlots);
		}
		__syncwarp();
	}
	for (uint32_t g0 = 0; g0 * 2 < end; g0 += 64) {
		uint32_t r[2][2];
		bool ok[2][2];
		long long v[2][2][NL > 0 ? NL : 1];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const uint32_t g = g0 + h * 32 + lane; // pair index
			const uint32_t packed = g * 2 < end ? cb32[g] : 0u;
			r[h][0] = packed & 0xffffu;
			r[h][1] = packed >> 16;
			ok[h][0] = g * 2 >= pad && g * 2 < end;
			ok[h][1] = g * 2 + 1 < end;
			if (NL > 0) {
#pragma unroll
				for (int e = 0; e < 2; e++) {
#pragma unroll
					for (int cc = 0; cc < NL; cc++) {
						if (PK) {
							v[h][e][cc] = load_col_staged(a.lcol[cc], hs + cc * kHdrSlots, local0, r[h][e], ok[h][e]);
						} else {
							v[h][e][cc] = ok[h][e] ? __ldg(a.lcol[cc].raw + local0 + r[h][e]) : 0;
						}
					}
				}
			}
		}
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const uint32_t g = g0 + h * 32 + lane;
			if (POS && ok[h][0] && ok[h][1]) {
				if (a.ids_out) {
					__stcs(reinterpret_cast<longlong2 *>(a.ids_out + obase + g * 2),
					       make_longlong2(row_origin + r[h][0], row_origin + r[h][1]));
				}
#pragma unroll
				for (int cc = 0; cc < NL; cc++) {
					if (a.lout[cc]) {
						__stcs(reinterpret_cast<longlong2 *>(a.lout[cc] + obase + g * 2),
						       make_longlong2(v[h][0][cc], v[h][1][cc]));
					}
				}
				if (NL > 0) {
					consume_row<NL, false>(a, 0, 0, v[h][0], agg);
					consume_row<NL, false>(a, 0, 0, v[h][1], agg);
				}
			} else {
#pragma unroll
				for (int e = 0; e < 2; e++) {
					if (ok[h][e]) {
						consume_row<NL, POS>(a, obase + g * 2 + e, row_origin + r[h][e], v[h][e], agg);
					}
				}
			}
		}
	}
}

// lane-local compaction of one 64-bit word: its two halves are two independent ctz chains.
// Branch-free: an exhausted chain keeps "storing" into a per-lane dummy slot behind the
// staging area, so the loop body is straight-line code (no divergence regions).
// (hi_off: row distance between the two halves — 32 for the halves of a 64-bit word, 16 when a 32-bit piece is split)
__device__ __forceinline__ void stage_word(uint16_t *cbuf, uint32_t p0, uint32_t wlo, uint32_t whi, uint32_t bit0,
                                           uint32_t dummy, uint32_t hi_off = 32u) {
#if !CUBIT_STAGE_ASM
	uint32_t p1 = p0 + __popc(wlo);
	uint32_t w0 = wlo, w1 = whi;
	const uint32_t b1 = bit0 + hi_off;
	while (w0 | w1) {
		const uint32_t i0 = w0 ? p0 : dummy, i1 = w1 ? p1 : dummy;
		cbuf[i0] = (uint16_t)(bit0 + (uint32_t)(__ffs(w0) - 1));
		cbuf[i1] = (uint16_t)(b1 + (uint32_t)(__ffs(w1) - 1));
		p0 += w0 != 0;
		p1 += w1 != 0;
		w0 &= w0 - 1; // 0 stays 0
		w1 &= w1 - 1;
	}
#else
	(void)dummy;
	// running shared-memory byte addresses of the two chains; one step = ctz, predicated 16-bit store,
	// predicated address bump, clear the lowest set bit (8 SASS instructions per chain, no select / re-derived
	// address as the compiler generates for the C form)
	uint32_t a0 = smem_u32(cbuf + p0);
	uint32_t a1 = a0 + 2u * (uint32_t)__popc(wlo);
	uint32_t w0 = wlo, w1 = whi;
	const uint32_t b1 = bit0 + hi_off;
	while (w0 | w1) {
		asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b32 f, g, t;\n\t.reg .b16 h;\n\t"
		             "setp.ne.u32 p, %2, 0;\n\t"
		             "setp.ne.u32 q, %3, 0;\n\t"
		             "brev.b32 f, %2;\n\t"
		             "brev.b32 g, %3;\n\t"
		             "clz.b32 f, f;\n\t"
		             "clz.b32 g, g;\n\t"
		             "add.u32 f, f, %4;\n\t"
		             "add.u32 g, g, %5;\n\t"
		             "cvt.u16.u32 h, f;\n\t"
		             "@p st.shared.u16 [%0], h;\n\t"
		             "cvt.u16.u32 h, g;\n\t"
		             "@q st.shared.u16 [%1], h;\n\t"
		             "@p add.u32 %0, %0, 2;\n\t"
		             "@q add.u32 %1, %1, 2;\n\t"
		             "add.u32 t, %2, -1;\n\t"
		             "and.b32 %2, %2, t;\n\t"
		             "add.u32 t, %3, -1;\n\t"
		             "and.b32 %3, %3, t;\n\t}"
		             : "+r"(a0), "+r"(a1), "+r"(w0), "+r"(w1)
		             : "r"(bit0), "r"(b1)
		             : "memory");
	}
#endif
}

// ---- emission of one warp's span of a merged segment.
// q[i] of lane l is word (i*32 + l) of the span, so slot i = 2048 consecutive rows and the
// output order is (slot, lane, bit).
template <int WPT, int NL, bool POS, bool PK = true>
__device__ __forceinline__ void emit_span(const ScanArgs &a, const uint64_t (&q)[WPT], uint16_t *cbuf,
                                          unsigned long long wbase, int64_t span_row0, int lane,
                                          Agg &agg) {
	uint32_t c[WPT], incl[WPT], lane_total = 0;
#pragma unroll
	for (int i = 0; i < WPT; i++) {
		c[i] = __popcll(q[i]);
		incl[i] = c[i];
		lane_total += c[i];
	}
	const uint32_t span_total = __reduce_add_sync(0xffffffffu, lane_total); // redux.sync: one instruction
	if (span_total == 0) {
		return; // sparse selections: most spans are empty
	}
	// WPT independent warp scans, interleaved (one scan's latency for all slots)
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
		for (int i = 0; i < WPT; i++) {
			const uint32_t n = __shfl_up_sync(0xffffffffu, incl[i], d);
			if (lane >= d) {
				incl[i] += n;
			}
		}
	}
	uint32_t slot_total[WPT];
#pragma unroll
	for (int i = 0; i < WPT; i++) {
		slot_total[i] = __shfl_sync(0xffffffffu, incl[i], 31);
	}
	const uint32_t dummy = (uint32_t)(kSlotRows + 8 + lane); // per-lane scratch slot behind the staging area
	if (span_total <= (uint32_t)kSlotRows && WPT * kSlotRows <= 65536) {
		// sparse / medium span: stage ALL slots at once (row numbers relative to the span fit
		// 16 bits), then one write-out — one synchronisation round instead of one per slot
		const uint32_t pad = (uint32_t)wbase & 1u;
		uint32_t base = pad;
#pragma unroll
		for (int i = 0; i < WPT; i++) {
			stage_word(cbuf, base + incl[i] - c[i], (uint32_t)q[i], (uint32_t)(q[i] >> 32),
			           (uint32_t)(i * kSlotRows + lane * 64), dummy);
			base += slot_total[i];
		}
		__syncwarp();
		write_out<NL, POS, PK>(a, cbuf, pad, span_total, wbase, span_row0, lane, agg);
		__syncwarp();
		return;
	}
	unsigned long long pos0 = wbase; // output position of the slot's first selected row
#pragma unroll
	for (int i = 0; i < WPT; i++) {
		if (slot_total[i] == 0) {
			continue;
		}
		const uint32_t pad = (uint32_t)pos0 & 1u;
		stage_word(cbuf, pad + incl[i] - c[i], (uint32_t)q[i], (uint32_t)(q[i] >> 32), (uint32_t)lane * 64u, dummy);
		__syncwarp();
		write_out<NL, POS, PK>(a, cbuf, pad, slot_total[i], pos0, span_row0 + (int64_t)i * kSlotRows, lane, agg);
		__syncwarp();
		pos0 += slot_total[i];
	}
}

// ---- the same emission cut into steps, so that the scan kernel can interleave the write-out of a pending
// segment with the fold of the next one (the ring keeps draining while a dense segment is written out).
// sparse / medium span: step 0 stages every slot, step 1 writes out; dense span: one step per slot.
constexpr uint32_t kEmitDone = 0xffu;
struct EmitState {
	uint32_t step = kEmitDone; // next step, kEmitDone = nothing (left) to do
	uint32_t span_total = 0;   // selected rows of the span
	unsigned long long pos0 = 0; // output position of the next row to emit
};

template <int WPT>
__device__ __forceinline__ void emit_begin(const uint64_t (&q)[WPT], unsigned long long wbase, EmitState &es) {
	uint32_t lane_total = 0;
#pragma unroll
	for (int i = 0; i < WPT; i++) {
		lane_total += __popcll(q[i]);
	}
	es.span_total = __reduce_add_sync(0xffffffffu, lane_total);
	es.pos0 = wbase;
	es.step = es.span_total ? 0u : kEmitDone;
}

template <int WPT, int NL, bool POS>
__device__ __forceinline__ void emit_step(const ScanArgs &a, const uint64_t (&q)[WPT], uint16_t *cbuf,
                                          int64_t span_row0, int lane, Agg &agg, EmitState &es) {
	const uint32_t dummy = (uint32_t)(kSlotRows + 8 + lane);
	if (es.span_total <= (uint32_t)kSlotRows && WPT * kSlotRows <= 65536) {
		const uint32_t pad = (uint32_t)es.pos0 & 1u;
		if (es.step == 0) {
			uint32_t c[WPT], incl[WPT];
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				c[i] = __popcll(q[i]);
				incl[i] = c[i];
			}
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
				for (int i = 0; i < WPT; i++) {
					const uint32_t n = __shfl_up_sync(0xffffffffu, incl[i], d);
					if (lane >= d) {
						incl[i] += n;
					}
				}
			}
			uint32_t base = pad;
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				stage_word(cbuf, base + incl[i] - c[i], (uint32_t)q[i], (uint32_t)(q[i] >> 32),
				           (uint32_t)(i * kSlotRows + lane * 64), dummy);
				base += __shfl_sync(0xffffffffu, incl[i], 31);
			}
			__syncwarp();
			es.step = 1;
		} else {
			write_out<NL, POS>(a, cbuf, pad, es.span_total, es.pos0, span_row0, lane, agg);
			__syncwarp();
			es.step = kEmitDone;
		}
		return;
	}
#pragma unroll
	for (int i = 0; i < WPT; i++) {
		if (es.step == (uint32_t)i) {
			const uint32_t c = __popcll(q[i]);
			uint32_t incl = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) {
					incl += n;
				}
			}
			const uint32_t slot_total = __shfl_sync(0xffffffffu, incl, 31);
			if (slot_total) {
				const uint32_t pad = (uint32_t)es.pos0 & 1u;
				stage_word(cbuf, pad + incl - c, (uint32_t)q[i], (uint32_t)(q[i] >> 32), (uint32_t)lane * 64u, dummy);
				__syncwarp();
				write_out<NL, POS>(a, cbuf, pad, slot_total, es.pos0, span_row0 + (int64_t)i * kSlotRows, lane, agg);
				__syncwarp();
				es.pos0 += slot_total;
			}
		}
	}
	es.step = es.step + 1 == (uint32_t)WPT ? kEmitDone : es.step + 1;
}

// WPT: 64-bit words of Q each consumer thread holds → segment = 256*WPT words
//      (WPT 2/4/8 ↔ 32768/65536/131072 rows per segment).
// NL : distinct int64 columns gathered at every selected row (fused probe).
} // namespace cubit
