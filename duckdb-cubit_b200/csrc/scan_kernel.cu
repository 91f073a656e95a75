// scan_kernel.cu — the fused CUBIT segment-merge + bit→row-ID decode (+ probe /
// aggregate) kernel for sm_100a.
//
// What it computes:
//     Q = AND_j ( OR_{i in R_j} ( B_i XOR D_i ) )          per fixed-size segment
//     ids = ascending positions of the set bits of Q (+ row_base)
//     vals[c] = col_c[ids], SUM / SUM(a*b) over them        (optional, fused)
// The reference tree has no bitmap code; conventions come from
//   bit order      src/include/duckdb/common/types/validity_mask.hpp:163-168
//   row-id order   src/execution/index/art/art.cpp:974-985 (sorted, unique)
//   SUM carry      src/include/duckdb/core_functions/aggregate/sum_helpers.hpp:92-113
//
// Shape of the kernel (B200-first, HBM-bound integer work, no tensor cores):
//   * TWO persistent CTAs per SM (8 consumer warps + 1 producer warp + 1 prefix
//     warp each); segments are handed out by a ticket counter, ONE segment per
//     ticket, so segments are merged in roughly global order (the look-back relies
//     on it) and every predecessor of a segment is already owned by a running CTA.
//     Tickets are drawn 2..8 segments ahead so the atomic never gates a copy.
//   * the producer warp streams whole segments of every queried bitvector with
//     1-D bulk async copies (cp.async.bulk → SASS UBLKCP, the TMA engine) into a
//     72 KiB shared-memory ring guarded by mbarriers; it runs ahead of the
//     consumers by the ring depth, which keeps the HBM pipe full (≥ 64 KiB in
//     flight per SM) while consumers decode and emit.  Consumers wait for up to
//     four ring stages with ONE SIMT try_wait (lane u polls stage u), so the
//     mbarrier round trip is paid once per four streams, not once per stream.
//   * consumer warps: apply the (sparse) pending-delta words of the staged
//     segment (XOR), fold the segment into registers (OR within a group, AND
//     across groups; a single OR group — the common range predicate — has its own
//     leaner fold) with conflict-free ld.shared.
//   * decode: __popcll, warp redux, block scan over the 8 warp totals; the
//     segment's aggregate is published at once.  The inter-segment prefix is a
//     single-pass decoupled look-back done by the PREFIX WARP, asynchronously (it
//     sums the published aggregates between this CTA's previous segment and the
//     new one — no chains through other CTAs' look-backs): consumers hand it
//     (segment, count) through shared memory + mbarrier and pick the answer up two
//     segments later.  The prefix warp serves a request one segment-time after it
//     was posted (when the next request arrives): by then the predecessors have
//     published and ONE round of status loads suffices — polling for them at once
//     cost the bulk-copy stream 12-15 % in L2 traffic.
//   * emit: the oldest pending segment is written out WHILE the current one is
//     merged (one emission step after every batch of ring stages, the rest after
//     the current aggregate is published), so the ring keeps draining.  Every warp
//     compacts the set bits of 2048 consecutive rows (or its whole 8192-row span
//     when sparse) into a private shared-memory staging buffer (16-bit local row
//     numbers; each lane walks the two 32-bit halves of its word as two
//     independent ctz chains with predicated stores, no atomics) and then writes
//     row IDs (and gathers/stores column values, accumulates SUMs) in position
//     order: 128 consecutive results per iteration as two contiguous 512-byte
//     stores.
#include "scan_common.cuh"

namespace cubit {

// NL > 0 (probe fused into the scan, not the default path): the per-warp staging rows also hold the pack-block
// headers of the probed columns; the ring gives up 8 KiB so that two CTAs still fit one SM.
// CMP (a queried index keeps roaring-style containers, container_kernels.cu): the ring gives up a quarter of its
// bytes for the per-stage ARRAY staging buffers (kArrayMax 16-bit positions each).
template <int WPT, int NL, bool CMP = false>
struct ScanSmem {
	static constexpr int kTileWords = kConsumerThreads * WPT;
	static constexpr int kTileBytes = kTileWords * 8;
	static constexpr int kStages = CMP ? (kScanRingBytes * 3 / 4) / kTileBytes : (kScanRingBytes - (NL > 0 ? 8192 : 0)) / kTileBytes;
	static constexpr int kCompactRow = kCompactHdrOff + NL * kHdrSlots * 8;
	alignas(128) uint64_t stage[kStages][kTileWords];
	alignas(16) uint16_t compact[kConsumerWarps][kCompactRow]; // per-warp staging of local row numbers (+ dummy slots)
	alignas(16) DeltaEnt dbuf[kStages][kDeltaStage];              // pending-delta words staged beside each segment
	alignas(16) uint16_t abuf[CMP ? kStages : 1][CMP ? kArrayMax : 8]; // ARRAY containers staged beside each segment
	// ARRAY containers without pending deltas are OR-ed into this tile-sized bitmap by ONE consumer warp per stage
	// (a 32-bit shared-memory atomic per list entry) and folded into the registers once per OR group — instead of all
	// eight warps filtering every list for their span (2,650 warp instructions per container stage, the whole run time
	// of a day-level index scan; profiles/r2_compressed_scan.md)
	alignas(16) uint64_t acc[CMP ? kTileWords : 1];
	unsigned long long pdir[CMP ? kMaxStreams : 1];                    // producer: directory entries of the current segment
	alignas(8) uint64_t full[kStages];
	uint64_t empty[kStages];
	uint64_t req_full[kReqSlots];  // consumers → prefix warp
	uint64_t resp_full[kReqSlots]; // prefix warp → consumers
	unsigned long long resp_excl[kReqSlots];
	uint32_t req_tile[kReqSlots];
	uint32_t req_total[kReqSlots];
	StageMeta meta[kStages];
	uint32_t warp_tot[2][kConsumerWarps];
	uint32_t poff[2 * kMaxStreams]; // producer: delta CSR offsets of the current segment (lo | hi)
	BlockPartial red[kConsumerWarps];
};

// ONEG: the predicate is ONE OR group (a range / IN predicate on one indexed column, the common case): the fold
// is a plain OR into q, without the per-stream group test and the AND / reset of the group accumulator (23 of the
// 47 SASS instructions the generic fold spends per stream and warp).
template <int WPT, bool HAS_DELTA, int NL, bool ONEG, bool CMP>
__global__ void __launch_bounds__(kScanThreads, 2) cubit_scan_kernel(const __grid_constant__ ScanArgs a) {
	using Smem = ScanSmem<WPT, NL, CMP>;
	constexpr int kStages = Smem::kStages;
	constexpr int kTileWords = Smem::kTileWords;
	constexpr int kTileBytes = Smem::kTileBytes;
	constexpr int kSpanWords = WPT * 32; // words of one warp's span
	constexpr int UB = (WPT >= 8 || CMP) ? 2 : (kWaitBatch < kStages ? kWaitBatch : kStages); // (CMP: container stages are handshake-bound — shorter batches let the producer run further ahead)
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	unsigned int *ticket = reinterpret_cast<unsigned int *>(a.ctrl);
	unsigned long long *status = a.ctrl + 1;
	const bool need_pos =
	    (a.ids_out != nullptr) || (NL > 0 && (a.lout[0] != nullptr || (NL > 1 && a.lout[NL - 1] != nullptr)));
	const bool need_emit = need_pos || (NL > 0 && a.agg_kind != 0);

	if (threadIdx.x == 0) {
		for (int s = 0; s < kStages; s++) {
			mbar_init(&sm.full[s], 1);
			mbar_init(&sm.empty[s], kConsumerWarps);
		}
		for (int s = 0; s < kReqSlots; s++) {
			mbar_init(&sm.req_full[s], 1);
			mbar_init(&sm.resp_full[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (CMP) {
		for (int w = threadIdx.x; w < kTileWords; w += kScanThreads) {
			sm.acc[w] = 0;
		}
	}
	__syncthreads();

	if (warp == kConsumerWarps) {
		// ------------------------------------------------------------ producer warp
		// Lane 0 walks the ring and issues the bulk copies.  With pending deltas the other lanes
		// fetch the delta CSR offsets of the NEXT segment (its ticket is known one segment ahead)
		// while lane 0 is busy, so those dependent loads never sit in front of a bulk copy.
		// Tickets (one segment each, so segments are merged in roughly global order, which the look-back
		// relies on) are drawn `a.ticket_depth` segments ahead: lane j of this warp holds the ticket of
		// iteration i ≡ j (mod depth) and re-draws right after it was consumed, so the atomic's round trip
		// to L2 (≈ 1 µs under load) overlaps `depth` segments of bulk copies instead of gating each one —
		// with few bitvectors per query a segment is only one or two 8 KiB copies.
		uint32_t stage = 0, phase = 0;
		const uint32_t td = a.ticket_depth; // 2..8
		// the first `depth` rounds are assigned statically (round j: segment j·grid + CTA) — drawing them
		// with one warp-wide atomic would hand each CTA `depth` CONSECUTIVE segments, which serialises the
		// look-back; the counter then continues from depth·grid
		const uint32_t tbase = td * gridDim.x;
		uint32_t my_ticket = (uint32_t)lane * gridDim.x + blockIdx.x;
		uint32_t tslot = 0; // lane holding the current iteration's ticket
		uint32_t tile = __shfl_sync(0xffffffffu, my_ticket, 0);
		uint32_t dlo[2] = {0, 0}, dhi[2] = {0, 0}; // CSR offsets of streams lane and lane+32 for `tile`
		auto load_offsets = [&](uint32_t tl) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const uint32_t s = (uint32_t)lane + 32u * h;
				dlo[h] = dhi[h] = 0;
				if (s < a.k && a.doff[s] && tl < a.n_seg) {
					dlo[h] = __ldg(a.doff[s] + tl);
					dhi[h] = __ldg(a.doff[s] + tl + 1);
				}
			}
		};
		// compressed streams: the directory entry of (stream, segment) says what to copy; fetched like the delta
		// offsets, one segment ahead by all lanes, so the dependent load never sits in front of a bulk copy
		unsigned long long cde[2] = {0, 0};
		auto load_dirs = [&](uint32_t tl) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const uint32_t s = (uint32_t)lane + 32u * h;
				cde[h] = 0;
				if (s < a.k && a.cdir[s] && tl < a.n_seg) {
					cde[h] = __ldg(a.cdir[s] + tl);
				}
			}
		};
		if (HAS_DELTA) {
			load_offsets(tile);
		}
		if (CMP) {
			load_dirs(tile);
		}
		while (true) {
			const bool valid = tile < a.n_seg;
			if (valid && lane == (int)tslot) {
				my_ticket = tbase + atomicAdd(ticket, 1u); // consumed `depth` iterations from now
			}
			tslot = tslot + 1 == td ? 0 : tslot + 1;
			const uint32_t next = __shfl_sync(0xffffffffu, my_ticket, tslot); // drawn depth-1 iterations ago
			if (HAS_DELTA) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					sm.poff[lane + 32 * h] = dlo[h];
					sm.poff[64 + lane + 32 * h] = dhi[h];
				}
				__syncwarp();
				if (valid) {
					load_offsets(next); // in flight while lane 0 issues this segment's copies
				}
			}
			if (CMP) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					sm.pdir[lane + 32 * h] = cde[h];
				}
				__syncwarp();
				if (valid) {
					load_dirs(next);
				}
			}
			if (lane == 0) {
				for (uint32_t s = 0; s < a.k; s++) {
#if CUBIT_PRODUCER_SLEEP_NS
					mbar_wait_sleep(&sm.empty[stage], phase ^ 1, CUBIT_PRODUCER_SLEEP_NS);
#else
					mbar_wait(&sm.empty[stage], phase ^ 1);
#endif
					sm.meta[stage].tile = valid ? tile : kNoTile;
					if (!valid) {
						mbar_arrive(&sm.full[stage]); // end marker: k empty stages, so batched waits stay uniform
					} else {
						uint32_t dbytes = 0;
						const DeltaEnt *dsrc = nullptr;
						if (HAS_DELTA) {
							uint32_t d0 = sm.poff[s], d1 = sm.poff[64 + s];
							if (a.debug & 64u) { // timing experiment: pretend there are no pending deltas
								d1 = d0;
							}
							sm.meta[stage].d0 = d0;
							sm.meta[stage].dcnt = d1 - d0;
							if (d1 > d0) { // stage the first kDeltaStage delta words next to the segment
								dbytes = ((d1 - d0) < (uint32_t)kDeltaStage ? (d1 - d0) : (uint32_t)kDeltaStage) *
								         (uint32_t)sizeof(DeltaEnt);
								dsrc = a.dent[s] + d0;
							}
						}
						if (CMP && a.cdir[s]) {
							// container of (stream, segment): BITMAP → the ring stage, ARRAY → its staging buffer,
							// EMPTY / FULL → nothing crosses HBM at all
							const unsigned long long de = sm.pdir[s];
							const uint32_t ctype = ct_type(de), ccnt = ct_count(de);
							sm.meta[stage].pad = ctype | (ccnt << 2);
							const uint8_t *csrc = reinterpret_cast<const uint8_t *>(a.bv[s]) + ct_offset(de);
							uint32_t cbytes =
							    ctype == CT_BITMAP ? (uint32_t)kTileBytes : (ctype == CT_ARRAY ? ((ccnt * 2u + 15u) & ~15u) : 0u);
							if (ctype == CT_ARRAY && (a.debug & 16u)) {
								cbytes = kArrayMax * 2; // experiment: always copy a whole staging row
							}
							mbar_arrive_expect_tx(&sm.full[stage], cbytes + dbytes);
							if (ctype == CT_BITMAP) {
								bulk_g2s(&sm.stage[stage][0], csrc, kTileBytes, &sm.full[stage]);
							} else if (ctype == CT_ARRAY) {
								bulk_g2s(&sm.abuf[stage][0], csrc, cbytes, &sm.full[stage]);
							}
						} else {
							if (CMP) {
								sm.meta[stage].pad = CT_BITMAP;
							}
							mbar_arrive_expect_tx(&sm.full[stage], kTileBytes + dbytes);
							bulk_g2s(&sm.stage[stage][0], a.bv[s] + (size_t)tile * kTileWords, kTileBytes, &sm.full[stage]);
						}
						if (dbytes) {
							bulk_g2s(&sm.dbuf[stage][0], dsrc, dbytes, &sm.full[stage]);
						}
					}
					stage++;
					if (stage == kStages) {
						stage = 0;
						phase ^= 1;
					}
				}
			}
			__syncwarp();
			if (!valid) {
				return;
			}
			tile = next;
		}
	}

	if (warp == kConsumerWarps + 1) {
		// -------------------------------------------------------------- prefix warp
		if (!need_pos) {
			return;
		}
		int64_t t0 = -1;            // last segment handled by this CTA
		unsigned long long s0 = 0; // popcount of segments [0, t0]
		// Request n is served once request n + kPrefixLag has been posted (the end marker counts), i.e.
		// kPrefixLag segment-times after it was posted and still a segment-time before its answer is
		// needed: by then its predecessors' aggregates have (almost always) landed, so the look-back is
		// ONE round of status loads.  Polling at once instead re-reads the holes for microseconds from
		// every CTA, and that L2 traffic costs the bulk-copy stream 10-15 % (profiles/r1_sweep_v13_prefix_lag.log).
		uint32_t seen = 0, exit_at = kNoTile;
		for (uint32_t n = 0;; n++) {
			while (seen <= n + (uint32_t)kPrefixLag && exit_at == kNoTile) {
				mbar_wait_sleep(&sm.req_full[seen % kReqSlots], (seen / kReqSlots) & 1, 200);
				if (sm.req_tile[seen % kReqSlots] == kNoTile) {
					exit_at = seen;
				}
				seen++;
			}
			if (n == exit_at) {
				return;
			}
			const uint32_t slot = n % kReqSlots;
			const uint32_t tile = sm.req_tile[slot];
			const uint32_t total = sm.req_total[slot];
#if CUBIT_PREFIX_DELAY_NS
			__nanosleep(CUBIT_PREFIX_DELAY_NS);
#endif
			const unsigned long long excl =
			    s0 + ((a.debug & 8u) ? 0ull : sum_aggregates(status, t0 + 1, (int64_t)tile, lane));
			if (lane == 0) {
				sm.resp_excl[slot] = excl;
				mbar_arrive(&sm.resp_full[slot]);
				if (a.tile_excl) {
					a.tile_excl[tile] = excl; // consumed by the bit-driven probe kernel
				}
			}
			t0 = tile;
			s0 = excl + total;
			__syncwarp();
		}
	}

	// ---------------------------------------------------------------- consumer warps
	// timing experiments only (see kernels.h): 2 = skip emission, 7 = no look-back traffic at all; anything
	// else could unbalance the request/response mbarriers, so it is ignored
	const unsigned dbg = (a.debug == 2u || a.debug == 7u || a.debug == 10u) ? a.debug : 0u;
	uint32_t stage = 0, phase = 0;
	unsigned long long blk_count = 0; // meaningful in thread 0
	Agg agg;
	uint32_t it = 0;

	// The PENDING segments: merged and counted up to kDefer iterations ago, aggregates published,
	// look-back requests handed to the prefix warp.  The oldest one is emitted WHILE the current
	// segment is merged, one emission step after every batch of ring stages (so the ring keeps
	// draining during a dense write-out) and the rest after the current aggregate is published;
	// by then the prefix warp has had its exclusive prefix ready for a whole segment-time.
	uint64_t pq[kDefer][WPT];
	uint32_t ptile[kDefer] = {}, ptotal[kDefer] = {}, pwexcl[kDefer] = {}, pit[kDefer] = {};
	bool have_p[kDefer] = {};
	const uint32_t n_batches = (a.k + (uint32_t)UB - 1) / (uint32_t)UB;
	bool draining = false; // input exhausted: only pending segments are left
#if CUBIT_SMEM_ADDR
	const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
	const uint32_t stage0 = smem_u32(&sm.stage[0][0]) + (uint32_t)(warp * kSpanWords + lane) * 8u;
#endif

	while (true) {
		uint64_t q[WPT], g[WPT];
#pragma unroll
		for (int i = 0; i < WPT; i++) {
			q[i] = ONEG ? 0ull : ~0ull;
			g[i] = 0;
		}
		uint32_t tile = kNoTile;
		uint32_t tile_total = 0, warp_excl = 0;
		bool acc_dirty = false; // (uniform over the consumer warps) the accumulation tile holds bits of the open OR group
		bool e_todo = have_p[kDefer - 1] && need_emit; // the oldest pending segment is still to be emitted
		bool e_open = false;
		EmitState es;
		const uint32_t nb_iter = draining ? 0u : n_batches;
		for (uint32_t b = 0;; b++) {
			if (b < nb_iter) {
				// ---- fold one batch of ring stages
				const uint32_t s = b * (uint32_t)UB;
				const uint32_t nb = (a.k - s) < (uint32_t)UB ? (a.k - s) : (uint32_t)UB;
				// lane u polls ring stage (stage + u): one mbarrier round trip per batch
				if (lane < (int)nb) {
					uint32_t st = stage + lane, ph = phase;
					if (st >= kStages) {
						st -= kStages;
						ph ^= 1;
					}
#if CUBIT_SMEM_ADDR
					mbar_wait_u32(full0 + st * 8u, ph);
#else
					mbar_wait(&sm.full[st], ph);
#endif
				}
				__syncwarp();
				if (s == 0) {
					tile = sm.meta[stage].tile;
				}
				// CMP: bit u of `skip` = stage u of this batch needs no fold from shared memory — an EMPTY container, or a
				// FULL / ARRAY container that was accumulated straight into the registers below
				uint32_t skip = 0;
				if (HAS_DELTA && !CMP && tile != kNoTile) {
					// Verbatim streams: the pending-delta entries of ALL stages of the batch are handled in one pass —
					// lane l looks at entry (l mod LPS) of stage (l / LPS), LPS = 32 / UB lanes per stage — instead of
					// one pass per stage: with a handful of entries per (stream, segment) the per-stage loop machinery,
					// not the XORs, was what the consumers paid for (≈ 25 instructions per stage and warp, +17 % scan time
					// at 1 % pending rows; profiles/r2_delta_scan.md).  Entries are unordered and may repeat a word
					// (device-side ingestion never sorts), so every warp XORs the ones that fall into ITS span of the
					// staged segment with 32-bit shared-memory atomics — no block-wide barrier — and the fold below
					// picks the words up.
					constexpr int LPS = 32 / UB;
					const int u_l = lane / LPS;
					const uint32_t j_l = (uint32_t)(lane % LPS);
					uint32_t st_l = stage + (uint32_t)u_l;
					st_l = st_l >= (uint32_t)kStages ? st_l - (uint32_t)kStages : st_l;
					const uint32_t dcnt_l = u_l < (int)nb ? sm.meta[st_l].dcnt : 0u;
					const uint32_t maxd = __reduce_max_sync(0xffffffffu, dcnt_l);
					bool wrote = false;
					for (uint32_t e0 = 0; e0 < maxd; e0 += (uint32_t)LPS) {
						const uint32_t e = e0 + j_l;
						if (e < dcnt_l) {
							uint4 raw;
							if (e < (uint32_t)kDeltaStage) {
								raw = *reinterpret_cast<const uint4 *>(&sm.dbuf[st_l][e]);
							} else {
								raw = __ldg(reinterpret_cast<const uint4 *>(a.dent[s + (uint32_t)u_l] + sm.meta[st_l].d0 + e));
							}
							const uint32_t rel = raw.x - (uint32_t)(warp * kSpanWords);
							if (rel < (uint32_t)kSpanWords) {
								unsigned int *w32 = reinterpret_cast<unsigned int *>(&sm.stage[st_l][raw.x]);
								if (raw.z) {
									atomicXor(w32, raw.z);
								}
								if (raw.w) {
									atomicXor(w32 + 1, raw.w);
								}
								wrote = true;
							}
						}
					}
					if (wrote) {
						fence_proxy_async_smem(); // generic-proxy writes before the stage is refilled by the async proxy
					}
					__syncwarp();
				} else if ((HAS_DELTA || CMP) && tile != kNoTile) {
					// Pending-delta entries of one (stream, segment) are unordered and may repeat a word (device-side
					// ingestion never sorts), so every warp XORs the entries that fall into ITS span of the staged segment
					// with 32-bit shared-memory atomics (native ATOMS.XOR; the 64-bit form is a CAS loop) — no block-wide
					// barrier — and the fold below picks the words up.
					// Containers without pending deltas never touch shared memory: FULL ORs all-ones into the
					// accumulator; ARRAY is decoded by the warp — lane l looks at list entries l, l+32, ..., a ballot finds
					// the ones inside the warp's span, and each of those (a handful per warp: sparse segments are what
					// ARRAY containers are for) is broadcast and OR-ed into the register of the lane that owns its word.
					bool wrote = false;
#pragma unroll
					for (int u = 0; u < UB; u++) {
						if (u >= (int)nb) {
							continue;
						}
						uint32_t st = stage + (uint32_t)u;
						st = st >= (uint32_t)kStages ? st - (uint32_t)kStages : st;
						const uint32_t dcnt = HAS_DELTA ? sm.meta[st].dcnt : 0u;
						if (CMP) {
							const uint32_t cmeta = sm.meta[st].pad, ctype = cmeta & 3u;
							if (ctype != CT_BITMAP && dcnt == 0) {
								skip |= 1u << u; // accumulated into the registers by the fold below, in stream order
								continue;
							}
							if (ctype != CT_BITMAP) { // a container WITH pending deltas: materialise the span, then XOR
								const uint64_t fill = ctype == CT_FULL ? ~0ull : 0ull;
								uint64_t *span = &sm.stage[st][warp * kSpanWords];
#pragma unroll
								for (int i = 0; i < WPT; i++) {
									span[i * 32 + lane] = fill;
								}
								wrote = true;
								if (ctype == CT_ARRAY) {
									__syncwarp();
									const uint32_t cnt = cmeta >> 2;
									for (uint32_t e = lane; e < cnt; e += 32) {
										const uint32_t p = sm.abuf[st][e];
										const uint32_t rel = (p >> 6) - (uint32_t)(warp * kSpanWords);
										if (rel < (uint32_t)kSpanWords) {
											atomicOr(reinterpret_cast<unsigned int *>(&sm.stage[st][0]) + (p >> 5), 1u << (p & 31u));
										}
									}
								}
								__syncwarp();
							}
						}
						for (uint32_t e = lane; e < dcnt; e += 32) {
							uint4 raw;
							if (e < (uint32_t)kDeltaStage) {
								raw = *reinterpret_cast<const uint4 *>(&sm.dbuf[st][e]);
							} else {
								raw = __ldg(reinterpret_cast<const uint4 *>(a.dent[s + u] + sm.meta[st].d0 + e));
							}
							const uint32_t rel = raw.x - (uint32_t)(warp * kSpanWords);
							if (rel < (uint32_t)kSpanWords) {
								unsigned int *w32 = reinterpret_cast<unsigned int *>(&sm.stage[st][raw.x]);
								if (raw.z) {
									atomicXor(w32, raw.z);
								}
								if (raw.w) {
									atomicXor(w32 + 1, raw.w);
								}
								wrote = true;
							}
						}
					}
					if (wrote) {
						fence_proxy_async_smem(); // generic-proxy writes before the stage is refilled by the async proxy
					}
					__syncwarp();
				}
				if (tile != kNoTile) {
#pragma unroll
					for (int u = 0; u < UB; u++) {
						if (u < (int)nb) {
							uint32_t st = stage + (uint32_t)u;
							st = st >= (uint32_t)kStages ? st - (uint32_t)kStages : st;
#if CUBIT_SMEM_ADDR
							const uint32_t src = stage0 + st * (uint32_t)kTileBytes;
#define CUBIT_LD(i) lds64(src + (uint32_t)(i) * 256u)
#else
							const uint64_t *src = &sm.stage[st][warp * kSpanWords];
#define CUBIT_LD(i) src[(i) * 32 + lane]
#endif
							if (CMP && ((skip >> u) & 1u)) {
								// container without pending deltas: straight into the accumulator (q for a single OR
								// group, g otherwise), nothing read from the ring stage
								const uint32_t cmeta = sm.meta[st].pad, ctype = cmeta & 3u;
								if (ctype == CT_FULL) {
#pragma unroll
									for (int i = 0; i < WPT; i++) {
										(ONEG ? q[i] : g[i]) = ~0ull;
									}
								} else if (ctype == CT_ARRAY && !(a.debug & (32u | 128u))) {
									// one warp per stage ORs the whole list into the accumulation tile
									if (warp == (int)((s + (uint32_t)u) & (uint32_t)(kConsumerWarps - 1))) {
										const uint32_t cnt = cmeta >> 2;
										unsigned int *acc32 = reinterpret_cast<unsigned int *>(&sm.acc[0]);
										for (uint32_t e = lane; e < cnt; e += 32) {
											const uint32_t p = sm.abuf[st][e];
											atomicOr(acc32 + (p >> 5), 1u << (p & 31u));
										}
									}
									acc_dirty = true;
								} else if (ctype == CT_ARRAY && !(a.debug & 32u)) { // (debug 128: the per-warp filter, for A/B runs)
									const uint32_t cnt = cmeta >> 2;
									for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {
										const uint32_t p = e0 + lane < cnt ? (uint32_t)sm.abuf[st][e0 + lane] : 0xffffffffu;
										// word of the span = (p >> 6) - warp * kSpanWords; inside the span iff < kSpanWords
										const uint32_t w = (p >> 6) - (uint32_t)(warp * kSpanWords);
										// The list is sorted, so the entries inside this warp's span are ONE contiguous run: a ballot
										// finds it, then every lane reads the run's entries straight from shared memory (uniform
										// address → broadcast) and compares each word against the four words IT owns.  No shuffles:
										// a *.sync op inside a loop whose trip count the compiler cannot prove uniform costs a
										// convergence barrier + branch resolve each (ncu: 42 % of all stall samples), and the
										// accumulator index stays a compile-time constant (a run-time q[w >> 5] goes to local memory).
										const uint32_t mine = __ballot_sync(0xffffffffu, p != 0xffffffffu && w < (uint32_t)kSpanWords);
										if (mine) {
											const uint32_t first = e0 + (uint32_t)(__ffs(mine) - 1), n_in = (uint32_t)__popc(mine);
											const uint32_t my0 = (uint32_t)(warp * kSpanWords + lane);
											for (uint32_t j = 0; j < n_in; j += 4) {
												uint32_t pj[4];
#pragma unroll
												for (int r = 0; r < 4; r++) { // (the staging row is kArrayMax entries: reading ≤ 3 past the run stays inside it)
													pj[r] = j + r < n_in ? (uint32_t)sm.abuf[st][(first + j + r) & (kArrayMax - 1)] : 0xffffffffu;
												}
#pragma unroll
												for (int r = 0; r < 4; r++) {
													const uint32_t pw = pj[r] >> 6;
													const uint64_t bit = 1ull << (pj[r] & 63u);
#pragma unroll
													for (int i = 0; i < WPT; i++) {
														(ONEG ? q[i] : g[i]) |= pw == my0 + (uint32_t)(i * 32) ? bit : 0ull;
													}
												}
											}
										}
									}
								}
								if (ONEG) {
									continue;
								}
							} else if (ONEG) {
#pragma unroll
								for (int i = 0; i < WPT; i++) {
									q[i] |= CUBIT_LD(i);
								}
								continue;
							} else {
#pragma unroll
								for (int i = 0; i < WPT; i++) {
									g[i] |= CUBIT_LD(i);
								}
							}
#undef CUBIT_LD
							if ((a.group_end >> (s + u)) & 1ull) {
								if (CMP && acc_dirty) { // the group's ARRAY containers: every warp's atomics, then this warp's span
									consumer_bar_sync();
#pragma unroll
									for (int i = 0; i < WPT; i++) {
										uint64_t *w = &sm.acc[warp * kSpanWords + i * 32 + lane];
										g[i] |= *w;
										*w = 0;
									}
									consumer_bar_sync(); // (before the next group's atomics land in a span that is still being read)
									acc_dirty = false;
								}
#pragma unroll
								for (int i = 0; i < WPT; i++) {
									q[i] &= g[i];
									g[i] = 0;
								}
							}
						}
					}
				}
				__syncwarp();
				if (lane < (int)nb) {
					uint32_t st = stage + (uint32_t)lane;
					st = st >= (uint32_t)kStages ? st - (uint32_t)kStages : st;
#if CUBIT_SMEM_ADDR
					mbar_arrive_u32(empty0 + st * 8u);
#else
					mbar_arrive(&sm.empty[st]);
#endif
				}
				stage += nb;
				if (stage >= kStages) {
					stage -= kStages;
					phase ^= 1;
				}
			} else if (b == nb_iter && !draining) {
				if (tile != kNoTile) {
					if (CMP && ONEG && acc_dirty) { // single OR group: the tile's ARRAY containers, folded once
						consumer_bar_sync();
#pragma unroll
						for (int i = 0; i < WPT; i++) {
							uint64_t *w = &sm.acc[warp * kSpanWords + i * 32 + lane];
							q[i] |= *w;
							*w = 0;
						}
						// (the block scan's barrier below separates this from the next tile's atomics)
					}
					// ---- merged bitvector out (optional)
					if (a.q_out) {
						uint64_t *dst = a.q_out + (size_t)tile * kTileWords + warp * kSpanWords;
#pragma unroll
						for (int i = 0; i < WPT; i++) {
							dst[i * 32 + lane] = q[i];
						}
					}
					// ---- decode, step 1: popcount, warp reduce, block scan over the warp totals
					uint32_t cnt = 0;
#pragma unroll
					for (int i = 0; i < WPT; i++) {
						cnt += __popcll(q[i]);
					}
					cnt = __reduce_add_sync(0xffffffffu, cnt);
					const int buf = it & 1;
					if (lane == 0) {
						sm.warp_tot[buf][warp] = cnt;
					}
					consumer_bar_sync();
					{ // block scan over the 8 warp totals: one shared-memory load and two redux per warp
						const uint32_t wt = lane < kConsumerWarps ? sm.warp_tot[buf][lane] : 0u;
						tile_total = __reduce_add_sync(0xffffffffu, wt);
						warp_excl = __reduce_add_sync(0xffffffffu, lane < warp ? wt : 0u);
					}
					if (threadIdx.x == 0) {
						blk_count += tile_total;
						if (need_pos && !(dbg & 4u)) {
							// publish this segment's aggregate NOW and hand the look-back to the prefix warp
							st_relaxed_u64(&status[tile], kFlagAgg | (unsigned long long)tile_total);
							sm.req_tile[it % kReqSlots] = tile;
							sm.req_total[it % kReqSlots] = tile_total;
							mbar_arrive(&sm.req_full[it % kReqSlots]);
						}
					}
				} else if (need_pos && threadIdx.x == 0) { // (debug & 4: the prefix warp still gets its exit request)
					const uint32_t eslot = (dbg & 4u) ? 0u : it % kReqSlots; // no requests were posted: it still waits on slot 0
					sm.req_tile[eslot] = kNoTile; // tell the prefix warp to exit
					mbar_arrive(&sm.req_full[eslot]);
				}
			}
			// ---- one emission step of the oldest pending segment
			if (e_todo) {
				if (!e_open) {
					e_open = true;
					unsigned long long excl = 0;
					if (need_pos && !(dbg & 1u)) {
						const uint32_t slot = pit[kDefer - 1] % kReqSlots, par = (pit[kDefer - 1] / kReqSlots) & 1;
						if (lane == 0) {
							mbar_wait(&sm.resp_full[slot], par);
						}
						__syncwarp();
						excl = sm.resp_excl[slot];
						if (a.span_excl && lane == 0) { // where this warp's share of the segment starts in the output
							a.span_excl[(size_t)ptile[kDefer - 1] * kConsumerWarps + warp] = excl + pwexcl[kDefer - 1];
						}
					}
					if (ptotal[kDefer - 1] > 0 && !(dbg & 2u)) {
						emit_begin<WPT>(pq[kDefer - 1], excl + pwexcl[kDefer - 1], es);
					}
				}
				if (es.step != kEmitDone) {
					const int64_t span_row0 =
					    a.row_base + ((int64_t)ptile[kDefer - 1] * kTileWords + (int64_t)warp * kSpanWords) * 64;
					if (need_pos) {
						emit_step<WPT, NL, true>(a, pq[kDefer - 1], sm.compact[warp], span_row0, lane, agg, es);
					} else {
						emit_step<WPT, NL, false>(a, pq[kDefer - 1], sm.compact[warp], span_row0, lane, agg, es);
					}
				}
				e_todo = es.step != kEmitDone;
			}
			if (b >= nb_iter && !e_todo) {
				break;
			}
		}
		const bool finished = draining || tile == kNoTile;

		// shift the queue
		bool any_pending = false;
#pragma unroll
		for (int d = kDefer - 1; d > 0; d--) {
			have_p[d] = have_p[d - 1];
			any_pending |= have_p[d];
			ptile[d] = ptile[d - 1];
			ptotal[d] = ptotal[d - 1];
			pwexcl[d] = pwexcl[d - 1];
			pit[d] = pit[d - 1];
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				pq[d][i] = pq[d - 1][i];
			}
		}
		have_p[0] = !finished;
		if (!finished) {
			ptile[0] = tile;
			ptotal[0] = tile_total;
			pwexcl[0] = warp_excl;
			pit[0] = it;
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				pq[0][i] = q[i];
			}
			it++;
		} else {
			draining = true;
			if (!any_pending) {
				break;
			}
		}
	}

	// ---- one exact atomic accumulate per warp / CTA (hdr is zeroed before the launch; no serial
	// last-block pass over per-CTA partials)
	if (NL > 0 && a.agg_kind != 0) {
		agg_flush_warp(agg, a.hdr, lane);
	}
	if (threadIdx.x == 0 && blk_count && !a.skip_count) {
		atomicAdd(&a.hdr->count, blk_count);
	}
}

// ------------------------------------------------------------ bit-driven probe
// The dense-selection probe (SURVEY §8a A3 "direct bit-driven masked streaming when density
// is high"): instead of re-reading 8-byte row IDs it re-decodes the merged bitvector Q
// (1 bit per row) with the per-segment prefixes left by the scan kernel, and gathers the
// columns with the same staged, position-ordered write-out — but as a plain, fully
// occupied grid (48 warps / SM), which is what the gathers need to keep HBM busy.
constexpr int kProbeBitsThreads = 256;

// PK: some probed column is FOR-bit-packed (its pack-block headers are staged per warp); PK = false is the
// all-raw instance, with a smaller staging row and a plain gather.
template <int WPT, int NL, bool POS, bool PK>
__global__ void __launch_bounds__(kProbeBitsThreads, 3) cubit_probe_bits_kernel(const __grid_constant__ ScanArgs a) {
	constexpr int kTileWords = kProbeBitsThreads * WPT;
	constexpr int kSpanWords = WPT * 32;
	__shared__ __align__(16) uint16_t compact[kProbeBitsThreads / 32][kCompactHdrOff + (PK ? NL * kHdrSlots * 8 : 0)];
	__shared__ uint32_t warp_tot[2][kProbeBitsThreads / 32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	Agg agg;
	uint32_t it = 0;
	unsigned long long my_rows = 0; // set bits seen by this thread (a.count_rows)
	// software pipeline: the words of the next TWO segments of this CTA are in flight while one is decoded (one
	// 8 KiB segment in flight per CTA left sparse selections latency-bound: 125 MB in 50 us)
	uint64_t qn[WPT], qn2[WPT];
	unsigned long long excl_n = 0, excl_n2 = 0;
	auto prefetch = [&](uint32_t tl, uint64_t (&dst)[WPT], unsigned long long &ex) {
		if (tl < a.n_seg) {
			const uint64_t *src = a.q_out + (size_t)tl * kTileWords + warp * kSpanWords;
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				dst[i] = __ldg(src + i * 32 + lane);
			}
			if (POS) {
				ex = __ldg(a.tile_excl + tl);
			}
		}
	};
	prefetch(blockIdx.x, qn, excl_n);
	prefetch(blockIdx.x + gridDim.x, qn2, excl_n2);
	for (uint32_t tile = blockIdx.x; tile < a.n_seg; tile += gridDim.x, it++) {
		uint64_t q[WPT];
		uint32_t cnt = 0;
#pragma unroll
		for (int i = 0; i < WPT; i++) {
			q[i] = qn[i];
			qn[i] = qn2[i];
			cnt += __popcll(q[i]);
		}
		my_rows += cnt;
		const unsigned long long tile_excl = excl_n;
		excl_n = excl_n2;
		prefetch(tile + 2 * gridDim.x, qn2, excl_n2);
		unsigned long long wbase = 0;
		if (POS) {
			cnt = __reduce_add_sync(0xffffffffu, cnt);
			if (lane == 0) {
				warp_tot[it & 1][warp] = cnt;
			}
			__syncthreads();
			uint32_t warp_excl = 0;
#pragma unroll
			for (int w = 0; w < kProbeBitsThreads / 32; w++) {
				warp_excl += (w < warp) ? warp_tot[it & 1][w] : 0u;
			}
			wbase = tile_excl + warp_excl;
		}
		const int64_t span_row0 = a.row_base + ((int64_t)tile * kTileWords + (int64_t)warp * kSpanWords) * 64;
		emit_span<WPT, NL, POS, PK>(a, q, compact[warp], wbase, span_row0, lane, agg);
	}
	if (a.agg_kind != 0) {
		agg_flush_warp(agg, a.hdr, lane);
	}
	if (a.count_rows) { // the probe ran directly on a value bitvector (single-bitvector predicate): COUNT comes from here
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			my_rows += __shfl_xor_sync(0xffffffffu, my_rows, d);
		}
		if (lane == 0 && my_rows) {
			atomicAdd(&a.hdr->count, my_rows);
		}
	}
}

template <int WPT, int NL, bool POS, bool PK>
static cudaError_t launch_probe_bits_p(const ScanArgs &args, int sm_count, cudaStream_t stream) {
	auto kern = cubit_probe_bits_kernel<WPT, NL, POS, PK>;
	static int blocks_per_sm = 0; // same for every B200
	if (blocks_per_sm == 0) {
		int b = 0;
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kProbeBitsThreads, 0);
		if (e != cudaSuccess) {
			return e;
		}
		blocks_per_sm = b < 1 ? 1 : b;
	}
	long long grid = (long long)sm_count * blocks_per_sm; // resident CTAs, segments strided, next segment prefetched
	if (grid > (long long)args.n_seg) {
		grid = args.n_seg;
	}
	if (grid < 1) {
		grid = 1;
	}
	kern<<<(unsigned)grid, kProbeBitsThreads, 0, stream>>>(args);
	return cudaGetLastError();
}

template <int WPT, int NL, bool POS>
static cudaError_t launch_probe_bits_t(const ScanArgs &args, int sm_count, cudaStream_t stream) {
	bool packed = false;
	for (int c = 0; c < NL; c++) {
		packed |= args.lcol[c].raw == nullptr;
	}
	return packed ? launch_probe_bits_p<WPT, NL, POS, true>(args, sm_count, stream)
	              : launch_probe_bits_p<WPT, NL, POS, false>(args, sm_count, stream);
}

template <int WPT>
static cudaError_t launch_probe_bits_w(const ScanArgs &args, bool positions, int sm_count, cudaStream_t stream) {
	if (args.n_load == 1) {
		return positions ? launch_probe_bits_t<WPT, 1, true>(args, sm_count, stream)
		                 : launch_probe_bits_t<WPT, 1, false>(args, sm_count, stream);
	}
	if (args.n_load == 2) {
		return positions ? launch_probe_bits_t<WPT, 2, true>(args, sm_count, stream)
		                 : launch_probe_bits_t<WPT, 2, false>(args, sm_count, stream);
	}
	return cudaErrorInvalidValue;
}

cudaError_t launch_probe_bits(const ScanArgs &args, uint32_t seg_words, bool positions, int sm_count,
                              cudaStream_t stream) {
	switch (seg_words) {
	case 512:
		return launch_probe_bits_w<2>(args, positions, sm_count, stream);
	case 1024:
		return launch_probe_bits_w<4>(args, positions, sm_count, stream);
	case 2048:
		return launch_probe_bits_w<8>(args, positions, sm_count, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

// --------------------------------------------------------------------- launch
template <int WPT, bool HAS_DELTA, int NL, bool ONEG, bool CMP>
static cudaError_t launch_scan_g(const ScanArgs &args, int sm_count, cudaStream_t stream, int *grid_out) {
	auto kern = cubit_scan_kernel<WPT, HAS_DELTA, NL, ONEG, CMP>;
	const size_t smem = sizeof(ScanSmem<WPT, NL, CMP>) + 128;
	// function attributes are per device: configure once per (template instance, device)
	static int blocks_per_sm_dev[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (blocks_per_sm_dev[dev] == 0) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) {
			return e;
		}
		int b = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kScanThreads, smem);
		if (e != cudaSuccess) {
			return e;
		}
		blocks_per_sm_dev[dev] = b < 1 ? 1 : (b > 2 ? 2 : b);
	}
	long long grid = (long long)sm_count * blocks_per_sm_dev[dev]; // persistent CTAs, all co-resident
	if (grid > (long long)args.n_seg) {
		grid = args.n_seg;
	}
	if (grid < 1) {
		grid = 1;
	}
	if (grid_out) {
		*grid_out = (int)grid;
	}
	// tickets in flight per CTA: ≈ 16 bulk copies' worth, 2..8 (see the producer warp)
	ScanArgs largs = args;
	const uint32_t td = 16u / (args.k ? args.k : 1u);
	largs.ticket_depth = td < 2u ? 2u : (td > 8u ? 8u : td);
	kern<<<(unsigned)grid, kScanThreads, smem, stream>>>(largs);
	return cudaGetLastError();
}

template <int WPT, bool HAS_DELTA, int NL, bool CMP>
static cudaError_t launch_scan_t(const ScanArgs &args, int sm_count, cudaStream_t stream, int *grid_out) {
	// one OR group ⇔ only the last stream closes a group (the specialised fold exists for the default, unfused path)
	const bool one_group = NL == 0 && args.k >= 1 && args.group_end == (1ull << (args.k - 1));
	if (NL == 0 && one_group) {
		return launch_scan_g<WPT, HAS_DELTA, NL, NL == 0, CMP>(args, sm_count, stream, grid_out);
	}
	return launch_scan_g<WPT, HAS_DELTA, NL, false, CMP>(args, sm_count, stream, grid_out);
}

template <int WPT, bool HAS_DELTA>
static cudaError_t launch_scan_nl(const ScanArgs &args, bool compressed, int sm_count, cudaStream_t stream, int *grid_out) {
	if (compressed) { // container streams: segments of ≤ 65536 rows, probe never fused (cubit_query.cu plans accordingly)
		if (WPT > 4 || args.n_load != 0) {
			return cudaErrorInvalidValue;
		}
		return launch_scan_t<(WPT > 4 ? 4 : WPT), HAS_DELTA, 0, true>(args, sm_count, stream, grid_out);
	}
	switch (args.n_load) {
	case 0:
		return launch_scan_t<WPT, HAS_DELTA, 0, false>(args, sm_count, stream, grid_out);
	case 1:
		return launch_scan_t<WPT, HAS_DELTA, 1, false>(args, sm_count, stream, grid_out);
	case 2:
		return launch_scan_t<WPT, HAS_DELTA, 2, false>(args, sm_count, stream, grid_out);
	default:
		return cudaErrorInvalidValue;
	}
}

cudaError_t launch_scan(const ScanArgs &args, uint32_t seg_words, bool has_delta, bool compressed, int sm_count,
                        cudaStream_t stream, int *grid_out) {
	switch (seg_words) {
	case 512:
		return has_delta ? launch_scan_nl<2, true>(args, compressed, sm_count, stream, grid_out)
		                 : launch_scan_nl<2, false>(args, compressed, sm_count, stream, grid_out);
	case 1024:
		return has_delta ? launch_scan_nl<4, true>(args, compressed, sm_count, stream, grid_out)
		                 : launch_scan_nl<4, false>(args, compressed, sm_count, stream, grid_out);
	case 2048:
		return has_delta ? launch_scan_nl<8, true>(args, compressed, sm_count, stream, grid_out)
		                 : launch_scan_nl<8, false>(args, compressed, sm_count, stream, grid_out);
	default:
		return cudaErrorInvalidValue;
	}
}

int scan_max_grid(uint32_t seg_words, int sm_count) {
	(void)seg_words;
	return sm_count * 2; // two CTAs per SM
}

} // namespace cubit
