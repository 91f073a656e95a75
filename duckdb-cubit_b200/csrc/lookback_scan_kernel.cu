// lookback_scan_kernel.cu — merge + decode of SHORT queries (k ≤ 3 value bitvectors, no pending deltas, verbatim
// bitvectors) that want row positions, in ONE pass with a decoupled look-back (sm_100a).
//
// What it computes is what scan_kernel.cu computes:
//     Q = AND_j ( OR_{i in R_j} B_i )          ids = ascending positions of the set bits of Q (+ row_base)
// Why a third implementation (profiles/r2_small_k.md): the ring kernel spends ≈ 1.9 µs of dependent latency per tile
// and CTA whatever the tile holds (1.05 TB/s on one bitvector); the two streaming passes of small_scan_kernels.cu
// read the bitvector twice and need three launches (≈ 2 TB/s).  Here every CTA
//   1. draws a tile of 32 units (32 KiB of every bitvector, 262,144 rows) from a ticket counter — tickets are handed
//      out in tile order to CTAs that are running, so a tile never waits for one that has not started,
//   2. loads its tile ONCE with plain coalesced 64-bit loads (16 per lane and bitvector, all in flight) and keeps the
//      merged words in registers,
//   3. counts, publishes the tile's total (status word = flag | value) and computes its output position chain-free:
//      inclusive prefix of tile − 1024 (finished long ago: only ≈ 600 tickets are in flight) + the totals of the tiles
//      in between, summed by the whole CTA in one L2 round trip; no tile waits for another tile's LOOK-BACK, only for
//      its count,
//   4. emits row IDs from the registers at the known position (the lane-by-lane store path for units that hold a
//      handful of rows, the staged position-ordered write-out of scan_common.cuh otherwise) and leaves the per-unit /
//      per-tile prefixes the probe kernels want.
// The latency of step 3 (≈ 1–2 µs) is hidden by the other CTAs of the SM (4 resident), not by a ring.
// Reference conventions as in scan_kernel.cu (bit order validity_mask.hpp:163-168, sorted unique row ids
// art.cpp:974-985).
#include "scan_common.cuh"

#include <cstdlib>

namespace cubit {

namespace {

constexpr int kLbThreads = 256;
constexpr int kLbWarps = kLbThreads / 32;
constexpr int kLbUnitWpt = 4;                  // 64-bit words per lane and unit (unit = 128 words = 8192 rows)
constexpr int kLbUnitWords = 32 * kLbUnitWpt;
constexpr int kLbUnitsPerWarp = 4;
constexpr int kLbTileUnits = kLbWarps * kLbUnitsPerWarp; // 32 units = 32 KiB per bitvector
constexpr int kLbAnchor = 1024;                // look-back: inclusive prefix of tile − 1024 + the totals in between
constexpr int kLbDirectMax = 96;               // rows per unit up to which row IDs are stored lane by lane
constexpr unsigned long long kLbFlag = 1ull << 62;     // status word published
constexpr unsigned long long kLbValMask = (1ull << 62) - 1;

} // namespace

template <bool ONEG>
__global__ void __launch_bounds__(kLbThreads, 4) cubit_scan_lookback_kernel(const __grid_constant__ SmallScanArgs s,
                                                                           const __grid_constant__ ScanArgs a,
                                                                           const uint32_t pf_dist) {
	__shared__ __align__(16) uint16_t compact[kLbWarps][kCompactHdrOff];
	__shared__ unsigned long long warp_tot[kLbWarps];
	__shared__ unsigned long long red[kLbWarps];
	__shared__ uint32_t tile_s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long *agg_w = a.ctrl + 1;  // [n_tiles], zeroed: flag | the tile's own total; a.ctrl[0] = the ticket counter
	unsigned long long *incl_w = s.chunk_tot; // [n_tiles], zeroed: flag | total of all tiles up to and including this one
	if (threadIdx.x == 0) {
		tile_s = (uint32_t)atomicAdd(a.ctrl, 1ull);
	}
	__syncthreads();
	const uint32_t tile = tile_s;
	const uint32_t u0 = tile * kLbTileUnits + warp * kLbUnitsPerWarp;
	// Experiment (CUBIT_LB_PREFETCH = tickets ahead, default off): ask L2 for the tile a later CTA will draw, one
	// 128-byte line per thread and bitvector.  Measured: no gain (k = 1: 43 → 47 µs, k = 2: 89 → 108 µs) — the kernel is
	// bound by the per-CTA chain (ticket → load → count → look-back → emit ≈ 6.7 µs × 6.4 waves), not by DRAM latency.
	if (pf_dist) {
		const uint64_t w = ((uint64_t)tile + pf_dist) * (kLbTileUnits * kLbUnitWords) + threadIdx.x * 16u;
		if (w < (uint64_t)s.n_units * kLbUnitWords) {
			for (uint32_t st = 0; st < s.k; st++) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(s.bv[st] + w));
			}
		}
	}

	// ---- load + merge: this warp's four units, every word in flight at once
	uint64_t q[kLbUnitsPerWarp][kLbUnitWpt];
	if (s.k == 1) {
#pragma unroll
		for (int h = 0; h < kLbUnitsPerWarp; h++) {
			const bool in = u0 + h < s.n_units;
			const uint64_t *src = s.bv[0] + (size_t)(u0 + h) * kLbUnitWords + lane;
#pragma unroll
			for (int i = 0; i < kLbUnitWpt; i++) {
				q[h][i] = in ? __ldg(src + i * 32) : 0ull;
			}
		}
	} else if (ONEG) {
#pragma unroll
		for (int h = 0; h < kLbUnitsPerWarp; h++) {
#pragma unroll
			for (int i = 0; i < kLbUnitWpt; i++) {
				q[h][i] = 0ull;
			}
		}
		for (uint32_t st = 0; st < s.k; st++) {
#pragma unroll
			for (int h = 0; h < kLbUnitsPerWarp; h++) {
				const bool in = u0 + h < s.n_units;
				const uint64_t *src = s.bv[st] + (size_t)(u0 + h) * kLbUnitWords + lane;
#pragma unroll
				for (int i = 0; i < kLbUnitWpt; i++) {
					q[h][i] |= in ? __ldg(src + i * 32) : 0ull;
				}
			}
		}
	} else {
		// AND of OR groups: one unit at a time keeps the group accumulator at four registers
#pragma unroll
		for (int h = 0; h < kLbUnitsPerWarp; h++) {
			const bool in = u0 + h < s.n_units;
			uint64_t g[kLbUnitWpt];
#pragma unroll
			for (int i = 0; i < kLbUnitWpt; i++) {
				q[h][i] = in ? ~0ull : 0ull; // (units past the end of the table select nothing)
				g[i] = 0;
			}
			for (uint32_t st = 0; st < s.k; st++) {
				const uint64_t *src = s.bv[st] + (size_t)(u0 + h) * kLbUnitWords + lane;
#pragma unroll
				for (int i = 0; i < kLbUnitWpt; i++) {
					g[i] |= in ? __ldg(src + i * 32) : 0ull;
				}
				if ((s.group_end >> st) & 1ull) {
#pragma unroll
					for (int i = 0; i < kLbUnitWpt; i++) {
						q[h][i] &= g[i];
						g[i] = 0;
					}
				}
			}
		}
	}
	if (s.q_out) { // the merged bitvector, for the probe kernels / the caller (k = 1 without it: Q is bv[0] itself)
#pragma unroll
		for (int h = 0; h < kLbUnitsPerWarp; h++) {
			if (u0 + h < s.n_units) {
#pragma unroll
				for (int i = 0; i < kLbUnitWpt; i++) {
					s.q_out[(size_t)(u0 + h) * kLbUnitWords + lane + i * 32] = q[h][i];
				}
			}
		}
	}

	// ---- count: per unit (uniform in the warp), per warp, per tile
	uint32_t ucnt[kLbUnitsPerWarp];
	uint32_t wsum = 0;
#pragma unroll
	for (int h = 0; h < kLbUnitsPerWarp; h++) {
		uint32_t c = 0;
#pragma unroll
		for (int i = 0; i < kLbUnitWpt; i++) {
			c += __popcll(q[h][i]);
		}
		ucnt[h] = __reduce_add_sync(0xffffffffu, c);
		wsum += ucnt[h];
	}
	if (lane == 0) {
		warp_tot[warp] = wsum;
	}
	__syncthreads();

	// ---- publish the total; look back WITHOUT a chain: the tiles in flight are the last ≈ 600 tickets, so tile
	// (tile − 1024) finished before this CTA started and its inclusive prefix is there; the totals of the ≤ 1023 tiles in
	// between are summed by the whole CTA (≤ 4 independent loads per thread): one L2 round trip, and the only thing a
	// tile ever waits for is a predecessor's COUNT.  (The classic "nearest inclusive prefix" walk measured 68 µs per
	// 125 MB: the nearest one is a whole wave back, and the polling hammers a handful of L2 lines.)
	unsigned long long total = 0;
#pragma unroll
	for (int w = 0; w < kLbWarps; w++) {
		total += warp_tot[w];
	}
	if (threadIdx.x == 0) {
		st_relaxed_u64(&agg_w[tile], kLbFlag | total);
	}
	const int64_t anchor = (int64_t)tile - kLbAnchor;
	unsigned long long part = 0;
	if (threadIdx.x == 0 && anchor >= 0) {
		unsigned long long v;
		do {
			v = ld_relaxed_u64(&incl_w[anchor]);
		} while (!(v & kLbFlag));
		part = v & kLbValMask;
	}
	for (int64_t idx = (anchor >= 0 ? anchor + 1 : 0) + threadIdx.x; idx < (int64_t)tile; idx += kLbThreads) {
		unsigned long long v;
		do {
			v = ld_relaxed_u64(&agg_w[idx]);
		} while (!(v & kLbFlag)); // that tile's CTA is running (tickets) and publishes right after counting
		part += v & kLbValMask;
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		part += __shfl_xor_sync(0xffffffffu, part, d);
	}
	if (lane == 0) {
		red[warp] = part;
	}
	__syncthreads();
	unsigned long long excl = 0;
#pragma unroll
	for (int w = 0; w < kLbWarps; w++) {
		excl += red[w];
	}
	if (threadIdx.x == 0) {
		st_relaxed_u64(&incl_w[tile], kLbFlag | (excl + total));
		if ((tile + 1) * (uint32_t)kLbTileUnits >= s.n_units) { // the last tile knows COUNT
			s.hdr->count = excl + total;
		}
	}

	// ---- emit from the registers at the known position
	unsigned long long pos = excl;
	for (int w = 0; w < warp; w++) {
		pos += warp_tot[w];
	}
	Agg agg;
#pragma unroll
	for (int h = 0; h < kLbUnitsPerWarp; h++) {
		const uint32_t u = u0 + h;
		if (u >= s.n_units) {
			break;
		}
		const uint32_t cnt = ucnt[h];
		if (lane == 0) {
			if (s.span_excl) {
				s.span_excl[u] = pos;
			}
			if (s.tile_excl && (u & 7u) == 0u) {
				s.tile_excl[u >> 3] = pos;
			}
		}
		if (cnt && a.ids_out) {
			const int64_t row0 = a.row_base + (int64_t)u * (kLbUnitWords * 64);
			if (cnt <= (uint32_t)kLbDirectMax) {
				// a handful of rows in 8192: four interleaved warp scans give every lane the rank of its words and the
				// lanes store their row IDs straight to global memory
				uint32_t c[kLbUnitWpt], incl[kLbUnitWpt];
#pragma unroll
				for (int i = 0; i < kLbUnitWpt; i++) {
					c[i] = (uint32_t)__popcll(q[h][i]);
					incl[i] = c[i];
				}
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
					for (int i = 0; i < kLbUnitWpt; i++) {
						const uint32_t n = __shfl_up_sync(0xffffffffu, incl[i], d);
						if (lane >= d) {
							incl[i] += n;
						}
					}
				}
				unsigned long long base = pos;
#pragma unroll
				for (int i = 0; i < kLbUnitWpt; i++) {
					const uint32_t slot_total = __shfl_sync(0xffffffffu, incl[i], 31);
					unsigned long long at = base + incl[i] - c[i];
					uint64_t w = q[h][i];
					const int64_t wrow = row0 + (int64_t)(i * 32 + lane) * 64;
					while (w) {
						a.ids_out[at++] = wrow + (__ffsll((long long)w) - 1);
						w &= w - 1;
					}
					base += slot_total;
				}
			} else {
				emit_span<kLbUnitWpt, 0, true, false>(a, q[h], compact[warp], pos, row0, lane, agg);
			}
		}
		pos += cnt;
	}
}

// ------------------------------------------------------------------------------------------------ launch
// a.ctrl: [1 + n_tiles] zeroed words (ticket counter, then every tile's total); s.chunk_tot: [n_tiles] zeroed words
// (every tile's inclusive prefix)
uint32_t lookback_scan_tiles(uint32_t n_units) {
	return (n_units + kLbTileUnits - 1) / kLbTileUnits;
}

cudaError_t launch_lookback_scan(const SmallScanArgs &s, const ScanArgs &a, int sm_count, cudaStream_t stream) {
	const uint32_t n_tiles = lookback_scan_tiles(s.n_units);
	(void)sm_count;
	uint32_t pf_dist = 0;
	if (const char *e = getenv("CUBIT_LB_PREFETCH")) { // experiment knob (profiles/r2_small_k.md)
		pf_dist = (uint32_t)strtoul(e, nullptr, 10);
	}
	const bool one_group = s.k >= 1 && s.group_end == (1ull << (s.k - 1));
	if (one_group) {
		cubit_scan_lookback_kernel<true><<<n_tiles, kLbThreads, 0, stream>>>(s, a, pf_dist);
	} else {
		cubit_scan_lookback_kernel<false><<<n_tiles, kLbThreads, 0, stream>>>(s, a, pf_dist);
	}
	return cudaGetLastError();
}

} // namespace cubit
