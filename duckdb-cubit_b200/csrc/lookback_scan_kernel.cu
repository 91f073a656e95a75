// lookback_scan_kernel.cu — merge + decode of SHORT queries (k ≤ 2 value bitvectors, no pending deltas, verbatim
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
//   3. counts, publishes the tile's total (status word = flag | value), and warp 0 looks back over the status words of
//      the tiles before it — 32 at a time — until it meets one whose inclusive prefix is known; totals are published
//      before any look-back starts, so no tile waits for another tile's LOOK-BACK, only for its count,
//   4. emits row IDs from the registers at the known position (the lane-by-lane store path for units that hold a
//      handful of rows, the staged position-ordered write-out of scan_common.cuh otherwise) and leaves the per-unit /
//      per-tile prefixes the probe kernels want.
// The latency of step 3 (≈ 1–2 µs) is hidden by the other CTAs of the SM (4 resident), not by a ring.
// Reference conventions as in scan_kernel.cu (bit order validity_mask.hpp:163-168, sorted unique row ids
// art.cpp:974-985).
#include "scan_common.cuh"

namespace cubit {

namespace {

constexpr int kLbThreads = 256;
constexpr int kLbWarps = kLbThreads / 32;
constexpr int kLbUnitWpt = 4;                  // 64-bit words per lane and unit (unit = 128 words = 8192 rows)
constexpr int kLbUnitWords = 32 * kLbUnitWpt;
constexpr int kLbUnitsPerWarp = 4;
constexpr int kLbTileUnits = kLbWarps * kLbUnitsPerWarp; // 32 units = 32 KiB per bitvector
constexpr int kLbDirectMax = 96;               // rows per unit up to which row IDs are stored lane by lane
constexpr unsigned long long kLbFlagAgg = 1ull << 62;  // value = the tile's own total
constexpr unsigned long long kLbFlagIncl = 2ull << 62; // value = total of all tiles up to and including this one
constexpr unsigned long long kLbValMask = (1ull << 62) - 1;

} // namespace

template <bool ONEG>
__global__ void __launch_bounds__(kLbThreads, 4) cubit_scan_lookback_kernel(const __grid_constant__ SmallScanArgs s,
                                                                           const __grid_constant__ ScanArgs a) {
	__shared__ __align__(16) uint16_t compact[kLbWarps][kCompactHdrOff];
	__shared__ unsigned long long warp_tot[kLbWarps];
	__shared__ unsigned long long tile_prefix;
	__shared__ uint32_t tile_s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long *status = a.ctrl + 1; // [n_tiles], zeroed; a.ctrl[0] = the ticket counter
	if (threadIdx.x == 0) {
		tile_s = (uint32_t)atomicAdd(a.ctrl, 1ull);
	}
	__syncthreads();
	const uint32_t tile = tile_s;
	const uint32_t u0 = tile * kLbTileUnits + warp * kLbUnitsPerWarp;

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

	// ---- publish the total, look back (warp 0), publish the inclusive prefix
	if (warp == 0) {
		unsigned long long total = 0;
#pragma unroll
		for (int w = 0; w < kLbWarps; w++) {
			total += warp_tot[w];
		}
		if (lane == 0) {
			st_relaxed_u64(&status[tile], (tile == 0 ? kLbFlagIncl : kLbFlagAgg) | total);
		}
		unsigned long long excl = 0;
		if (tile > 0) {
			int64_t hi = (int64_t)tile - 1; // window [hi - 31, hi], lane l reads tile hi - l
			while (true) {
				const int64_t idx = hi - lane;
				unsigned long long v = kLbFlagIncl; // tiles before 0: an inclusive prefix of 0
				if (idx >= 0) {
					do {
						v = ld_relaxed_u64(&status[idx]);
					} while ((v >> 62) == 0ull); // that tile's CTA is running (tickets) and publishes right after counting
				}
				const uint32_t incl_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
				// nearest tile with a known inclusive prefix = lowest such lane; sum the totals of the lanes before it
				const int stop = incl_mask ? __ffs((int)incl_mask) - 1 : 32;
				const unsigned long long part = lane <= stop ? (v & kLbValMask) : 0ull;
				unsigned long long sum = part;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) {
					sum += __shfl_xor_sync(0xffffffffu, sum, d);
				}
				excl += sum;
				if (incl_mask) {
					break;
				}
				hi -= 32;
			}
			if (lane == 0) {
				st_relaxed_u64(&status[tile], kLbFlagIncl | (excl + total));
			}
		}
		if (lane == 0) {
			tile_prefix = excl;
			if ((tile + 1) * (uint32_t)kLbTileUnits >= s.n_units) { // the last tile knows COUNT
				s.hdr->count = excl + total;
			}
		}
	}
	__syncthreads();

	// ---- emit from the registers at the known position
	unsigned long long pos = tile_prefix;
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
// ctrl: [1 + n_tiles] words, zeroed (ticket counter, then one status word per tile)
uint32_t lookback_scan_tiles(uint32_t n_units) {
	return (n_units + kLbTileUnits - 1) / kLbTileUnits;
}

cudaError_t launch_lookback_scan(const SmallScanArgs &s, const ScanArgs &a, cudaStream_t stream) {
	const uint32_t n_tiles = lookback_scan_tiles(s.n_units);
	const bool one_group = s.k >= 1 && s.group_end == (1ull << (s.k - 1));
	if (one_group) {
		cubit_scan_lookback_kernel<true><<<n_tiles, kLbThreads, 0, stream>>>(s, a);
	} else {
		cubit_scan_lookback_kernel<false><<<n_tiles, kLbThreads, 0, stream>>>(s, a);
	}
	return cudaGetLastError();
}

} // namespace cubit
