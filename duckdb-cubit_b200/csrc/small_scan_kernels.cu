// small_scan_kernels.cu — the merge + decode of SHORT queries (k ≤ 4 value bitvectors, no pending deltas, verbatim
// bitvectors) as two streaming passes instead of the single-pass ring kernel (sm_100a).
//
// What it computes is what scan_kernel.cu computes:
//     Q = AND_j ( OR_{i in R_j} B_i )          ids = ascending positions of the set bits of Q (+ row_base)
// Why a second implementation: the ring kernel orders its output with a per-tile count → publish → look-back →
// emit chain that costs a CTA ≈ 1–2 µs of dependent latency per tile whatever the tile holds.  With ten
// bitvectors per tile that hides behind the bulk copies (0.91 of the HBM roofline); with one it IS the run time:
// 125 MB in 52 µs count-only, 127 µs with row IDs (profiles/r2_small_k.md).  For few bitvectors reading the
// (one) merged bitvector a second time is cheaper than that chain:
//   pass A  cubit_merge_count_kernel   every warp owns a contiguous CHUNK of 8192-row units; per unit it folds the k
//           words per lane straight from global memory (plain coalesced 64-bit loads, 4·k in flight per lane, no
//           ring, no barrier), writes Q when k > 1 (for k = 1 Q is the bitvector itself), and adds the unit's
//           popcount to its chunk total.  Count-only / bitvector-only / aggregate-only queries stop here.
//   pass B  cubit_chunk_prefix_kernel  one CTA: exclusive prefix over the ≤ 9,472 chunk totals (+ COUNT).
//   pass C  cubit_decode_kernel        every warp walks ITS chunk again — backwards, so that what pass A touched last
//           is what pass C touches first and the 126 MB L2 serves part of the re-read — re-reads only Q (one
//           bitvector whatever k was), recomputes the unit counts and emits the row IDs at known positions with the
//           staged, position-ordered write-out of scan_common.cuh.  No tickets, no look-back, no inter-CTA
//           dependency at all.  It also leaves the per-span / per-tile prefixes the probe kernels want.
// Reference conventions as in scan_kernel.cu (bit order validity_mask.hpp:163-168, sorted unique row ids
// art.cpp:974-985).
#include "scan_common.cuh"

namespace cubit {

constexpr int kUnitWpt = 4;                    // 64-bit words per lane and unit
constexpr int kUnitWords = 32 * kUnitWpt;      // 128 words = 8192 rows = one scan-kernel span of a 65536-row tile
constexpr int kSmallThreads = 256;
constexpr int kDirectMax = 96;                // rows per unit up to which pass C stores row IDs lane by lane

// ---------------------------------------------------------------------------------------------- pass A
template <bool ONEG>
__global__ void __launch_bounds__(kSmallThreads) cubit_merge_count_kernel(const __grid_constant__ SmallScanArgs s) {
	const int lane = threadIdx.x & 31;
	const uint32_t chunk = (blockIdx.x * kSmallThreads + threadIdx.x) >> 5;
	if (chunk >= s.n_chunks) {
		return;
	}
	const uint32_t u0 = chunk * s.units_per_chunk;
	const uint32_t u1 = min(u0 + s.units_per_chunk, s.n_units);
	unsigned long long total = 0;
	// two units per iteration: 8·k independent loads in flight per lane (one unit at a time left short queries
	// latency-bound: 125 MB in 30 µs)
	for (uint32_t u = u0; u < u1; u += 2) {
		const bool two = u + 1 < u1;
		uint64_t q[2][kUnitWpt], g[2][kUnitWpt];
#pragma unroll
		for (int h = 0; h < 2; h++) {
#pragma unroll
			for (int i = 0; i < kUnitWpt; i++) {
				q[h][i] = ONEG ? 0ull : ~0ull;
				g[h][i] = 0;
			}
		}
#pragma unroll 4
		for (uint32_t st = 0; st < s.k; st++) {
			const uint64_t *src = s.bv[st] + (size_t)u * kUnitWords + lane;
#pragma unroll
			for (int h = 0; h < 2; h++) {
				if (h == 0 || two) {
#pragma unroll
					for (int i = 0; i < kUnitWpt; i++) {
						(ONEG ? q[h][i] : g[h][i]) |= __ldg(src + h * kUnitWords + i * 32);
					}
				}
			}
			if (!ONEG && ((s.group_end >> st) & 1ull)) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
#pragma unroll
					for (int i = 0; i < kUnitWpt; i++) {
						q[h][i] &= g[h][i];
						g[h][i] = 0;
					}
				}
			}
		}
		uint32_t cnt = 0;
#pragma unroll
		for (int h = 0; h < 2; h++) {
			if (h == 0 || two) {
#pragma unroll
				for (int i = 0; i < kUnitWpt; i++) {
					cnt += __popcll(q[h][i]);
					if (s.q_out) {
						s.q_out[(size_t)(u + h) * kUnitWords + lane + i * 32] = q[h][i];
					}
				}
			}
		}
		total += cnt;
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		total += __shfl_xor_sync(0xffffffffu, total, d);
	}
	if (lane == 0) {
		s.chunk_tot[chunk] = total;
		if (s.count_here && total) { // no pass B follows (the query wants no row positions): COUNT straight from here
			atomicAdd(&s.hdr->count, total);
		}
	}
}

// ---------------------------------------------------------------------------------------------- pass B
// chunk_tot[0 .. n) → exclusive prefixes in place, chunk_tot[n] = total; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) cubit_chunk_prefix_kernel(unsigned long long *chunk_tot, uint32_t n, ResultHeader *hdr) {
	__shared__ unsigned long long warp_sum[32];
	__shared__ unsigned long long carry_s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		carry_s = 0;
	}
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const unsigned long long v = i < n ? chunk_tot[i] : 0ull;
		unsigned long long incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) {
				incl += o;
			}
		}
		if (lane == 31) {
			warp_sum[warp] = incl;
		}
		__syncthreads();
		unsigned long long wex = 0;
		for (int w = 0; w < warp; w++) {
			wex += warp_sum[w];
		}
		const unsigned long long carry = carry_s;
		if (i < n) {
			chunk_tot[i] = carry + wex + incl - v;
		}
		__syncthreads();
		if (threadIdx.x == 1023) {
			carry_s = carry + wex + incl;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		chunk_tot[n] = carry_s;
		if (hdr) {
			hdr->count += carry_s;
		}
	}
}

// ---------------------------------------------------------------------------------------------- pass C
__global__ void __launch_bounds__(kSmallThreads) cubit_decode_kernel(const __grid_constant__ SmallScanArgs s,
                                                                    const __grid_constant__ ScanArgs a) {
	__shared__ __align__(16) uint16_t compact[kSmallThreads / 32][kCompactHdrOff];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// (chunks in reverse launch order: the first CTAs re-read what pass A finished with)
	const uint32_t wid = (blockIdx.x * kSmallThreads + threadIdx.x) >> 5;
	if (wid >= s.n_chunks) {
		return;
	}
	const uint32_t chunk = s.n_chunks - 1u - wid;
	const uint32_t u0 = chunk * s.units_per_chunk;
	const uint32_t u1 = min(u0 + s.units_per_chunk, s.n_units);
	unsigned long long pos_end = __ldg(s.chunk_tot + chunk + 1); // exclusive prefix of the NEXT chunk = end of this one
	Agg agg;
	uint64_t qn[kUnitWpt], qn2[kUnitWpt]; // the next two units (walking backwards) are in flight while one is written out
	auto load = [&](uint32_t u, uint64_t (&dst)[kUnitWpt]) {
		const uint64_t *src = s.q_in + (size_t)u * kUnitWords + lane;
#pragma unroll
		for (int i = 0; i < kUnitWpt; i++) {
			dst[i] = __ldg(src + i * 32);
		}
	};
	if (u1 > u0) {
		load(u1 - 1, qn);
	}
	if (u1 > u0 + 1) {
		load(u1 - 2, qn2);
	}
	for (uint32_t u = u1; u-- > u0;) {
		uint64_t q[kUnitWpt];
		uint32_t cnt = 0;
#pragma unroll
		for (int i = 0; i < kUnitWpt; i++) {
			q[i] = qn[i];
			qn[i] = qn2[i];
			cnt += __popcll(q[i]);
		}
		if (u > u0 + 1) {
			load(u - 2, qn2);
		}
		cnt = __reduce_add_sync(0xffffffffu, cnt);
		pos_end -= cnt;
		if (lane == 0) {
			if (s.span_excl) {
				s.span_excl[u] = pos_end;
			}
			if (s.tile_excl && (u & 7u) == 0u) {
				s.tile_excl[u >> 3] = pos_end;
			}
		}
		if (cnt && a.ids_out) {
			const int64_t row0 = a.row_base + (int64_t)u * (kUnitWords * 64);
			if (cnt <= (uint32_t)kDirectMax) {
				// a handful of rows in 8192: four interleaved warp scans give every lane the rank of its words, and the
				// lanes store their row IDs straight to global memory (≈ 90 instructions per unit instead of the ≈ 400
				// of the staged, coalesced write-out, which pays off only when there is something to coalesce)
				uint32_t c[kUnitWpt], incl[kUnitWpt];
#pragma unroll
				for (int i = 0; i < kUnitWpt; i++) {
					c[i] = (uint32_t)__popcll(q[i]);
					incl[i] = c[i];
				}
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
					for (int i = 0; i < kUnitWpt; i++) {
						const uint32_t n = __shfl_up_sync(0xffffffffu, incl[i], d);
						if (lane >= d) {
							incl[i] += n;
						}
					}
				}
				unsigned long long base = pos_end;
#pragma unroll
				for (int i = 0; i < kUnitWpt; i++) {
					const uint32_t slot_total = __shfl_sync(0xffffffffu, incl[i], 31);
					unsigned long long at = base + incl[i] - c[i];
					uint64_t w = q[i];
					const int64_t wrow = row0 + (int64_t)(i * 32 + lane) * 64;
					while (w) {
						a.ids_out[at++] = wrow + (__ffsll((long long)w) - 1);
						w &= w - 1;
					}
					base += slot_total;
				}
			} else {
				emit_span<kUnitWpt, 0, true, false>(a, q, compact[warp], pos_end, row0, lane, agg);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ launch
cudaError_t launch_small_merge_count(const SmallScanArgs &s, cudaStream_t stream) {
	const unsigned grid = (s.n_chunks * 32u + kSmallThreads - 1) / kSmallThreads;
	const bool one_group = s.k >= 1 && s.group_end == (1ull << (s.k - 1));
	if (one_group) {
		cubit_merge_count_kernel<true><<<grid, kSmallThreads, 0, stream>>>(s);
	} else {
		cubit_merge_count_kernel<false><<<grid, kSmallThreads, 0, stream>>>(s);
	}
	return cudaGetLastError();
}

cudaError_t launch_small_prefix(const SmallScanArgs &s, cudaStream_t stream) {
	cubit_chunk_prefix_kernel<<<1, 1024, 0, stream>>>(s.chunk_tot, s.n_chunks, s.count_here ? nullptr : s.hdr);
	return cudaGetLastError();
}

cudaError_t launch_small_decode(const SmallScanArgs &s, const ScanArgs &a, cudaStream_t stream) {
	const unsigned grid = (s.n_chunks * 32u + kSmallThreads - 1) / kSmallThreads;
	cubit_decode_kernel<<<grid, kSmallThreads, 0, stream>>>(s, a);
	return cudaGetLastError();
}

// chunks = the warps one resident wave of the decode pass holds (a second, partial wave doubled its time), at least one unit each
void small_scan_plan(SmallScanArgs &s, uint32_t n_units, int sm_count) {
	const uint32_t warps = (uint32_t)sm_count * 48u; // what the decode pass keeps resident (6 CTAs of 8 warps per SM): one wave
	s.n_units = n_units;
	s.units_per_chunk = (n_units + warps - 1) / warps;
	if (s.units_per_chunk < 1) {
		s.units_per_chunk = 1;
	}
	s.n_chunks = (n_units + s.units_per_chunk - 1) / s.units_per_chunk;
}

} // namespace cubit
