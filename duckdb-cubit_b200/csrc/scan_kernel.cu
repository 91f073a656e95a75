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
//   * persistent CTAs (SM count × occupancy), segments handed out by a ticket
//     counter, so every predecessor of a segment is already owned by a running
//     CTA — the decoupled look-back below can never wait on unscheduled work.
//   * one producer warp streams whole segments of every queried bitvector with
//     1-D bulk async copies (cp.async.bulk → SASS UBLKCP, the TMA engine) into a
//     kScanStages-deep shared-memory ring guarded by mbarriers; it runs ahead
//     of the consumers by the ring depth, which is what keeps ≥64 KiB per SM in
//     flight while consumers are busy decoding or waiting in the look-back.
//   * 8 consumer warps: apply the (sparse) pending-delta words of the staged
//     segment (XOR), fold the segment into registers (OR within a group, AND
//     across groups), 128-bit ld.shared per lane, conflict free.
//   * decode: __popcll per word, warp __shfl_up scan, block scan over 8 warp
//     totals, single-pass inter-segment prefix by decoupled look-back (one
//     64-bit status word per segment, no atomics on the data path), then every
//     set bit is written by "lane = bit" so each store instruction writes one
//     contiguous run of row IDs (coalesced, sorted, no atomics).
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
__device__ __forceinline__ void st_stream_s64(long long *p, long long v) {
	asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_nc_s64(const long long *p) {
	long long v;
	asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}

// look-back status word: [63:62] flag, [61:0] value
constexpr unsigned long long kFlagAgg = 1ull << 62;    // value = popcount of this segment
constexpr unsigned long long kFlagPrefix = 2ull << 62; // value = popcount of segments [0, this]
constexpr unsigned long long kValMask = (1ull << 62) - 1;

struct StageMeta {
	uint32_t tile; // 0xffffffff = no more work
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

template <int WPT>
struct ScanSmem {
	static constexpr int kTileWords = kConsumerThreads * WPT;
	static constexpr int kTileBytes = kTileWords * 8;
	alignas(128) uint64_t stage[kScanStages][kTileWords];
	alignas(8) uint64_t full[kScanStages];
	uint64_t empty[kScanStages];
	StageMeta meta[kScanStages];
	uint32_t warp_tot[2][kConsumerWarps];
	unsigned long long tile_base[2];
	BlockPartial red[kConsumerWarps];
};

// WPT: 64-bit words of Q each consumer thread holds → segment = 256*WPT words
//      (WPT 2/4/8 ↔ 32768/65536/131072 rows per segment).
template <int WPT, bool HAS_DELTA>
__global__ void __launch_bounds__(kScanThreads) cubit_scan_kernel(const __grid_constant__ ScanArgs a) {
	using Smem = ScanSmem<WPT>;
	constexpr int kTileWords = Smem::kTileWords;
	constexpr int kTileBytes = Smem::kTileBytes;
	constexpr int NCH = WPT / 2; // 16-byte chunks per lane
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	unsigned int *ticket = reinterpret_cast<unsigned int *>(a.ctrl);
	unsigned int *done_ctr = ticket + 1;
	unsigned long long *status = a.ctrl + 1;

	if (threadIdx.x == 0) {
		for (int s = 0; s < kScanStages; s++) {
			mbar_init(&sm.full[s], 1);
			mbar_init(&sm.empty[s], kConsumerWarps);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == kConsumerWarps) {
		// ------------------------------------------------------------ producer warp
		uint32_t stage = 0, phase = 0;
		uint32_t tile = 0;
		if (lane == 0) {
			tile = atomicAdd(ticket, 1u);
		}
		tile = __shfl_sync(0xffffffffu, tile, 0);
		while (true) {
			const bool valid = tile < a.n_seg;
			uint32_t next = 0;
			if (valid && lane == 0) {
				next = atomicAdd(ticket, 1u); // prefetched: consumed one segment later
			}
			if (!valid) {
				if (lane == 0) {
					mbar_wait(&sm.empty[stage], phase ^ 1);
					sm.meta[stage].tile = 0xffffffffu;
					mbar_arrive(&sm.full[stage]);
				}
				break;
			}
			// delta CSR offsets of this segment: lane i serves streams i and i+32
			uint32_t dlo0 = 0, dhi0 = 0, dlo1 = 0, dhi1 = 0;
			if (HAS_DELTA) {
				if (lane < (int)a.k && a.doff[lane]) {
					dlo0 = __ldg(a.doff[lane] + tile);
					dhi0 = __ldg(a.doff[lane] + tile + 1);
				}
				if (lane + 32 < (int)a.k && a.doff[lane + 32]) {
					dlo1 = __ldg(a.doff[lane + 32] + tile);
					dhi1 = __ldg(a.doff[lane + 32] + tile + 1);
				}
			}
			for (uint32_t s = 0; s < a.k; s++) {
				uint32_t d0 = 0, d1 = 0;
				if (HAS_DELTA) {
					d0 = __shfl_sync(0xffffffffu, s < 32 ? dlo0 : dlo1, s & 31);
					d1 = __shfl_sync(0xffffffffu, s < 32 ? dhi0 : dhi1, s & 31);
				}
				if (lane == 0) {
					mbar_wait(&sm.empty[stage], phase ^ 1);
					sm.meta[stage].tile = tile;
					sm.meta[stage].d0 = d0;
					sm.meta[stage].dcnt = d1 - d0;
					mbar_arrive_expect_tx(&sm.full[stage], kTileBytes);
					bulk_g2s(&sm.stage[stage][0], a.bv[s] + (size_t)tile * kTileWords, kTileBytes, &sm.full[stage]);
				}
				stage++;
				if (stage == kScanStages) {
					stage = 0;
					phase ^= 1;
				}
			}
			next = __shfl_sync(0xffffffffu, next, 0);
			tile = next;
		}
		return;
	}

	// ---------------------------------------------------------------- consumer warps
	const unsigned lanemask_lt = (1u << lane) - 1u;
	uint32_t stage = 0, phase = 0;
	unsigned long long blk_count = 0; // meaningful in thread 0
	unsigned long long sum_lo = 0;
	long long sum_hi = 0;
	unsigned int overflow = 0;
	uint32_t it = 0;

	while (true) {
		uint64_t q[WPT], g[WPT];
#pragma unroll
		for (int i = 0; i < WPT; i++) {
			q[i] = ~0ull;
			g[i] = 0;
		}
		uint32_t tile = 0;
		bool finished = false;
		for (uint32_t s = 0; s < a.k; s++) {
			mbar_wait(&sm.full[stage], phase);
			if (s == 0) {
				tile = sm.meta[stage].tile;
				if (tile == 0xffffffffu) {
					finished = true;
					break;
				}
			}
			if (HAS_DELTA) {
				const uint32_t dcnt = sm.meta[stage].dcnt;
				if (dcnt) { // uniform across the consumer warps
					const DeltaEnt *ent = a.dent[s] + sm.meta[stage].d0;
					for (uint32_t e = threadIdx.x; e < dcnt; e += kConsumerThreads) {
						const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(ent + e));
						const uint64_t mask = ((uint64_t)raw.w << 32) | raw.z;
						sm.stage[stage][raw.x] ^= mask; // words are unique per (stream, segment)
					}
					fence_proxy_async_smem();
					consumer_bar_sync();
				}
			}
			const uint4 *src = reinterpret_cast<const uint4 *>(&sm.stage[stage][warp * (WPT * 32)]);
#pragma unroll
			for (int j = 0; j < NCH; j++) {
				const uint4 v = src[j * 32 + lane];
				g[2 * j] |= ((uint64_t)v.y << 32) | v.x;
				g[2 * j + 1] |= ((uint64_t)v.w << 32) | v.z;
			}
			if ((a.group_end >> s) & 1ull) {
#pragma unroll
				for (int i = 0; i < WPT; i++) {
					q[i] &= g[i];
					g[i] = 0;
				}
			}
			__syncwarp();
			if (lane == 0) {
				mbar_arrive(&sm.empty[stage]);
			}
			stage++;
			if (stage == kScanStages) {
				stage = 0;
				phase ^= 1;
			}
		}
		if (finished) {
			break;
		}

		// ---- merged bitvector out (optional)
		if (a.q_out) {
			uint4 *dst = reinterpret_cast<uint4 *>(a.q_out + (size_t)tile * kTileWords + warp * (WPT * 32));
#pragma unroll
			for (int j = 0; j < NCH; j++) {
				uint4 v;
				v.x = (uint32_t)q[2 * j];
				v.y = (uint32_t)(q[2 * j] >> 32);
				v.z = (uint32_t)q[2 * j + 1];
				v.w = (uint32_t)(q[2 * j + 1] >> 32);
				dst[j * 32 + lane] = v;
			}
		}

		// ---- decode: popcount, warp scan, block scan
		uint32_t off[WPT]; // exclusive offset of every word inside this warp's span
		uint32_t warp_total = 0;
#pragma unroll
		for (int j = 0; j < NCH; j++) {
			const uint32_t c0 = __popcll(q[2 * j]);
			const uint32_t c1 = __popcll(q[2 * j + 1]);
			uint32_t incl = c0 + c1;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) {
					incl += n;
				}
			}
			const uint32_t excl = incl - (c0 + c1) + warp_total;
			off[2 * j] = excl;
			off[2 * j + 1] = excl + c0;
			warp_total += __shfl_sync(0xffffffffu, incl, 31);
		}
		const int buf = it & 1;
		if (lane == 0) {
			sm.warp_tot[buf][warp] = warp_total;
		}
		consumer_bar_sync();
		uint32_t tile_total = 0, warp_excl = 0;
#pragma unroll
		for (int w = 0; w < kConsumerWarps; w++) {
			const uint32_t t = sm.warp_tot[buf][w];
			warp_excl += (w < warp) ? t : 0u;
			tile_total += t;
		}
		if (threadIdx.x == 0) {
			blk_count += tile_total;
		}

		const bool need_pos = (a.ids_out != nullptr) || (a.n_vcols > 0);
		if (need_pos) {
			// ---- single-pass inter-segment prefix: decoupled look-back by warp 0
			if (warp == 0) {
				unsigned long long excl = 0;
				if (tile == 0) {
					if (lane == 0) {
						st_relaxed_u64(&status[0], kFlagPrefix | (unsigned long long)tile_total);
					}
				} else {
					if (lane == 0) {
						st_relaxed_u64(&status[tile], kFlagAgg | (unsigned long long)tile_total);
					}
					int64_t look = (int64_t)tile - 1; // window [look-31, look]
					while (true) {
						const int64_t idx = look - lane;
						unsigned long long sv = kFlagPrefix; // out-of-range slots act as a zero prefix
						if (idx >= 0) {
							do {
								sv = ld_relaxed_u64(&status[idx]);
							} while ((sv >> 62) == 0);
						}
						const unsigned pmask = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
						// lanes nearer than the first inclusive prefix contribute their aggregates
						const int first = pmask ? (__ffs(pmask) - 1) : 32;
						unsigned long long contrib = (lane <= first) ? (sv & kValMask) : 0ull;
#pragma unroll
						for (int d = 16; d > 0; d >>= 1) {
							contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
						}
						excl += contrib;
						if (pmask) {
							break;
						}
						look -= 32;
					}
					if (lane == 0) {
						st_relaxed_u64(&status[tile], kFlagPrefix | (excl + tile_total));
					}
				}
				if (lane == 0) {
					sm.tile_base[buf] = excl;
				}
			}
			consumer_bar_sync();
			const unsigned long long wbase = sm.tile_base[buf] + warp_excl;
			const int64_t span_row0 =
			    a.row_base + ((int64_t)tile * kTileWords + (int64_t)warp * (WPT * 32)) * 64; // global id of span bit 0
			const int64_t local_adj = -a.row_base; // global id → local row for column probes

			// ---- emit: lane = bit.  For each non-zero word the warp writes its set
			// bits as one contiguous run of row IDs (two 32-bit halves).
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				const int j = i >> 1, h = i & 1;
				unsigned nz = __ballot_sync(0xffffffffu, q[i] != 0);
				while (nz) {
					const int src = __ffs(nz) - 1;
					nz &= nz - 1;
					const uint32_t wlo = __shfl_sync(0xffffffffu, (uint32_t)q[i], src);
					const uint32_t whi = __shfl_sync(0xffffffffu, (uint32_t)(q[i] >> 32), src);
					const uint32_t o = __shfl_sync(0xffffffffu, off[i], src);
					const int64_t row0 = span_row0 + (int64_t)(j * 64 + src * 2 + h) * 64;
					const unsigned long long p0 = wbase + o;
#pragma unroll
					for (int half = 0; half < 2; half++) {
						const uint32_t wv = half ? whi : wlo;
						if ((wv >> lane) & 1u) {
							const unsigned long long p =
							    p0 + (half ? __popc(wlo) : 0) + __popc(wv & lanemask_lt);
							const int64_t rid = row0 + half * 32 + lane;
							if (p < a.ids_cap) {
								if (a.ids_out) {
									st_stream_s64(a.ids_out + p, rid);
								}
								for (int c = 0; c < a.n_vcols; c++) {
									st_stream_s64(a.vout[c] + p, ld_nc_s64(a.vcol[c] + (rid + local_adj)));
								}
							}
							if (a.agg_kind == 1) {
								add128(sum_lo, sum_hi, ld_nc_s64(a.agg_a + (rid + local_adj)));
							} else if (a.agg_kind == 2) {
								const long long x = ld_nc_s64(a.agg_a + (rid + local_adj));
								const long long y = ld_nc_s64(a.agg_b + (rid + local_adj));
								const long long pr = x * y;
								if (__mul64hi(x, y) != (pr >> 63)) {
									overflow = 1;
								}
								add128(sum_lo, sum_hi, pr);
							}
						}
					}
				}
			}
		} else if (a.agg_kind != 0) {
			// aggregate only: bit-driven masked probe, no row IDs materialised
			const int64_t span_local0 = ((int64_t)tile * kTileWords + (int64_t)warp * (WPT * 32)) * 64;
#pragma unroll
			for (int i = 0; i < WPT; i++) {
				const int j = i >> 1, h = i & 1;
				unsigned nz = __ballot_sync(0xffffffffu, q[i] != 0);
				while (nz) {
					const int src = __ffs(nz) - 1;
					nz &= nz - 1;
					const uint32_t wlo = __shfl_sync(0xffffffffu, (uint32_t)q[i], src);
					const uint32_t whi = __shfl_sync(0xffffffffu, (uint32_t)(q[i] >> 32), src);
					const int64_t row0 = span_local0 + (int64_t)(j * 64 + src * 2 + h) * 64;
#pragma unroll
					for (int half = 0; half < 2; half++) {
						const uint32_t wv = half ? whi : wlo;
						if ((wv >> lane) & 1u) {
							const int64_t r = row0 + half * 32 + lane;
							if (a.agg_kind == 1) {
								add128(sum_lo, sum_hi, ld_nc_s64(a.agg_a + r));
							} else {
								const long long x = ld_nc_s64(a.agg_a + r);
								const long long y = ld_nc_s64(a.agg_b + r);
								const long long pr = x * y;
								if (__mul64hi(x, y) != (pr >> 63)) {
									overflow = 1;
								}
								add128(sum_lo, sum_hi, pr);
							}
						}
					}
				}
			}
		}
		it++;
	}

	// ---- block reduction of count / 128-bit sum, then last-block finalisation
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const unsigned long long olo = __shfl_xor_sync(0xffffffffu, sum_lo, d);
		const long long ohi = __shfl_xor_sync(0xffffffffu, sum_hi, d);
		add128(sum_lo, sum_hi, olo, ohi);
		overflow |= __shfl_xor_sync(0xffffffffu, overflow, d);
	}
	if (lane == 0) {
		sm.red[warp].sum_lo = sum_lo;
		sm.red[warp].sum_hi = sum_hi;
		sm.red[warp].pad = overflow;
	}
	consumer_bar_sync();
	if (threadIdx.x == 0) {
		unsigned long long lo = 0;
		long long hi = 0;
		unsigned long long ovf = 0;
		for (int w = 0; w < kConsumerWarps; w++) {
			add128(lo, hi, sm.red[w].sum_lo, sm.red[w].sum_hi);
			ovf |= sm.red[w].pad;
		}
		BlockPartial bp;
		bp.count = blk_count;
		bp.sum_lo = lo;
		bp.sum_hi = hi;
		bp.pad = ovf;
		volatile BlockPartial *dst = a.partials + blockIdx.x;
		dst->count = bp.count;
		dst->sum_lo = bp.sum_lo;
		dst->sum_hi = bp.sum_hi;
		dst->pad = bp.pad;
		__threadfence();
		const unsigned int prev = atomicAdd(done_ctr, 1u);
		if (prev == gridDim.x - 1) {
			__threadfence();
			unsigned long long cnt = 0, tlo = 0, tovf = 0;
			long long thi = 0;
			for (unsigned b = 0; b < gridDim.x; b++) {
				const volatile BlockPartial *p = a.partials + b;
				cnt += p->count;
				add128(tlo, thi, p->sum_lo, p->sum_hi);
				tovf |= p->pad;
			}
			a.hdr->count = cnt;
			a.hdr->sum_lo = tlo;
			a.hdr->sum_hi = thi;
			a.hdr->overflow = (unsigned int)tovf;
		}
	}
}

// --------------------------------------------------------------------- launch
template <int WPT, bool HAS_DELTA>
static cudaError_t launch_scan_t(const ScanArgs &args, int sm_count, cudaStream_t stream, int *grid_out) {
	auto kern = cubit_scan_kernel<WPT, HAS_DELTA>;
	const size_t smem = sizeof(ScanSmem<WPT>) + 128;
	static bool configured = false; // per template instance
	static int blocks_per_sm = 1;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) {
			return e;
		}
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kScanThreads, smem);
		if (e != cudaSuccess) {
			return e;
		}
		if (blocks_per_sm < 1) {
			blocks_per_sm = 1;
		}
		configured = true;
	}
	long long grid = (long long)sm_count * blocks_per_sm;
	if (grid > (long long)args.n_seg) {
		grid = args.n_seg;
	}
	if (grid < 1) {
		grid = 1;
	}
	if (grid_out) {
		*grid_out = (int)grid;
	}
	kern<<<(unsigned)grid, kScanThreads, smem, stream>>>(args);
	return cudaGetLastError();
}

cudaError_t launch_scan(const ScanArgs &args, uint32_t seg_words, bool has_delta, int sm_count, cudaStream_t stream,
                        int *grid_out) {
	switch (seg_words) {
	case 512:
		return has_delta ? launch_scan_t<2, true>(args, sm_count, stream, grid_out)
		                 : launch_scan_t<2, false>(args, sm_count, stream, grid_out);
	case 1024:
		return has_delta ? launch_scan_t<4, true>(args, sm_count, stream, grid_out)
		                 : launch_scan_t<4, false>(args, sm_count, stream, grid_out);
	case 2048:
		return has_delta ? launch_scan_t<8, true>(args, sm_count, stream, grid_out)
		                 : launch_scan_t<8, false>(args, sm_count, stream, grid_out);
	default:
		return cudaErrorInvalidValue;
	}
}

int scan_max_grid(uint32_t seg_words, int sm_count) {
	// upper bound used to size the per-block partial array: smallest stage
	// footprint (WPT=2) allows the most CTAs per SM; 16 is a safe ceiling.
	(void)seg_words;
	return sm_count * 16;
}

} // namespace cubit
