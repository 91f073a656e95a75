// scan_kernel_ws.cu — warp-specialised variant of the fused segment-merge + decode kernel.
//
// Same algorithm, bit order, look-back and emission as scan_kernel.cu (see there for the
// description and the reference citations); what changes is WHO does what inside the CTA, so
// that merging and emitting overlap instead of alternating:
//
//   warp 0        producer   : ticket + cp.async.bulk (TMA) ring of bitvector segments
//   warp 1        prefix     : chain-free decoupled look-back (sum of published aggregates)
//   warps 2..5    FOLD       : wait ring stages, XOR pending deltas, OR/AND into registers,
//                              write the merged segment Q into a shared-memory Q ring together
//                              with its per-span popcounts, publish the aggregate — never emit
//   warps 6..21   EMIT       : take (segment, 64-word span) items off the Q ring, compact the
//                              set bits and write row IDs / gather / sum — never wait for HBM
//
// One persistent CTA per SM (704 threads, ~200 KiB of shared memory).  Fold warps run at the
// speed of the TMA ring (they do what the count-only kernel does), emit warps at the speed of
// the stores; the Q ring (4 segments) decouples the two, so the kernel's time approaches
// max(merge, emit, HBM) instead of merge + emit.
#include "scan_common.cuh"

namespace cubit {

constexpr int kWsFoldWarps = 4;
constexpr int kWsFoldThreads = kWsFoldWarps * 32;
constexpr int kWsEmitWarps = 16;
constexpr int kWsThreads = (2 + kWsFoldWarps + kWsEmitWarps) * 32; // 704
constexpr int kWsRingBytes = 96 * 1024;
constexpr int kWsQRingBytes = 32 * 1024;
constexpr int kSpanWordsWs = 64; // words per emit item = 2 words per lane = 4096 rows

template <int SEG> // segment size in 64-bit words: 512 / 1024 / 2048
struct WsSmem {
	static constexpr int kTileBytes = SEG * 8;
	static constexpr int kStages = kWsRingBytes / kTileBytes;
	static constexpr int kQSlots = kWsQRingBytes / kTileBytes;
	static constexpr int kSpans = SEG / kSpanWordsWs; // emit items per segment
	alignas(128) uint64_t stage[kStages][SEG];
	alignas(128) uint64_t qslot[kQSlots][SEG];
	alignas(16) uint16_t compact[kWsEmitWarps][kSlotRows + 8 + 32];
	alignas(16) DeltaEnt dbuf[kStages][kDeltaStage];
	alignas(8) uint64_t full[kStages];
	uint64_t empty[kStages];
	uint64_t qfull[kQSlots];
	uint64_t qempty[kQSlots];
	uint64_t req_full[kQSlots];
	uint64_t resp_full[kQSlots];
	unsigned long long resp_excl[kQSlots];
	uint32_t req_tile[kQSlots];
	uint32_t req_total[kQSlots];
	uint32_t qtile[kQSlots];
	uint32_t qcnt[kQSlots][kSpans];
	StageMeta meta[kStages];
	uint32_t poff[2 * kMaxStreams];
};

__device__ __forceinline__ void fold_bar_sync() {
	asm volatile("bar.sync 2, %0;" ::"n"(kWsFoldThreads) : "memory");
}

template <int SEG, bool HAS_DELTA, int NL>
__global__ void __launch_bounds__(kWsThreads, 1) cubit_scan_ws_kernel(const __grid_constant__ ScanArgs a) {
	using Smem = WsSmem<SEG>;
	constexpr int kStages = Smem::kStages;
	constexpr int kQSlots = Smem::kQSlots;
	constexpr int kSpans = Smem::kSpans;
	constexpr int kTileBytes = Smem::kTileBytes;
	constexpr int WF = SEG / kWsFoldThreads; // words per fold thread (4 / 8 / 16)
	constexpr int NCHF = WF / 2;             // 16-byte chunks per fold thread
	constexpr int UB = SEG >= 2048 ? 2 : (kWaitBatch < kStages ? kWaitBatch : kStages);
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
			mbar_init(&sm.empty[s], kWsFoldWarps);
		}
		for (int s = 0; s < kQSlots; s++) {
			mbar_init(&sm.qfull[s], 1);
			mbar_init(&sm.qempty[s], kSpans);
			mbar_init(&sm.req_full[s], 1);
			mbar_init(&sm.resp_full[s], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == 0) {
		// ------------------------------------------------------------ producer warp
		uint32_t stage = 0, phase = 0;
		uint32_t tile = 0;
		if (lane == 0) {
			tile = atomicAdd(ticket, 1u);
		}
		tile = __shfl_sync(0xffffffffu, tile, 0);
		uint32_t dlo[2] = {0, 0}, dhi[2] = {0, 0};
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
		if (HAS_DELTA) {
			load_offsets(tile);
		}
		while (true) {
			const bool valid = tile < a.n_seg;
			uint32_t next = 0;
			if (valid && lane == 0) {
				next = atomicAdd(ticket, 1u);
			}
			next = __shfl_sync(0xffffffffu, next, 0);
			if (HAS_DELTA) {
#pragma unroll
				for (int h = 0; h < 2; h++) {
					sm.poff[lane + 32 * h] = dlo[h];
					sm.poff[64 + lane + 32 * h] = dhi[h];
				}
				__syncwarp();
				if (valid) {
					load_offsets(next);
				}
			}
			if (lane == 0) {
				for (uint32_t s = 0; s < a.k; s++) {
					mbar_wait(&sm.empty[stage], phase ^ 1);
					sm.meta[stage].tile = valid ? tile : kNoTile;
					if (!valid) {
						mbar_arrive(&sm.full[stage]);
					} else {
						uint32_t dbytes = 0;
						const DeltaEnt *dsrc = nullptr;
						if (HAS_DELTA) {
							const uint32_t d0 = sm.poff[s], d1 = sm.poff[64 + s];
							sm.meta[stage].d0 = d0;
							sm.meta[stage].dcnt = d1 - d0;
							if (d1 > d0) {
								dbytes = ((d1 - d0) < (uint32_t)kDeltaStage ? (d1 - d0) : (uint32_t)kDeltaStage) *
								         (uint32_t)sizeof(DeltaEnt);
								dsrc = a.dent[s] + d0;
							}
						}
						mbar_arrive_expect_tx(&sm.full[stage], kTileBytes + dbytes);
						bulk_g2s(&sm.stage[stage][0], a.bv[s] + (size_t)tile * SEG, kTileBytes, &sm.full[stage]);
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

	if (warp == 1) {
		// -------------------------------------------------------------- prefix warp
		if (!need_pos) {
			return;
		}
		int64_t t0 = -1;
		unsigned long long s0 = 0;
		for (uint32_t n = 0;; n++) {
			const uint32_t slot = n % kQSlots, par = (n / kQSlots) & 1;
			mbar_wait(&sm.req_full[slot], par);
			const uint32_t tile = sm.req_tile[slot];
			if (tile == kNoTile) {
				return;
			}
			const uint32_t total = sm.req_total[slot];
			const unsigned long long excl = s0 + sum_aggregates(status, t0 + 1, (int64_t)tile, lane);
			if (lane == 0) {
				sm.resp_excl[slot] = excl;
				mbar_arrive(&sm.resp_full[slot]);
				if (a.tile_excl) {
					a.tile_excl[tile] = excl;
				}
			}
			t0 = tile;
			s0 = excl + total;
			__syncwarp();
		}
	}

	if (warp < 2 + kWsFoldWarps) {
		// ---------------------------------------------------------------- fold warps
		const int fw = warp - 2;                 // fold warp index 0..3
		const int ft = fw * 32 + lane;           // fold thread index 0..127
		uint32_t stage = 0, phase = 0;
		unsigned long long my_count = 0; // count-only mode: per-lane popcount; else thread 0: CTA total
		for (uint32_t n = 0;; n++) {
			uint64_t q[WF], g[WF];
#pragma unroll
			for (int i = 0; i < WF; i++) {
				q[i] = ~0ull;
				g[i] = 0;
			}
			uint32_t tile = 0;
			for (uint32_t s = 0; s < a.k; s += UB) {
				const uint32_t nb = (a.k - s) < (uint32_t)UB ? (a.k - s) : (uint32_t)UB;
				if (lane < (int)nb) {
					uint32_t st = stage + lane, ph = phase;
					if (st >= kStages) {
						st -= kStages;
						ph ^= 1;
					}
					mbar_wait(&sm.full[st], ph);
				}
				__syncwarp();
				if (s == 0) {
					tile = sm.meta[stage].tile;
				}
				if (HAS_DELTA && tile != kNoTile) {
					// this warp owns the words whose 64-word block index is ≡ fw (mod 4)
					bool wrote = false;
					for (uint32_t u = 0; u < nb; u++) {
						const uint32_t st = (stage + u) % kStages;
						const uint32_t dcnt = sm.meta[st].dcnt;
						for (uint32_t e = lane; e < dcnt; e += 32) {
							uint4 raw;
							if (e < (uint32_t)kDeltaStage) {
								raw = *reinterpret_cast<const uint4 *>(&sm.dbuf[st][e]);
							} else {
								raw = __ldg(reinterpret_cast<const uint4 *>(a.dent[s + u] + sm.meta[st].d0 + e));
							}
							if (((raw.x >> 6) & 3u) == (uint32_t)fw) {
								sm.stage[st][raw.x] ^= ((uint64_t)raw.w << 32) | raw.z;
								wrote = true;
							}
						}
					}
					if (wrote) {
						fence_proxy_async_smem();
					}
					__syncwarp();
				}
				if (tile != kNoTile) {
#pragma unroll
					for (int u = 0; u < UB; u++) {
						if (u < (int)nb) {
							const uint64_t *src = &sm.stage[(stage + u) % kStages][0];
#pragma unroll
							for (int j = 0; j < NCHF; j++) {
								const uint4 v = *reinterpret_cast<const uint4 *>(src + j * 256 + ft * 2);
								g[2 * j] |= ((uint64_t)v.y << 32) | v.x;
								g[2 * j + 1] |= ((uint64_t)v.w << 32) | v.z;
							}
							if ((a.group_end >> (s + u)) & 1ull) {
#pragma unroll
								for (int i = 0; i < WF; i++) {
									q[i] &= g[i];
									g[i] = 0;
								}
							}
						}
					}
				}
				__syncwarp();
				if (lane < (int)nb) {
					mbar_arrive(&sm.empty[(stage + lane) % kStages]);
				}
				stage += nb;
				if (stage >= kStages) {
					stage -= kStages;
					phase ^= 1;
				}
			}
			const bool finished = tile == kNoTile;
			const uint32_t slot = n % kQSlots, par = (n / kQSlots) & 1;

			if (!finished) {
				if (a.q_out) {
					uint64_t *dst = a.q_out + (size_t)tile * SEG;
#pragma unroll
					for (int j = 0; j < NCHF; j++) {
						uint4 v;
						v.x = (uint32_t)q[2 * j];
						v.y = (uint32_t)(q[2 * j] >> 32);
						v.z = (uint32_t)q[2 * j + 1];
						v.w = (uint32_t)(q[2 * j + 1] >> 32);
						*reinterpret_cast<uint4 *>(dst + j * 256 + ft * 2) = v;
					}
				}
				if (!need_emit) {
					// count only: no Q ring, no per-segment synchronisation at all
#pragma unroll
					for (int i = 0; i < WF; i++) {
						my_count += __popcll(q[i]);
					}
					continue;
				}
				// hand the merged segment to the emit warps through the Q ring
				if (lane == 0) {
					mbar_wait(&sm.qempty[slot], par ^ 1);
				}
				__syncwarp();
#pragma unroll
				for (int j = 0; j < NCHF; j++) {
					uint4 v;
					v.x = (uint32_t)q[2 * j];
					v.y = (uint32_t)(q[2 * j] >> 32);
					v.z = (uint32_t)q[2 * j + 1];
					v.w = (uint32_t)(q[2 * j + 1] >> 32);
					*reinterpret_cast<uint4 *>(&sm.qslot[slot][j * 256 + ft * 2]) = v;
					// chunk j of fold warp fw is exactly emit span j*4 + fw
					const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)(__popcll(q[2 * j]) + __popcll(q[2 * j + 1])));
					if (lane == 0) {
						sm.qcnt[slot][j * 4 + fw] = c;
					}
				}
			}
			fold_bar_sync();
			if (ft == 0) {
				if (!finished) {
					uint32_t total = 0;
#pragma unroll
					for (int sp = 0; sp < kSpans; sp++) {
						total += sm.qcnt[slot][sp];
					}
					my_count += total;
					if (need_pos) {
						st_relaxed_u64(&status[tile], kFlagAgg | (unsigned long long)total);
						sm.req_tile[slot] = tile;
						sm.req_total[slot] = total;
						mbar_arrive(&sm.req_full[slot]);
					}
					sm.qtile[slot] = tile;
					mbar_arrive(&sm.qfull[slot]);
				} else if (need_emit) {
					if (need_pos) {
						sm.req_tile[slot] = kNoTile; // prefix warp: exit
						mbar_arrive(&sm.req_full[slot]);
					}
					// end markers in every Q slot: whichever item an emit warp asks for next ends it
					for (uint32_t j = 0; j < (uint32_t)kQSlots; j++) {
						const uint32_t n2 = n + j, sl = n2 % kQSlots, pr = (n2 / kQSlots) & 1;
						mbar_wait(&sm.qempty[sl], pr ^ 1);
						sm.qtile[sl] = kNoTile;
						mbar_arrive(&sm.qfull[sl]);
					}
				}
			}
			if (finished) {
				break;
			}
		}
		// count: one atomic per CTA (count-only mode: per warp)
		if (!need_emit) {
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				my_count += __shfl_xor_sync(0xffffffffu, my_count, d);
			}
			if (lane == 0 && my_count && !a.skip_count) {
				atomicAdd(&a.hdr->count, my_count);
			}
		} else if (ft == 0 && my_count && !a.skip_count) {
			atomicAdd(&a.hdr->count, my_count);
		}
		return;
	}

	// ------------------------------------------------------------------ emit warps
	if (!need_emit) {
		return;
	}
	const int ew = warp - 2 - kWsFoldWarps; // 0..15
	unsigned long long sum_lo = 0;
	long long sum_hi = 0;
	unsigned int overflow = 0;
	for (uint32_t item = ew;; item += kWsEmitWarps) {
		const uint32_t n = item / kSpans, sp = item % kSpans;
		const uint32_t slot = n % kQSlots, par = (n / kQSlots) & 1;
		if (lane == 0) {
			mbar_wait(&sm.qfull[slot], par);
		}
		__syncwarp();
		const uint32_t tile = sm.qtile[slot];
		if (tile == kNoTile) {
			break;
		}
		unsigned long long wbase = 0;
		if (need_pos) {
			if (lane == 0) {
				mbar_wait(&sm.resp_full[slot], par);
			}
			__syncwarp();
			const uint32_t mine = (lane < (int)sp && lane < kSpans) ? sm.qcnt[slot][lane] : 0u; // spans before mine
			wbase = sm.resp_excl[slot] + __reduce_add_sync(0xffffffffu, mine);
		}
		uint64_t q[2];
		q[0] = sm.qslot[slot][sp * kSpanWordsWs + lane];
		q[1] = sm.qslot[slot][sp * kSpanWordsWs + 32 + lane];
		const int64_t span_row0 = a.row_base + ((int64_t)tile * SEG + (int64_t)sp * kSpanWordsWs) * 64;
		if (need_pos) {
			emit_span<2, NL, true>(a, q, sm.compact[ew], wbase, span_row0, lane, sum_lo, sum_hi, overflow);
		} else {
			emit_span<2, NL, false>(a, q, sm.compact[ew], 0, span_row0, lane, sum_lo, sum_hi, overflow);
		}
		__syncwarp();
		if (lane == 0) {
			mbar_arrive(&sm.qempty[slot]);
		}
	}
	if (NL > 0 && a.agg_kind != 0) {
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			const unsigned long long olo = __shfl_xor_sync(0xffffffffu, sum_lo, d);
			const long long ohi = __shfl_xor_sync(0xffffffffu, sum_hi, d);
			add128(sum_lo, sum_hi, olo, ohi);
			overflow |= __shfl_xor_sync(0xffffffffu, overflow, d);
		}
		if (lane == 0) {
			if (sum_lo | (unsigned long long)sum_hi) {
				const unsigned long long old = atomicAdd(&a.hdr->sum_lo, sum_lo);
				const long long carry = (old + sum_lo) < old ? 1 : 0;
				atomicAdd(reinterpret_cast<unsigned long long *>(&a.hdr->sum_hi), (unsigned long long)(sum_hi + carry));
			}
			if (overflow) {
				atomicOr(&a.hdr->overflow, 1u);
			}
		}
	}
}

// --------------------------------------------------------------------- launch
template <int SEG, bool HAS_DELTA, int NL>
static cudaError_t launch_ws_t(const ScanArgs &args, int sm_count, cudaStream_t stream) {
	auto kern = cubit_scan_ws_kernel<SEG, HAS_DELTA, NL>;
	const size_t smem = sizeof(WsSmem<SEG>) + 128;
	static bool configured[64] = {}; // function attributes are per device
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (!configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) {
			return e;
		}
		configured[dev] = true;
	}
	long long grid = sm_count; // one persistent CTA per SM
	if (grid > (long long)args.n_seg) {
		grid = args.n_seg;
	}
	if (grid < 1) {
		grid = 1;
	}
	kern<<<(unsigned)grid, kWsThreads, smem, stream>>>(args);
	return cudaGetLastError();
}

template <int SEG, bool HAS_DELTA>
static cudaError_t launch_ws_nl(const ScanArgs &args, int sm_count, cudaStream_t stream) {
	switch (args.n_load) {
	case 0:
		return launch_ws_t<SEG, HAS_DELTA, 0>(args, sm_count, stream);
	case 1:
		return launch_ws_t<SEG, HAS_DELTA, 1>(args, sm_count, stream);
	case 2:
		return launch_ws_t<SEG, HAS_DELTA, 2>(args, sm_count, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

cudaError_t launch_scan_ws(const ScanArgs &args, uint32_t seg_words, bool has_delta, int sm_count, cudaStream_t stream) {
	switch (seg_words) {
	case 512:
		return has_delta ? launch_ws_nl<512, true>(args, sm_count, stream) : launch_ws_nl<512, false>(args, sm_count, stream);
	case 1024:
		return has_delta ? launch_ws_nl<1024, true>(args, sm_count, stream)
		                 : launch_ws_nl<1024, false>(args, sm_count, stream);
	case 2048:
		return has_delta ? launch_ws_nl<2048, true>(args, sm_count, stream)
		                 : launch_ws_nl<2048, false>(args, sm_count, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace cubit
