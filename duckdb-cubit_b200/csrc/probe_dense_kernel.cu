// probe_dense_kernel.cu — the column probe for DENSE selections over FOR-bit-packed columns (sm_100a).
//
// What it computes: vals[c] = col_c[rows selected by Q], in row order, at the output
// positions the scan kernel gave the row IDs, plus SUM / SUM(a*b) over them.  Semantics as everywhere else:
//   fetch of a row's value   src/storage/compression/bitpacking.cpp:879 (BitpackingFetchRow: FOR base + packed delta)
//   SUM carry                src/include/duckdb/core_functions/aggregate/sum_helpers.hpp:92-113
//
// Why a second probe kernel.  The bit-driven probe of scan_kernel.cu gathers every selected value straight from
// global memory: two dependent loads, a header look-up and 64-bit address arithmetic per value — 87 thread
// instructions per value, long-scoreboard bound at 24 warps / SM (profiles/r2_probe_dense.md: 3.6 ms for 5·10^8
// values of a 24-bit column, DRAM pipe 24 % busy).  From a few percent of the rows upward, however, every DRAM
// line of a packed column is touched anyway, so this kernel STREAMS the column instead:
//   * work unit = one pack block (1024 rows: 128·width bytes of payload, one 16-byte header) of one warp's span;
//     every warp owns a double-buffered shared-memory stage per probed column and lane 0 keeps ONE 1-D bulk async
//     copy (cp.async.bulk → UBLKCP, the TMA engine) in flight for the warp's next non-empty block while the
//     current one is decoded — completion on a per-warp mbarrier, no CTA-wide barrier anywhere in the kernel
//     (the exclusive output prefix of every span comes from the scan kernel, ScanArgs::span_excl).  Blocks
//     without a selected row are never copied.
//   * decode from shared memory in OUTPUT-POSITION order: the set bits of the block are compacted to 10-bit row
//     numbers in a per-warp staging row (two ctz chains per lane, the hand-written step of scan_common.cuh), then
//     lane l handles positions 2l, 2l+1 (+64 …): index multiply, two ld.shared.b32 of neighbouring words (position
//     order keeps the lanes' addresses within a few banks of each other — lane order would hit 8-way conflicts at
//     width 24), one funnel shift, a mask, a 64-bit add of the block's base.  Values leave as contiguous
//     512-byte st.global.cs.v2 stores, exactly like the row IDs did.
// Raw (unpacked) columns may ride along (plain gather); widths above 32 bits are left to the gather probe.
#include "scan_common.cuh"

namespace cubit {

constexpr int kDenseWarps = 8;
constexpr int kDenseThreads = kDenseWarps * 32;
// staging row of 10-bit row numbers: one block (+ pad slot) and a whole write-out iteration of slack, so that the
// write-out reads its 128 slots unconditionally
constexpr int kDenseRowsBytes = (kPackBlock + 128 + 8) * 2;
constexpr int kDenseMaxStages = 4;
constexpr int kDenseWarpFixed = kDenseRowsBytes + 8 * kDenseMaxStages; // + the warp's mbarriers (one per stage)

// how the aggregate is accumulated (kernel template parameter, so the write-out loop carries no dispatch):
//   0  no aggregate
//   1  SUM over a bit-packed column: per block the lanes add up the 32-bit FOR deltas and count their rows; the
//      block's base is multiplied in once per block (two instructions per value instead of a 128-bit add)
//   2  everything else (SUM over a raw column, SUM(a*b) with the overflow check, SUM over DOUBLE): per value
enum { DENSE_AGG_NONE = 0, DENSE_AGG_DELTA = 1, DENSE_AGG_GENERAL = 2 };

template <int NL>
struct BlockHdrs { // headers of the block being decoded, one per probed column (warp-uniform registers)
	long long base[NL];
	uint32_t width[NL];
	uint32_t mask[NL];
	uint32_t pk[NL]; // shared-window address of the staged payload
};

// FOR delta of row r of a staged block: bits [r*width, (r+1)*width) of the payload
__device__ __forceinline__ uint32_t dense_delta(uint32_t pk, uint32_t r, uint32_t width, uint32_t mask) {
	const uint32_t bit = r * width;
	const uint32_t ad = pk + ((bit >> 5) << 2);
	uint32_t lo, hi;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lo) : "r"(ad));
	asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(hi) : "r"(ad));
	return __funnelshift_r(lo, hi, bit) & mask; // (the shift amount is taken modulo 32)
}

struct DenseAcc { // per-lane accumulators of the block being decoded (DENSE_AGG_DELTA)
	unsigned long long dsum = 0;
	uint32_t rows = 0;
};

// One write-out iteration: staged rows [2*g0, 2*g0 + 128) of the block → 128 consecutive output positions; lane l
// handles pairs g0 + l and g0 + 32 + l.  FULL: every one of the 128 slots holds a selected row (no predicates).
template <int NL, bool POS, bool RAW, int AGGM, bool FULL>
__device__ __forceinline__ void dense_iter(const DenseProbeArgs &a, const uint32_t *cb32, uint32_t g0, uint32_t pad,
                                           uint32_t end, unsigned long long obase, long long local0,
                                           const BlockHdrs<NL> &bh, int lane, Agg &agg, DenseAcc &acc) {
	long long v[2][2][NL];
	uint32_t dl[2][2]; // FOR deltas of the aggregate column (DENSE_AGG_DELTA)
	bool ok[2][2];
	constexpr int CA = 0; // DENSE_AGG_DELTA: the aggregate column is column agg_ia; resolved below
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const uint32_t g = g0 + h * 32 + lane; // pair index
		const uint32_t packed = cb32[g];
		// (10-bit row numbers: the mask keeps the decode of a slot past the list — stale staging bytes — inside the stage)
		const uint32_t r[2] = {FULL ? packed & 0xffffu : packed & (uint32_t)(kPackBlock - 1),
		                       FULL ? packed >> 16 : (packed >> 16) & (uint32_t)(kPackBlock - 1)};
		ok[h][0] = FULL || (g * 2 >= pad && g * 2 < end);
		ok[h][1] = FULL || g * 2 + 1 < end;
#pragma unroll
		for (int e = 0; e < 2; e++) {
#pragma unroll
			for (int c = 0; c < NL; c++) {
				if (RAW && a.lcol[c].raw) {
					v[h][e][c] = ok[h][e] ? __ldg(a.lcol[c].raw + local0 + r[e]) : 0;
				} else {
					const uint32_t d = dense_delta(bh.pk[c], r[e], bh.width[c], bh.mask[c]);
					if (AGGM == DENSE_AGG_DELTA && (NL == 1 || c == a.agg_ia)) {
						dl[h][e] = d;
					}
					if (POS || AGGM == DENSE_AGG_GENERAL) {
						v[h][e][c] = bh.base[c] + (long long)d;
					}
				}
			}
		}
	}
	(void)CA;
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const uint32_t g = g0 + h * 32 + lane;
		if (POS) {
#pragma unroll
			for (int c = 0; c < NL; c++) {
				if (a.lout[c]) {
					long long *dst = a.lout[c] + obase + g * 2;
					if (FULL) {
						__stcs(reinterpret_cast<longlong2 *>(dst), make_longlong2(v[h][0][c], v[h][1][c]));
					} else {
						// the pair as one 128-bit store when both slots hold a row, else whichever does — predicated,
						// no branches (the compiler's if / else cost the partial iteration twice the full one's instructions)
						asm volatile("{\n\t.reg .pred p, q, b, nb;\n\t"
						             "setp.ne.u32 p, %3, 0;\n\t"
						             "setp.ne.u32 q, %4, 0;\n\t"
						             "and.pred b, p, q;\n\t"
						             "not.pred nb, b;\n\t"
						             "and.pred p, p, nb;\n\t"
						             "and.pred q, q, nb;\n\t"
						             "@b st.global.cs.v2.s64 [%0], {%1, %2};\n\t"
						             "@p st.global.cs.s64 [%0], %1;\n\t"
						             "@q st.global.cs.s64 [%0+8], %2;\n\t}"
						             :
						             : "l"(dst), "l"(v[h][0][c]), "l"(v[h][1][c]), "r"((uint32_t)ok[h][0]), "r"((uint32_t)ok[h][1])
						             : "memory");
					}
				}
			}
		}
#pragma unroll
		for (int e = 0; e < 2; e++) {
			if (AGGM == DENSE_AGG_DELTA) {
				acc.dsum += FULL || ok[h][e] ? dl[h][e] : 0u;
				if (!FULL) {
					acc.rows += ok[h][e] ? 1u : 0u;
				}
			} else if (AGGM == DENSE_AGG_GENERAL) {
				if (FULL || ok[h][e]) {
					const long long x = (NL > 1 && a.agg_ia == 1) ? v[h][e][NL - 1] : v[h][e][0];
					if (a.agg_kind == 1) {
						add128(agg.lo, agg.hi, x);
					} else if (a.agg_kind == 2) {
						const long long y = (NL > 1 && a.agg_ib == 1) ? v[h][e][NL - 1] : v[h][e][0];
						const long long pr = x * y;
						if (__mul64hi(x, y) != (pr >> 63)) {
							agg.overflow = 1;
						}
						add128(agg.lo, agg.hi, pr);
					} else {
						agg.f += __longlong_as_double(x);
					}
				}
			}
		}
	}
	if (AGGM == DENSE_AGG_DELTA && FULL) {
		acc.rows += 4u;
	}
}

// position-ordered write-out of `count` selected rows of ONE pack block (row numbers staged at cbuf[pad ..))
template <int NL, bool POS, bool RAW, int AGGM>
__device__ __forceinline__ void dense_write_out(const DenseProbeArgs &a, const uint16_t *cbuf, uint32_t pad, uint32_t count,
                                                unsigned long long pos0, long long local0, const BlockHdrs<NL> &bh, int lane,
                                                Agg &agg) {
	const unsigned long long obase = pos0 - pad; // output position of staging index 0 (even)
	const uint32_t end = pad + count;
	const uint32_t *cb32 = reinterpret_cast<const uint32_t *>(cbuf);
	DenseAcc acc;
	for (uint32_t g0 = 0; g0 * 2 < end; g0 += 64) {
		if (g0 * 2 >= pad && g0 * 2 + 128 <= end) { // (warp-uniform)
			dense_iter<NL, POS, RAW, AGGM, true>(a, cb32, g0, pad, end, obase, local0, bh, lane, agg, acc);
		} else {
			dense_iter<NL, POS, RAW, AGGM, false>(a, cb32, g0, pad, end, obase, local0, bh, lane, agg, acc);
		}
	}
	if (AGGM == DENSE_AGG_DELTA) { // Σ (base + delta) over this lane's rows of the block = rows·base + Σ delta
		const int ca = NL == 1 ? 0 : a.agg_ia;
		const long long b = NL == 1 ? bh.base[0] : (ca == 1 ? bh.base[NL - 1] : bh.base[0]);
		// rows < 2^11: the 128-bit product from two 32 x 32 → 64-bit multiplies and a sign correction (b = ub − 2^64
		// for negative b) — a generic signed 64 x 64 → 128 multiply costs 35 instructions per block
		const unsigned long long ub = (unsigned long long)b;
		const unsigned long long p0 = (unsigned long long)(uint32_t)ub * acc.rows;
		const unsigned long long p1 = (unsigned long long)(uint32_t)(ub >> 32) * acc.rows;
		const unsigned long long plo = p0 + (p1 << 32);
		long long phi = (long long)((p1 >> 32) + (plo < p0 ? 1ull : 0ull));
		phi -= b < 0 ? (long long)acc.rows : 0ll;
		add128(agg.lo, agg.hi, plo, phi);
		add128(agg.lo, agg.hi, acc.dsum, 0ll);
	}
}

// SB: pack blocks per span (a span = one consumer warp's share of a segment in the scan kernel: 2·WPT blocks)
// RAW: some probed column is a raw array (gathered); AGGM: DENSE_AGG_*; NSTG: ring stages per column (2..4)
template <int SB, int NL, bool POS, bool RAW, int AGGM, int NSTG>
__global__ void __launch_bounds__(kDenseThreads, 3) cubit_probe_dense_kernel(const __grid_constant__ DenseProbeArgs a) {
	extern __shared__ __align__(128) unsigned char dense_smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned char *wbase = dense_smem + (size_t)warp * a.warp_bytes;
	uint16_t *cbuf = reinterpret_cast<uint16_t *>(wbase);
	uint64_t *full = reinterpret_cast<uint64_t *>(wbase + kDenseRowsBytes);
	constexpr uint32_t NST = (uint32_t)NSTG; // stages per column (dense_probe_plan: what shared memory allows)
	uint32_t pk0[NL], pk_stride[NL]; // shared-window address of stage 0 of every column, bytes between its stages
	{
		uint32_t off = kDenseWarpFixed;
#pragma unroll
		for (int c = 0; c < NL; c++) {
			pk0[c] = smem_u32(wbase + off);
			pk_stride[c] = a.stage_bytes[c];
			off += NST * a.stage_bytes[c];
		}
	}
	if (lane == 0) {
		for (uint32_t i = 0; i < (uint32_t)kDenseMaxStages; i++) {
			mbar_init(&full[i], 1);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	const uint32_t gw = blockIdx.x * kDenseWarps + warp, nw = gridDim.x * kDenseWarps;
	// the selection bits of one span as 32-bit pieces: piece (j, lane) = rows [j*1024 + 32*lane, +32) of the span
	auto load_bits = [&](uint32_t sp, uint32_t (&w)[SB]) {
		const uint32_t *src = reinterpret_cast<const uint32_t *>(a.q) + (size_t)sp * (SB * 32);
#pragma unroll
		for (int j = 0; j < SB; j++) {
			w[j] = sp < a.n_span ? __ldg(src + j * 32 + lane) : 0u;
		}
	};
	// lane j holds the headers of the span's block j
	auto load_hdrs = [&](uint32_t sp, uint4 (&h)[NL]) {
		const uint64_t blk = (uint64_t)sp * SB + (uint64_t)lane;
#pragma unroll
		for (int c = 0; c < NL; c++) {
			h[c] = make_uint4(0, 0, 0, 0);
			if (!(RAW && a.lcol[c].raw) && sp < a.n_span && lane < SB && blk < a.n_blk) {
				h[c] = __ldg(reinterpret_cast<const uint4 *>(a.lcol[c].hdr + blk));
			}
		}
	};
	auto nonempty = [&](const uint32_t (&w)[SB]) {
		uint32_t m = 0;
#pragma unroll
		for (int j = 0; j < SB; j++) {
			m |= (__any_sync(0xffffffffu, w[j] != 0u) ? 1u : 0u) << j;
		}
		return m;
	};
	// ring bookkeeping (warp-uniform): stage the next copy goes to, stage / parity of the next block to decode, copies
	// issued but not yet decoded
	uint32_t st_i = 0, st_c = 0, ph_c = 0, inflight = 0;
	// one bulk copy per packed column of block j of the span whose headers are `h`, into stage st_i
	auto issue = [&](const uint4 (&h)[NL], uint32_t j) {
		uint32_t off[NL], bytes[NL], total = 0;
#pragma unroll
		for (int c = 0; c < NL; c++) {
			off[c] = __shfl_sync(0xffffffffu, h[c].z, j);
			bytes[c] = (RAW && a.lcol[c].raw) ? 0u : __shfl_sync(0xffffffffu, h[c].w, j) * (uint32_t)(kPackBlock / 8);
			total += bytes[c];
		}
		if (lane == 0) {
			mbar_arrive_expect_tx(&full[st_i], total);
#pragma unroll
			for (int c = 0; c < NL; c++) {
				if (bytes[c]) {
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					                 pk0[c] + st_i * pk_stride[c]),
					             "l"(a.lcol[c].words + off[c]), "r"(bytes[c]), "r"(smem_u32(&full[st_i]))
					             : "memory");
				}
			}
		}
		st_i = st_i + 1 == NST ? 0 : st_i + 1;
		inflight++;
	};

	// Only the CURRENT span's bits stay in registers.  The next span's bits are loaded once at the top of a span to
	// find its non-empty blocks (whose copies are issued while this span's last blocks are decoded) and loaded again
	// — an L1 / L2 hit by then — when the span becomes current: 8..16 registers per thread for one short stall per span.
	// Copies run up to NST - 1 blocks ahead of the decode, across the span boundary.  Wide columns (3–4 KiB per block)
	// gain 10 % from a ring deeper than two stages; narrow ones do not — their blocks wait for the decode, not for the
	// copy — and keep two (dense_probe_plan; profiles/r2_probe_dense.md).
	uint32_t cur[SB];
	uint4 hc[NL], hn[NL];
	Agg agg;
	uint32_t mi_next; // blocks of the span about to become current whose copies have not been issued yet
	{
		uint32_t nb[SB];
		load_bits(gw, nb);
		mi_next = nonempty(nb);
	}
	load_hdrs(gw, hn);
	for (uint32_t sp = gw; sp < a.n_span; sp += nw) {
		load_bits(sp, cur);
		unsigned long long pos = 0;
		if (POS) {
			pos = __ldg(a.span_excl + sp);
		}
#pragma unroll
		for (int c = 0; c < NL; c++) {
			hc[c] = hn[c];
		}
		uint32_t mi = mi_next, mni; // not yet issued: of this span / of the next one
		{
			uint32_t nb[SB];
			load_bits(sp + nw, nb);
			load_hdrs(sp + nw, hn);
			mni = nonempty(nb);
		}
		uint32_t m = nonempty(cur);
		while (m) {
			const uint32_t j = (uint32_t)__ffs(m) - 1u;
			m &= m - 1u;
			// top the ring up: the blocks of this span in order, then those of the next span (block j itself is always
			// the oldest block not yet decoded, so it is issued first if it was not yet)
			while (inflight < NST) {
				if (mi) {
					const uint32_t ji = (uint32_t)__ffs(mi) - 1u;
					mi &= mi - 1u;
					issue(hc, ji);
				} else if (mni) {
					const uint32_t ji = (uint32_t)__ffs(mni) - 1u;
					mni &= mni - 1u;
					issue(hn, ji);
				} else {
					break;
				}
			}
			uint32_t w = cur[0];
#pragma unroll
			for (int i = 1; i < SB; i++) {
				w = j == (uint32_t)i ? cur[i] : w;
			}
			const uint32_t c = (uint32_t)__popc(w);
			uint32_t incl = c;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) {
					incl += n;
				}
			}
			const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
			const uint32_t pad = POS ? (uint32_t)pos & 1u : 0u;
			stage_word(cbuf, pad + incl - c, w & 0xffffu, w >> 16, (uint32_t)lane * 32u, 0u, 16u);
			BlockHdrs<NL> bh;
#pragma unroll
			for (int cc = 0; cc < NL; cc++) {
				const uint32_t blo = __shfl_sync(0xffffffffu, hc[cc].x, j), bhi = __shfl_sync(0xffffffffu, hc[cc].y, j);
				bh.base[cc] = (long long)(((unsigned long long)bhi << 32) | blo);
				bh.width[cc] = __shfl_sync(0xffffffffu, hc[cc].w, j);
				bh.mask[cc] = bh.width[cc] >= 32u ? 0xffffffffu : (1u << bh.width[cc]) - 1u;
				bh.pk[cc] = pk0[cc] + st_c * pk_stride[cc];
			}
			__syncwarp();
			mbar_wait(&full[st_c], ph_c);
			dense_write_out<NL, POS, RAW, AGGM>(a, cbuf, pad, total, pos, ((long long)sp * SB + j) * kPackBlock, bh, lane, agg);
			__syncwarp(); // every lane is done with the staging row and with stage st_c before either is refilled
			pos += total;
			inflight--;
			if (++st_c == NST) {
				st_c = 0;
				ph_c ^= 1u;
			}
		}
		mi_next = mni;
	}
	if (AGGM != DENSE_AGG_NONE) {
		agg_flush_warp(agg, a.hdr, lane);
	}
}

// ------------------------------------------------------------------------------------------ launch
template <int SB, int NL, bool POS, bool RAW, int AGGM, int NSTG>
static cudaError_t launch_dense_n(const DenseProbeArgs &args, int sm_count, cudaStream_t stream) {
	auto kern = cubit_probe_dense_kernel<SB, NL, POS, RAW, AGGM, NSTG>;
	const size_t smem = (size_t)args.warp_bytes * kDenseWarps;
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	static size_t smem_set[64] = {}; // largest dynamic shared memory this instance was configured for, per device
	if (smem > smem_set[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) {
			return e;
		}
		smem_set[dev] = smem;
	}
	int b = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kDenseThreads, smem);
	if (e != cudaSuccess) {
		return e;
	}
	if (b < 1) {
		return cudaErrorInvalidConfiguration;
	}
	long long grid = (long long)sm_count * b; // resident CTAs; spans strided over all warps of the grid
	const long long need = ((long long)args.n_span + kDenseWarps - 1) / kDenseWarps;
	grid = grid > need ? need : grid;
	grid = grid < 1 ? 1 : grid;
	kern<<<(unsigned)grid, kDenseThreads, smem, stream>>>(args);
	return cudaGetLastError();
}

template <int SB, int NL, bool POS, bool RAW, int AGGM>
static cudaError_t launch_dense_t(const DenseProbeArgs &args, int sm_count, cudaStream_t stream) {
	switch (args.n_stages) {
	case 2:
		return launch_dense_n<SB, NL, POS, RAW, AGGM, 2>(args, sm_count, stream);
	case 3:
		return launch_dense_n<SB, NL, POS, RAW, AGGM, 3>(args, sm_count, stream);
	case 4:
		return launch_dense_n<SB, NL, POS, RAW, AGGM, 4>(args, sm_count, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

template <int SB, int NL, bool POS, bool RAW>
static cudaError_t launch_dense_agg(const DenseProbeArgs &args, int sm_count, cudaStream_t stream) {
	if (args.agg_kind == 0) {
		if (!POS) {
			return cudaSuccess; // nothing to write, nothing to add
		}
		return launch_dense_t<SB, NL, POS, RAW, DENSE_AGG_NONE>(args, sm_count, stream);
	}
	const int ca = NL == 1 ? 0 : args.agg_ia;
	if (args.agg_kind == 1 && !args.lcol[ca].raw) { // SUM over a bit-packed column
		return launch_dense_t<SB, NL, POS, RAW, DENSE_AGG_DELTA>(args, sm_count, stream);
	}
	return launch_dense_t<SB, NL, POS, RAW, DENSE_AGG_GENERAL>(args, sm_count, stream);
}

template <int SB>
static cudaError_t launch_dense_sb(const DenseProbeArgs &args, bool positions, int sm_count, cudaStream_t stream) {
	bool raw = false;
	for (int c = 0; c < args.n_load; c++) {
		raw |= args.lcol[c].raw != nullptr;
	}
	if (args.n_load == 1 && !raw) {
		return positions ? launch_dense_agg<SB, 1, true, false>(args, sm_count, stream)
		                 : launch_dense_agg<SB, 1, false, false>(args, sm_count, stream);
	}
	if (args.n_load == 2) {
		if (raw) {
			return positions ? launch_dense_agg<SB, 2, true, true>(args, sm_count, stream)
			                 : launch_dense_agg<SB, 2, false, true>(args, sm_count, stream);
		}
		return positions ? launch_dense_agg<SB, 2, true, false>(args, sm_count, stream)
		                 : launch_dense_agg<SB, 2, false, false>(args, sm_count, stream);
	}
	return cudaErrorInvalidValue;
}

bool dense_probe_plan(DenseProbeArgs &args, const uint32_t *max_width) {
	uint32_t per_stage = 0;
	bool any_packed = false;
	for (int c = 0; c < args.n_load; c++) {
		args.stage_bytes[c] = 0;
		if (args.lcol[c].raw) {
			continue;
		}
		if (max_width[c] > 32u) {
			return false;
		}
		any_packed = true;
		// one block's payload + 16 spare bytes (the decoder reads one 32-bit word past a value)
		args.stage_bytes[c] = max_width[c] * (uint32_t)(kPackBlock / 8) + 16u;
		per_stage += args.stage_bytes[c];
	}
	// stages per column: as deep a ring as shared memory allows with 3 CTAs per SM, else with 2 (227 KiB per SM, 1 KiB
	// reserved per CTA); the copies run stages - 1 blocks ahead of the decode
	auto warp_bytes = [&](uint32_t nst) { return ((uint32_t)kDenseWarpFixed + nst * per_stage + 127u) & ~127u; };
	auto fits = [&](uint32_t ctas, uint32_t nst) {
		return (uint64_t)warp_bytes(nst) * kDenseWarps <= (227u * 1024u - ctas * 1024u) / ctas;
	};
	// (narrow columns gain nothing from a deeper ring — their blocks are ≤ 2 KiB and the decode, not the copy, is what
	// a block waits for: 0.49 / 0.70 / 0.97 ms with 2 stages against 0.51 / 0.73 / 0.99 ms with 4 on the 10-bit payload,
	// while the 24-bit payload goes from 0.68 / 0.92 / 1.30 to 0.61 / 0.84 / 1.16 ms; profiles/r2_probe_dense.md)
	uint32_t nst = 2;
	if (per_stage < 2048u) {
		nst = 2;
	} else if (fits(3, 4)) {
		nst = 4;
	} else if (fits(3, 3)) {
		nst = 3;
	} else if (fits(2, 4)) {
		nst = 4;
	} else if (fits(2, 3)) {
		nst = 3;
	}
	args.n_stages = nst;
	args.warp_bytes = warp_bytes(nst);
	return any_packed && args.n_load >= 1 && args.n_load <= kMaxFusedCols;
}

cudaError_t launch_probe_dense(const DenseProbeArgs &args, uint32_t seg_words, bool positions, int sm_count,
                               cudaStream_t stream) {
	switch (seg_words) {
	case 512:
		return launch_dense_sb<4>(args, positions, sm_count, stream);
	case 1024:
		return launch_dense_sb<8>(args, positions, sm_count, stream);
	case 2048:
		return launch_dense_sb<16>(args, positions, sm_count, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace cubit
