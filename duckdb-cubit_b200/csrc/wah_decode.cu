// wah_decode.cu — GPU expansion of a WAH-compressed bitvector (FastBit ibis::bitvector form, the way the
// upstream CUBIT library stores its value bitvectors) into the verbatim, segment-padded bitvector the scan
// kernel streams (input side of §8a A1).
//
// The compressed words cross PCIe as they are.  Word i covers g_i = 1 (literal) or count (fill) groups of 31
// bits, starting at bit 31 * Σ_{j<i} g_j: the host supplies that prefix once per block of 1024 words (it walks
// the words anyway to validate them), the CTA finishes it with a thread-local + warp-shuffle + 8-warp scan.
//   literal : its 31 bits are stored first-bit-most-significant → __brev, then OR-ed into the one or two
//             64-bit output words they fall into (atomicOr: two neighbouring literals share an output word)
//   0-fill  : nothing to do (the destination is zeroed first)
//   1-fill  : written by the whole warp — plain 64-bit stores for the fully covered words, atomicOr for the
//             two boundary words
// Not on the query path: one launch per uploaded bitvector.  HBM/atomic bound; bytes = 4 per compressed word
// + 8 per touched output word (+ the memset of the destination).
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {

namespace {

constexpr int kWahThreads = 256;
constexpr int kWahPerThread = 4;

__device__ __forceinline__ void place_bits(unsigned long long *out, unsigned long long pos, unsigned long long bits,
                                           uint32_t nb) {
	const unsigned long long w = pos >> 6;
	const uint32_t sh = (uint32_t)(pos & 63u);
	if (bits) {
		atomicOr(out + w, bits << sh);
		if (sh + nb > 64u && (bits >> (64u - sh))) {
			atomicOr(out + w + 1, bits >> (64u - sh));
		}
	}
}

__device__ __forceinline__ void warp_fill_ones(unsigned long long *out, unsigned long long start,
                                               unsigned long long len, int lane) {
	const unsigned long long endb = start + len - 1; // last bit of the run
	const unsigned long long first = start >> 6, last = endb >> 6;
	const unsigned long long head = ~0ull << (start & 63u), tail = ~0ull >> (63u - (endb & 63u));
	if (first == last) {
		if (lane == 0) {
			atomicOr(out + first, head & tail);
		}
		return;
	}
	if (lane == 0) {
		atomicOr(out + first, head);
	} else if (lane == 1) {
		atomicOr(out + last, tail);
	}
	for (unsigned long long w = first + 1 + lane; w < last; w += 32) {
		out[w] = ~0ull; // covered by this fill alone
	}
}

__global__ void __launch_bounds__(kWahThreads) cubit_wah_expand_kernel(const uint32_t *__restrict__ wah, uint64_t n_wah,
                                                                      const unsigned long long *__restrict__ block_group0,
                                                                      unsigned long long total_groups,
                                                                      uint32_t active_val, uint32_t active_nbits,
                                                                      unsigned long long *__restrict__ out) {
	__shared__ unsigned long long warp_tot[kWahThreads / 32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint64_t i0 = ((uint64_t)blockIdx.x * kWahThreads + threadIdx.x) * kWahPerThread;
	uint32_t w[kWahPerThread];
	unsigned long long g[kWahPerThread], mine = 0;
	if (i0 + kWahPerThread <= n_wah) { // 16-byte aligned: the staged array starts on an allocation boundary
		const uint4 v = __ldg(reinterpret_cast<const uint4 *>(wah + i0));
		w[0] = v.x;
		w[1] = v.y;
		w[2] = v.z;
		w[3] = v.w;
	} else {
#pragma unroll
		for (int j = 0; j < kWahPerThread; j++) {
			w[j] = i0 + j < n_wah ? __ldg(wah + i0 + j) : 0x80000000u; // padding: a zero-length 0-fill
		}
	}
#pragma unroll
	for (int j = 0; j < kWahPerThread; j++) {
		g[j] = (w[j] & 0x80000000u) ? (unsigned long long)(w[j] & 0x3fffffffu) : 1ull;
		mine += g[j];
	}
	// exclusive scan of the per-thread group counts over the CTA
	unsigned long long incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) {
			incl += y;
		}
	}
	if (lane == 31) {
		warp_tot[warp] = incl;
	}
	__syncthreads();
	unsigned long long base = block_group0[blockIdx.x] + incl - mine;
#pragma unroll
	for (int k = 0; k < kWahThreads / 32; k++) {
		base += k < warp ? warp_tot[k] : 0ull;
	}
	unsigned long long gstart[kWahPerThread];
#pragma unroll
	for (int j = 0; j < kWahPerThread; j++) {
		gstart[j] = base;
		base += g[j];
		if (!(w[j] & 0x80000000u)) { // literal: first bit is the most significant of the 31
			place_bits(out, gstart[j] * 31ull, (unsigned long long)(__brev(w[j]) >> 1), 31u);
		}
	}
#pragma unroll
	for (int j = 0; j < kWahPerThread; j++) {
		const bool ones = (w[j] & 0xc0000000u) == 0xc0000000u && g[j] != 0;
		unsigned m = __ballot_sync(0xffffffffu, ones);
		while (m) {
			const int src = __ffs(m) - 1;
			const unsigned long long s = __shfl_sync(0xffffffffu, gstart[j], src) * 31ull;
			const unsigned long long l = __shfl_sync(0xffffffffu, g[j], src) * 31ull;
			warp_fill_ones(out, s, l, lane);
			m &= m - 1;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0 && active_nbits) {
		place_bits(out, total_groups * 31ull, (unsigned long long)(__brev(active_val) >> (32u - active_nbits)),
		           active_nbits);
	}
}

} // namespace

cudaError_t launch_wah_expand(const uint32_t *wah, uint64_t n_wah, const unsigned long long *block_group0,
                              unsigned long long total_groups, uint32_t active_val, uint32_t active_nbits,
                              unsigned long long *out, cudaStream_t stream) {
	uint64_t blocks = (n_wah + kWahBlockWords - 1) / kWahBlockWords;
	if (blocks == 0) {
		blocks = 1; // only an active word
	}
	cubit_wah_expand_kernel<<<(unsigned)blocks, kWahThreads, 0, stream>>>(wah, n_wah, block_group0, total_groups,
	                                                                     active_val, active_nbits, out);
	return cudaGetLastError();
}

} // namespace cubit
