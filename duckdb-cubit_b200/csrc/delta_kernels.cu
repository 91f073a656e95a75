// delta_kernels.cu — device-side maintenance of the pending-delta lists (the index side of
// BoundIndex::Append/Delete/Insert, src/include/duckdb/execution/index/bound_index.hpp:71-97).
//
// The pending deltas of one index are a CSR over keys (value * n_seg + segment) of 16-byte entries
// {word-in-segment, key, 64-bit mask}; the scan kernel bulk-copies the entries of (value, segment) next to the
// staged segment and XORs them in with shared-memory atomics, so entries of one key are unordered and may repeat
// a word.  Ingesting n newly flipped (value, row) pairs is therefore a counting sort by key with NO general sort
// and no deduplication:
//     1. histogram of the new pairs per key                      (cubit_delta_hist_kernel, atomicAdd)
//     2. new offsets = exclusive scan of (old count + new count) (three small scan kernels)
//     3. old entries move to their key's new range               (cubit_delta_move_kernel, key kept in the entry)
//     4. new entries are scattered behind them                   (cubit_delta_scatter_kernel, atomicSub cursor)
// all enqueued on the table's kernel stream: scans before the call read the old lists, scans after it the new
// ones, and the host never waits.  Bytes moved: 16·(m + n) + 8·n_keys — HBM-bound, tens of microseconds for the
// 1 % / 10 M-row delta of SURVEY §8d config 4.
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {

namespace {

constexpr int kThreads = 256;
constexpr int kScanItems = 16;                      // keys per thread in the scan kernels
constexpr int kScanChunk = kThreads * kScanItems;   // keys per CTA

__device__ __forceinline__ uint32_t key_of(uint32_t value, long long row, uint32_t n_seg, uint32_t seg_shift) {
	return value * n_seg + (uint32_t)((unsigned long long)row >> seg_shift);
}

__global__ void __launch_bounds__(kThreads) cubit_delta_hist_kernel(const long long *__restrict__ rows,
                                                                    const uint32_t *__restrict__ values, uint32_t one_value,
                                                                    uint64_t n, uint32_t n_seg, uint32_t seg_shift,
                                                                    uint32_t *__restrict__ cnt) {
	for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
		const uint32_t v = values ? __ldg(values + i) : one_value;
		atomicAdd(cnt + key_of(v, __ldg(rows + i), n_seg, seg_shift), 1u);
	}
}

// count of key `k` in the merged list: surviving old entries + new ones
__device__ __forceinline__ uint32_t merged_count(const uint32_t *old_off, const uint32_t *cnt, uint64_t k, uint32_t n_seg,
                                                 uint32_t drop_value) {
	uint32_t c = cnt ? cnt[k] : 0u;
	if (old_off && (uint32_t)(k / n_seg) != drop_value) {
		c += old_off[k + 1] - old_off[k];
	}
	return c;
}

__global__ void __launch_bounds__(kThreads) cubit_delta_scan_reduce_kernel(const uint32_t *__restrict__ old_off,
                                                                           const uint32_t *__restrict__ cnt, uint64_t n_keys,
                                                                           uint32_t n_seg, uint32_t drop_value,
                                                                           uint32_t *__restrict__ block_sum) {
	__shared__ uint32_t wsum[kThreads / 32];
	const uint64_t k0 = (uint64_t)blockIdx.x * kScanChunk;
	uint32_t s = 0;
	for (int j = 0; j < kScanItems; j++) {
		const uint64_t k = k0 + (uint64_t)j * kThreads + threadIdx.x;
		if (k < n_keys) {
			s += merged_count(old_off, cnt, k, n_seg, drop_value);
		}
	}
	s = __reduce_add_sync(0xffffffffu, s);
	if ((threadIdx.x & 31) == 0) {
		wsum[threadIdx.x >> 5] = s;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t tot = 0;
		for (int w = 0; w < kThreads / 32; w++) {
			tot += wsum[w];
		}
		block_sum[blockIdx.x] = tot;
	}
}

// one CTA: exclusive scan of the per-CTA sums in place (n_blocks ≤ 1024 * 64)
__global__ void __launch_bounds__(1024) cubit_delta_scan_sums_kernel(uint32_t *__restrict__ block_sum, uint32_t n_blocks) {
	__shared__ uint32_t wtot[32];
	const uint32_t per = (n_blocks + 1023u) / 1024u;
	const uint32_t b0 = threadIdx.x * per;
	uint32_t mine = 0;
	for (uint32_t j = 0; j < per; j++) {
		if (b0 + j < n_blocks) {
			mine += block_sum[b0 + j];
		}
	}
	uint32_t incl = mine;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) {
			incl += y;
		}
	}
	if (lane == 31) {
		wtot[warp] = incl;
	}
	__syncthreads();
	uint32_t base = incl - mine;
	for (int w = 0; w < warp; w++) {
		base += wtot[w];
	}
	for (uint32_t j = 0; j < per; j++) {
		if (b0 + j < n_blocks) {
			const uint32_t x = block_sum[b0 + j];
			block_sum[b0 + j] = base;
			base += x;
		}
	}
}

__global__ void __launch_bounds__(kThreads) cubit_delta_scan_write_kernel(const uint32_t *__restrict__ old_off,
                                                                          const uint32_t *__restrict__ cnt, uint64_t n_keys,
                                                                          uint32_t n_seg, uint32_t drop_value,
                                                                          const uint32_t *__restrict__ block_base,
                                                                          uint32_t *__restrict__ new_off) {
	// thread t owns kScanItems CONSECUTIVE keys of the CTA's chunk (a blocked arrangement keeps the scan trivial;
	// the loads are strided but this kernel moves only 12 bytes per key)
	__shared__ uint32_t wtot[kThreads / 32];
	const uint64_t k0 = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * kScanItems;
	uint32_t c[kScanItems], mine = 0;
#pragma unroll
	for (int j = 0; j < kScanItems; j++) {
		c[j] = k0 + j < n_keys ? merged_count(old_off, cnt, k0 + j, n_seg, drop_value) : 0u;
		mine += c[j];
	}
	uint32_t incl = mine;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) {
			incl += y;
		}
	}
	if (lane == 31) {
		wtot[warp] = incl;
	}
	__syncthreads();
	uint32_t base = block_base[blockIdx.x] + incl - mine;
	for (int w = 0; w < warp; w++) {
		base += wtot[w];
	}
#pragma unroll
	for (int j = 0; j < kScanItems; j++) {
		if (k0 + j < n_keys) {
			new_off[k0 + j] = base;
		}
		base += c[j];
		if (k0 + j + 1 == n_keys) {
			new_off[n_keys] = base; // grand total closes the CSR
		}
	}
}

__global__ void __launch_bounds__(kThreads) cubit_delta_move_kernel(const DeltaEnt *__restrict__ old_ent, uint64_t m,
                                                                    const uint32_t *__restrict__ old_off,
                                                                    const uint32_t *__restrict__ new_off, uint32_t n_seg,
                                                                    uint32_t drop_value, DeltaEnt *__restrict__ new_ent) {
	for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < m; i += (uint64_t)gridDim.x * kThreads) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(old_ent + i));
		const uint32_t key = raw.y;
		if (key / n_seg == drop_value) {
			continue;
		}
		const uint32_t pos = new_off[key] + ((uint32_t)i - old_off[key]);
		*reinterpret_cast<uint4 *>(new_ent + pos) = raw;
	}
}

__global__ void __launch_bounds__(kThreads) cubit_delta_scatter_kernel(const long long *__restrict__ rows,
                                                                       const uint32_t *__restrict__ values, uint32_t one_value,
                                                                       uint64_t n, uint32_t n_seg, uint32_t seg_shift,
                                                                       const uint32_t *__restrict__ new_off,
                                                                       uint32_t *__restrict__ cnt, DeltaEnt *__restrict__ new_ent) {
	const uint32_t seg_mask = (1u << seg_shift) - 1u;
	for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
		const uint32_t v = values ? __ldg(values + i) : one_value;
		const long long row = __ldg(rows + i);
		const uint32_t key = key_of(v, row, n_seg, seg_shift);
		const uint32_t c = atomicSub(cnt + key, 1u); // 1-based slot from the END of the key's range
		const uint32_t pos = new_off[key + 1] - c;
		uint4 e;
		e.x = ((uint32_t)row & seg_mask) >> 6;
		e.y = key;
		const unsigned long long mask = 1ull << ((unsigned)row & 63u);
		e.z = (uint32_t)mask;
		e.w = (uint32_t)(mask >> 32);
		*reinterpret_cast<uint4 *>(new_ent + pos) = e;
	}
}

// merge-back: B_v ^= D_v for every pending entry (two entries may hit one word: atomic)
__global__ void __launch_bounds__(kThreads) cubit_delta_apply_kernel(const DeltaEnt *__restrict__ ent, uint64_t e0, uint64_t e1,
                                                                     uint32_t n_seg, uint32_t seg_words,
                                                                     unsigned long long *__restrict__ bits,
                                                                     uint64_t words_per_bv, uint32_t value_base) {
	for (uint64_t i = e0 + (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < e1; i += (uint64_t)gridDim.x * kThreads) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(ent + i));
		const uint32_t v = raw.y / n_seg, seg = raw.y % n_seg;
		atomicXor(bits + (uint64_t)(v - value_base) * words_per_bv + (uint64_t)seg * seg_words + raw.x,
		          ((unsigned long long)raw.w << 32) | raw.z);
	}
}

// the segment count changed (append): offsets re-keyed, entries keep their order
__global__ void __launch_bounds__(kThreads) cubit_delta_restride_off_kernel(const uint32_t *__restrict__ old_off,
                                                                            uint32_t card, uint32_t old_n_seg, uint32_t new_n_seg,
                                                                            uint32_t *__restrict__ new_off) {
	const uint64_t n_new = (uint64_t)card * new_n_seg;
	for (uint64_t k = (uint64_t)blockIdx.x * kThreads + threadIdx.x; k <= n_new; k += (uint64_t)gridDim.x * kThreads) {
		if (k == n_new) {
			new_off[k] = old_off[(uint64_t)card * old_n_seg];
			continue;
		}
		const uint32_t v = (uint32_t)(k / new_n_seg), s = (uint32_t)(k % new_n_seg);
		new_off[k] = s < old_n_seg ? old_off[(uint64_t)v * old_n_seg + s] : old_off[(uint64_t)(v + 1) * old_n_seg];
	}
}

__global__ void __launch_bounds__(kThreads) cubit_delta_restride_ent_kernel(DeltaEnt *__restrict__ ent, uint64_t m,
                                                                            uint32_t old_n_seg, uint32_t new_n_seg) {
	for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < m; i += (uint64_t)gridDim.x * kThreads) {
		const uint32_t key = ent[i].pad;
		ent[i].pad = (key / old_n_seg) * new_n_seg + key % old_n_seg;
	}
}

__global__ void cubit_delta_value_offsets_kernel(const uint32_t *__restrict__ off, uint32_t n_seg, uint32_t card,
                                                uint32_t *__restrict__ out) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v <= card) {
		out[v] = off[(uint64_t)v * n_seg];
	}
}

unsigned grid_for(uint64_t n, int sm_count) {
	uint64_t g = (n + kThreads - 1) / kThreads;
	const uint64_t cap = (uint64_t)sm_count * 16;
	if (g > cap) {
		g = cap;
	}
	return g < 1 ? 1u : (unsigned)g;
}

} // namespace

cudaError_t launch_delta_ingest(const DeltaIngest &a, int sm_count, cudaStream_t stream, int *n_launches) {
	int launches = 0;
	const uint32_t n_blocks = (uint32_t)((a.n_keys + kScanChunk - 1) / kScanChunk);
	if (n_blocks > 1024u * 64u) {
		return cudaErrorInvalidValue;
	}
	if (a.n_new) {
		cubit_delta_hist_kernel<<<grid_for(a.n_new, sm_count), kThreads, 0, stream>>>(a.rows, a.values, a.one_value, a.n_new,
		                                                                              a.n_seg, a.seg_shift, a.cnt);
		launches++;
	}
	cubit_delta_scan_reduce_kernel<<<n_blocks, kThreads, 0, stream>>>(a.old_off, a.n_new ? a.cnt : nullptr, a.n_keys, a.n_seg,
	                                                                 a.drop_value, a.block_sum);
	cubit_delta_scan_sums_kernel<<<1, 1024, 0, stream>>>(a.block_sum, n_blocks);
	cubit_delta_scan_write_kernel<<<n_blocks, kThreads, 0, stream>>>(a.old_off, a.n_new ? a.cnt : nullptr, a.n_keys, a.n_seg,
	                                                                a.drop_value, a.block_sum, a.new_off);
	launches += 3;
	if (a.n_old) {
		cubit_delta_move_kernel<<<grid_for(a.n_old, sm_count), kThreads, 0, stream>>>(a.old_ent, a.n_old, a.old_off, a.new_off,
		                                                                              a.n_seg, a.drop_value, a.new_ent);
		launches++;
	}
	if (a.n_new) {
		cubit_delta_scatter_kernel<<<grid_for(a.n_new, sm_count), kThreads, 0, stream>>>(
		    a.rows, a.values, a.one_value, a.n_new, a.n_seg, a.seg_shift, a.new_off, a.cnt, a.new_ent);
		launches++;
	}
	if (n_launches) {
		*n_launches = launches;
	}
	return cudaGetLastError();
}

cudaError_t launch_delta_apply(const DeltaEnt *ent, uint64_t e0, uint64_t e1, uint32_t n_seg, uint32_t seg_words,
                               uint64_t *bits, uint64_t words_per_bv, uint32_t value_base, int sm_count,
                               cudaStream_t stream) {
	if (e1 <= e0) {
		return cudaSuccess;
	}
	cubit_delta_apply_kernel<<<grid_for(e1 - e0, sm_count), kThreads, 0, stream>>>(
	    ent, e0, e1, n_seg, seg_words, reinterpret_cast<unsigned long long *>(bits), words_per_bv, value_base);
	return cudaGetLastError();
}

cudaError_t launch_delta_value_offsets(const uint32_t *off, uint32_t n_seg, uint32_t card, uint32_t *out, cudaStream_t stream) {
	cubit_delta_value_offsets_kernel<<<(card + 1 + 255) / 256, 256, 0, stream>>>(off, n_seg, card, out);
	return cudaGetLastError();
}

cudaError_t launch_delta_restride(const uint32_t *old_off, uint32_t *new_off, DeltaEnt *ent, uint64_t n_ent, uint32_t card,
                                  uint32_t old_n_seg, uint32_t new_n_seg, int sm_count, cudaStream_t stream) {
	cubit_delta_restride_off_kernel<<<grid_for((uint64_t)card * new_n_seg + 1, sm_count), kThreads, 0, stream>>>(
	    old_off, card, old_n_seg, new_n_seg, new_off);
	if (n_ent) {
		cubit_delta_restride_ent_kernel<<<grid_for(n_ent, sm_count), kThreads, 0, stream>>>(ent, n_ent, old_n_seg, new_n_seg);
	}
	return cudaGetLastError();
}

} // namespace cubit
