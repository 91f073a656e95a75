// column_decode.cu — GPU decode of the reference's on-disk column segments into the HBM-resident column
// (the input side of the probe, §8a A3).
//
// The reference stores lineitem's numeric columns BitPacking-compressed (src/storage/compression/bitpacking.cpp):
// a segment is a run of 2048-value metadata groups, each in one of four modes.  The CPU scan decodes group by
// group (BitpackingScanPartial, bitpacking.cpp:776-860; a point fetch decodes a 32-value group per row,
// BitpackingFetchRow :879-950, and for DELTA_FOR everything before the row in its 2048-group).  Here the
// compressed bytes cross PCIe as they are and ONE CTA decodes ONE metadata group:
//   * the packed words of the group are staged in shared memory with coalesced 32-bit loads (group data is only
//     4-byte aligned inside a segment: headers and packed runs of earlier groups have arbitrary 4-byte sizes);
//   * value i is the `width` bits at bit position i*width of the group's little-endian 32-bit word stream
//     (duckdb_fastpforlib::fastunpack, third_party/fastpforlib/bitpacking.cpp:173-226) — extracted from three
//     shared-memory words, no per-width code;
//   * + frame of reference (wrapping, ApplyFrameOfReference :583-593); DELTA_FOR then needs the running sum
//     over the whole group seeded with delta_offset (DeltaDecode :596-620): warp w owns values
//     [256w, 256w+256) as 8 rounds of 32 consecutive values (5-step shuffle scan per round, carry between
//     rounds), then an exclusive scan over the 8 warp totals through shared memory;
//   * stores are 256 contiguous bytes per warp instruction.
// Roofline: HBM (and PCIe before it).  Algorithmic bytes per group: packed bytes read + 2048*sizeof(T) written.
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {

namespace {

constexpr int kDecodeThreads = 256;
constexpr int kGroup = 2048; // BITPACKING_METADATA_GROUP_SIZE (bitpacking.cpp:22)

template <typename T>
struct Bits;
template <>
struct Bits<uint64_t> {
	static constexpr uint32_t kHdrWords = 2; // 32-bit words per header field
	static constexpr uint32_t kMaxWords = kGroup * 64 / 32;
	__device__ static uint64_t field(const uint32_t *w, int i) { // header field i (4-byte aligned only)
		return ((uint64_t)__ldg(w + 2 * i + 1) << 32) | __ldg(w + 2 * i);
	}
	__device__ static uint64_t extract(const uint32_t *s, uint32_t i, uint32_t width) {
		const uint32_t pos = i * width, wi = pos >> 5, sh = pos & 31u;
		const uint64_t a = ((uint64_t)s[wi + 1] << 32) | s[wi];
		uint64_t v = a >> sh;
		if (sh) {
			v |= (uint64_t)s[wi + 2] << (64u - sh);
		}
		return width < 64 ? v & ((1ull << width) - 1ull) : v;
	}
};
template <>
struct Bits<uint32_t> {
	static constexpr uint32_t kHdrWords = 1;
	static constexpr uint32_t kMaxWords = kGroup * 32 / 32;
	__device__ static uint32_t field(const uint32_t *w, int i) {
		return __ldg(w + i);
	}
	__device__ static uint32_t extract(const uint32_t *s, uint32_t i, uint32_t width) {
		const uint32_t pos = i * width, wi = pos >> 5, sh = pos & 31u;
		const uint64_t a = ((uint64_t)s[wi + 1] << 32) | s[wi];
		const uint32_t v = (uint32_t)(a >> sh);
		return width < 32 ? v & ((1u << width) - 1u) : v;
	}
};

template <typename T>
__global__ void __launch_bounds__(kDecodeThreads) cubit_bp_decode_kernel(const uint8_t *__restrict__ blob,
                                                                         const BpGroup *__restrict__ groups,
                                                                         T *__restrict__ out) {
	__shared__ uint32_t s[Bits<T>::kMaxWords + 2];
	__shared__ T warp_tot[kDecodeThreads / 32];
	const BpGroup g = groups[blockIdx.x];
	const uint32_t *w32 = reinterpret_cast<const uint32_t *>(blob + g.data_off);
	T *dst = out + g.row0;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (g.mode == BP_CONSTANT) { // also a whole Constant-compressed segment: n may exceed one group
		const T c = Bits<T>::field(w32, 0);
		for (uint32_t i = threadIdx.x; i < g.n; i += kDecodeThreads) {
			dst[i] = c;
		}
		return;
	}
	if (g.mode == BP_CONSTANT_DELTA) { // value i = frame + i * delta (bitpacking.cpp:815-826)
		const T frame = Bits<T>::field(w32, 0), delta = Bits<T>::field(w32, 1);
		for (uint32_t i = threadIdx.x; i < g.n; i += kDecodeThreads) {
			dst[i] = frame + delta * (T)i;
		}
		return;
	}
	const bool is_delta = g.mode == BP_DELTA_FOR;
	const T frame = Bits<T>::field(w32, 0);
	const uint32_t width = (uint32_t)Bits<T>::field(w32, 1) & 0xffu;
	const T delta_offset = is_delta ? Bits<T>::field(w32, 2) : (T)0;
	const uint32_t *packed = w32 + (is_delta ? 3 : 2) * Bits<T>::kHdrWords;
	const uint32_t n_words = ((g.n + 31u) >> 5) * width; // whole 32-value groups are stored (GetRequiredSize)
	for (uint32_t i = threadIdx.x; i < n_words; i += kDecodeThreads) {
		s[i] = __ldg(packed + i);
	}
	if (threadIdx.x < 2) {
		s[n_words + threadIdx.x] = 0; // the extractor reads up to two words past the last value's first word
	}
	__syncthreads();

	T v[8];
	T carry = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const uint32_t i = (uint32_t)warp * 256u + (uint32_t)j * 32u + (uint32_t)lane;
		T x = i < g.n ? (T)(Bits<T>::extract(s, i, width) + frame) : (T)0;
		if (is_delta) { // inclusive scan of this round, on top of the warp's previous rounds
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const T y = __shfl_up_sync(0xffffffffu, x, d);
				if (lane >= d) {
					x += y;
				}
			}
			x += carry;
			carry = __shfl_sync(0xffffffffu, x, 31);
		}
		v[j] = x;
	}
	T base = delta_offset;
	if (is_delta) {
		if (lane == 0) {
			warp_tot[warp] = carry;
		}
		__syncthreads();
#pragma unroll
		for (int w = 0; w < kDecodeThreads / 32; w++) {
			base += w < warp ? warp_tot[w] : (T)0;
		}
	}
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const uint32_t i = (uint32_t)warp * 256u + (uint32_t)j * 32u + (uint32_t)lane;
		if (i < g.n) {
			dst[i] = is_delta ? (T)(v[j] + base) : v[j];
		}
	}
}

// ---- RLE segments (src/storage/compression/rle.cpp).  The CPU scan walks the runs sequentially
// (RLEScanPartialInternal, :338-364) and a point fetch skips from the start of the segment (RLEFetchRow, :380-392).
// Here one CTA takes a tile of ≤ 1024 runs: inclusive block scan of the 16-bit run lengths → run end positions in
// shared memory; then the tile's rows are produced in order, 256 consecutive rows per step, each thread locating
// its run with a 10-step binary search over the shared array, so the stores are fully coalesced whatever the run
// lengths are.  Algorithmic bytes: (sizeof(T) + 2) per run read + sizeof(T) per row written.
template <typename T>
__global__ void __launch_bounds__(kDecodeThreads) cubit_rle_decode_kernel(const uint8_t *__restrict__ blob,
                                                                          const RleTile *__restrict__ tiles,
                                                                          T *__restrict__ out) {
	constexpr int kPer = kRleTileRuns / kDecodeThreads; // runs per thread
	__shared__ uint32_t ends[kRleTileRuns];
	__shared__ uint32_t warp_tot[kDecodeThreads / 32];
	const RleTile tl = tiles[blockIdx.x];
	const uint16_t *cnt = reinterpret_cast<const uint16_t *>(blob + tl.cnt_off);
	const T *val = reinterpret_cast<const T *>(blob + tl.val_off);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t c[kPer], sum = 0;
#pragma unroll
	for (int j = 0; j < kPer; j++) {
		const uint32_t r = threadIdx.x * kPer + j;
		c[j] = r < tl.n_runs ? (uint32_t)__ldg(cnt + r) : 0u;
		sum += c[j];
	}
	uint32_t incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) {
			incl += y;
		}
	}
	if (lane == 31) {
		warp_tot[warp] = incl;
	}
	__syncthreads();
	uint32_t run_end = incl - sum;
#pragma unroll
	for (int w = 0; w < kDecodeThreads / 32; w++) {
		run_end += w < warp ? warp_tot[w] : 0u;
	}
#pragma unroll
	for (int j = 0; j < kPer; j++) {
		run_end += c[j];
		ends[threadIdx.x * kPer + j] = run_end; // runs past n_runs repeat the total: never selected below
	}
	__syncthreads();
	T *dst = out + tl.row0;
	for (uint32_t r = threadIdx.x; r < tl.n_rows; r += kDecodeThreads) {
		uint32_t lo = 0, hi = tl.n_runs - 1; // first run whose end is > r
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (ends[mid] > r) {
				hi = mid;
			} else {
				lo = mid + 1;
			}
		}
		dst[r] = __ldg(val + lo);
	}
}

} // namespace

cudaError_t launch_rle_decode(const uint8_t *blob, const RleTile *tiles, uint32_t n_tiles, void *out, uint32_t elem_bytes,
                              cudaStream_t stream) {
	if (n_tiles == 0) {
		return cudaSuccess;
	}
	if (elem_bytes == 8) {
		cubit_rle_decode_kernel<uint64_t><<<n_tiles, kDecodeThreads, 0, stream>>>(blob, tiles, static_cast<uint64_t *>(out));
	} else {
		cubit_rle_decode_kernel<uint32_t><<<n_tiles, kDecodeThreads, 0, stream>>>(blob, tiles, static_cast<uint32_t *>(out));
	}
	return cudaGetLastError();
}

cudaError_t launch_bp_decode(const uint8_t *blob, const BpGroup *groups, uint32_t n_groups, void *out,
                             uint32_t elem_bytes, cudaStream_t stream) {
	if (n_groups == 0) {
		return cudaSuccess;
	}
	if (elem_bytes == 8) {
		cubit_bp_decode_kernel<uint64_t><<<n_groups, kDecodeThreads, 0, stream>>>(blob, groups, static_cast<uint64_t *>(out));
	} else {
		cubit_bp_decode_kernel<uint32_t><<<n_groups, kDecodeThreads, 0, stream>>>(blob, groups, static_cast<uint32_t *>(out));
	}
	return cudaGetLastError();
}

} // namespace cubit
