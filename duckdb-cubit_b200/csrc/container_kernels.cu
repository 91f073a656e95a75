// container_kernels.cu — roaring-style containers of a COMPRESSED CUBIT index.
//
// The upstream CUBIT library keeps every value bitvector WAH-compressed (FastBit ibis::bitvector); a verbatim
// high-cardinality index does not fit HBM (the day-level l_shipdate index of SURVEY §8d is 2,526 × 75 MB = 190 GB
// at SF100).  A run-length stream like WAH has to be decoded front to back, which a GPU kernel that hands out
// fixed-size segments to CTAs cannot do — so HBM holds one CONTAINER per (value, segment), the Roaring layout with
// the CUBIT segment as the container:
//     EMPTY   no payload              (scan kernel: the stream contributes nothing to this segment)
//     FULL    no payload
//     ARRAY   ≤ kArrayMax sorted 16-bit row positions, padded to 16 bytes — bulk-copied next to the ring stage and
//             expanded in shared memory by the consumer warps
//     BITMAP  the verbatim segment — bulk-copied into the ring stage like an uncompressed bitvector
// and a directory entry per (value, segment) = type | set-bit count | pool offset (kernels.h).  WAH-compressed
// uploads are expanded once on the GPU (wah_decode.cu) and stored as containers, so the compressed words cross PCIe
// as they are and HBM never holds the verbatim form of more than a scratch batch.
//
// Kernels here are maintenance-side (build / upload / download / merge-back), HBM-bound, one CTA per (value,
// segment): compress (verbatim → container, pool space from an atomic bump cursor), expand, counts.
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {

namespace {

constexpr int kThreads = 256;

// verbatim segment → container.  grid = (n_seg, nv).
template <int WPTC> // 64-bit words per thread: seg_words / 256
__global__ void __launch_bounds__(kThreads) cubit_container_compress_kernel(const uint64_t *__restrict__ src,
                                                                            uint64_t words_per_bv, uint32_t seg_words,
                                                                            unsigned long long *__restrict__ dir,
                                                                            uint64_t dir_stride, uint8_t *__restrict__ pool,
                                                                            unsigned long long *__restrict__ cursor) {
	__shared__ uint32_t wtot[kThreads / 32];
	__shared__ unsigned long long s_off;
	const uint32_t seg = blockIdx.x, v = blockIdx.y;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// thread t owns WPTC CONSECUTIVE words, so that its set bits are consecutive in the sorted output
	const uint64_t *w = src + (uint64_t)v * words_per_bv + (uint64_t)seg * seg_words + threadIdx.x * WPTC;
	uint64_t x[WPTC];
	uint32_t mine = 0;
#pragma unroll
	for (int j = 0; j < WPTC; j++) {
		x[j] = __ldg(w + j);
		mine += __popcll(x[j]);
	}
	uint32_t incl = mine;
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
	uint32_t excl = incl - mine, total = 0;
#pragma unroll
	for (int k = 0; k < kThreads / 32; k++) {
		excl += k < warp ? wtot[k] : 0u;
		total += wtot[k];
	}
	const uint32_t seg_bits = seg_words * 64u;
	uint32_t type, bytes;
	if (total == 0) {
		type = CT_EMPTY;
		bytes = 0;
	} else if (total == seg_bits) {
		type = CT_FULL;
		bytes = 0;
	} else if (total <= (uint32_t)kArrayMax) {
		type = CT_ARRAY;
		bytes = (total * 2u + 15u) & ~15u;
	} else {
		type = CT_BITMAP;
		bytes = seg_words * 8u;
	}
	unsigned long long *de = dir + (uint64_t)v * dir_stride + seg;
	if (threadIdx.x == 0) {
		const unsigned long long old = *de;
		const uint32_t ot = ct_type(old);
		const unsigned long long old_bytes =
		    ot == CT_BITMAP ? seg_words * 8ull : (ot == CT_ARRAY ? ((ct_count(old) * 2ull + 15ull) & ~15ull) : 0ull);
		if (old_bytes) {
			atomicAdd(cursor + 1, old_bytes); // the replaced container is garbage from now on
		}
		const unsigned long long off = bytes ? atomicAdd(cursor, (unsigned long long)bytes) : 0ull;
		s_off = off;
		*de = ct_make(type, total, off);
	}
	__syncthreads();
	if (type == CT_BITMAP) {
		uint64_t *dst = reinterpret_cast<uint64_t *>(pool + s_off) + threadIdx.x * WPTC;
#pragma unroll
		for (int j = 0; j < WPTC; j++) {
			dst[j] = x[j];
		}
	} else if (type == CT_ARRAY) {
		uint16_t *dst = reinterpret_cast<uint16_t *>(pool + s_off);
		uint32_t p = excl;
#pragma unroll
		for (int j = 0; j < WPTC; j++) {
			const uint32_t bit0 = (threadIdx.x * WPTC + j) * 64u;
			for (uint64_t m = x[j]; m; m &= m - 1) {
				dst[p++] = (uint16_t)(bit0 + (uint32_t)__ffsll((long long)m) - 1u);
			}
		}
		if (threadIdx.x == 0) { // padding up to the 16-byte granule the bulk copy moves
			for (uint32_t q = total; q * 2u < bytes; q++) {
				dst[q] = 0;
			}
		}
	}
}

// container → verbatim segment.  grid = n_seg_alloc.
__global__ void __launch_bounds__(kThreads) cubit_container_expand_kernel(const unsigned long long *__restrict__ dir,
                                                                          const uint8_t *__restrict__ pool, uint32_t n_seg,
                                                                          uint32_t seg_words, uint64_t *__restrict__ dst) {
	__shared__ unsigned long long words[2048];
	const uint32_t seg = blockIdx.x;
	uint64_t *out = dst + (uint64_t)seg * seg_words;
	const unsigned long long d = seg < n_seg ? dir[seg] : 0ull;
	const uint32_t type = ct_type(d);
	if (type == CT_BITMAP) {
		const uint64_t *src = reinterpret_cast<const uint64_t *>(pool + ct_offset(d));
		for (uint32_t i = threadIdx.x; i < seg_words; i += kThreads) {
			out[i] = __ldg(src + i);
		}
		return;
	}
	if (type != CT_ARRAY) {
		const uint64_t fill = type == CT_FULL ? ~0ull : 0ull;
		for (uint32_t i = threadIdx.x; i < seg_words; i += kThreads) {
			out[i] = fill;
		}
		return;
	}
	for (uint32_t i = threadIdx.x; i < seg_words; i += kThreads) {
		words[i] = 0;
	}
	__syncthreads();
	const uint16_t *src = reinterpret_cast<const uint16_t *>(pool + ct_offset(d));
	const uint32_t cnt = ct_count(d);
	for (uint32_t i = threadIdx.x; i < cnt; i += kThreads) {
		const uint32_t p = __ldg(src + i);
		atomicOr(&words[p >> 6], 1ull << (p & 63u));
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < seg_words; i += kThreads) {
		out[i] = words[i];
	}
}

// popcount of every B_v = Σ of the counts the directory keeps.  grid = card.
__global__ void __launch_bounds__(kThreads) cubit_compressed_counts_kernel(const unsigned long long *__restrict__ dir,
                                                                           uint64_t dir_stride, uint32_t n_seg,
                                                                           unsigned long long *__restrict__ out) {
	__shared__ unsigned long long wsum[kThreads / 32];
	const unsigned long long *d = dir + (uint64_t)blockIdx.x * dir_stride;
	unsigned long long c = 0;
	for (uint32_t s = threadIdx.x; s < n_seg; s += kThreads) {
		c += ct_count(__ldg(d + s));
	}
#pragma unroll
	for (int k = 16; k > 0; k >>= 1) {
		c += __shfl_xor_sync(0xffffffffu, c, k);
	}
	if ((threadIdx.x & 31) == 0) {
		wsum[threadIdx.x >> 5] = c;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long t = 0;
		for (int w = 0; w < kThreads / 32; w++) {
			t += wsum[w];
		}
		out[blockIdx.x] = t;
	}
}

__global__ void cubit_add_limbs_kernel(const ResultHeader *__restrict__ hdr, long long *__restrict__ dst) {
	if (threadIdx.x == 0) {
		const unsigned long long lo = hdr->sum_lo;
		const long long hi = hdr->sum_hi;
		dst[0] += (long long)hdr->count;
		dst[1] += (long long)(lo & 0xffffffffull);
		dst[2] += (long long)(lo >> 32);
		dst[3] += (long long)((unsigned long long)hi & 0xffffffffull);
		dst[4] += hi >> 32; // the signed top limb
	}
}

} // namespace

cudaError_t launch_container_compress(const uint64_t *src, uint64_t words_per_bv, uint32_t nv, uint32_t n_seg,
                                      uint32_t seg_words, unsigned long long *dir, uint64_t dir_stride, uint8_t *pool,
                                      unsigned long long *cursor, cudaStream_t stream) {
	// grid.y is limited to 65535: batches of values
	for (uint32_t v0 = 0; v0 < nv; v0 += 32768) {
		const uint32_t n = (nv - v0) < 32768u ? (nv - v0) : 32768u;
		dim3 grid(n_seg, n);
		const uint64_t *s = src + (uint64_t)v0 * words_per_bv;
		unsigned long long *d = dir + (uint64_t)v0 * dir_stride;
		if (seg_words == 512) {
			cubit_container_compress_kernel<2><<<grid, kThreads, 0, stream>>>(s, words_per_bv, seg_words, d, dir_stride, pool, cursor);
		} else if (seg_words == 1024) {
			cubit_container_compress_kernel<4><<<grid, kThreads, 0, stream>>>(s, words_per_bv, seg_words, d, dir_stride, pool, cursor);
		} else {
			return cudaErrorInvalidValue;
		}
	}
	return cudaGetLastError();
}

cudaError_t launch_container_expand(const unsigned long long *dir, const uint8_t *pool, uint64_t n_seg_alloc, uint32_t n_seg,
                                    uint32_t seg_words, uint64_t *dst, cudaStream_t stream) {
	if (seg_words > 2048 || n_seg_alloc == 0) {
		return cudaErrorInvalidValue;
	}
	cubit_container_expand_kernel<<<(unsigned)n_seg_alloc, kThreads, 0, stream>>>(dir, pool, n_seg, seg_words, dst);
	return cudaGetLastError();
}

cudaError_t launch_compressed_counts(const unsigned long long *dir, uint64_t dir_stride, uint32_t n_seg, uint32_t card,
                                     const uint8_t *pool, uint32_t seg_words, unsigned long long *out, cudaStream_t stream) {
	(void)pool;
	(void)seg_words;
	cubit_compressed_counts_kernel<<<card, kThreads, 0, stream>>>(dir, dir_stride, n_seg, out);
	return cudaGetLastError();
}

cudaError_t launch_add_limbs(const ResultHeader *hdr, long long *dst, cudaStream_t stream) {
	cubit_add_limbs_kernel<<<1, 32, 0, stream>>>(hdr, dst);
	return cudaGetLastError();
}

} // namespace cubit
