// aux_kernels.cu — the kernels either side of the fused scan:
//   * cubit_probe_kernel      column probe at sorted row IDs (+ SUM / SUM(a*b))
//                             — the DataTable::Fetch analog
//                             (reference: src/storage/data_table.cpp:373-377 →
//                              row_group_collection.cpp:264-288 → FixedSizeFetchRow
//                              fixed_size_uncompressed.cpp:169-178)
//   * cubit_index_build_kernel  per-value bitvectors from a column scan
//                             (CREATE INDEX analog, plan_create_index.cpp:16-130)
//   * popcount / delta merge-back / synthetic column generators
#include "col_ref.cuh"
#include "kernels.h"

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace cubit {

__device__ __forceinline__ void add128(unsigned long long &lo, long long &hi, long long v) {
	unsigned long long uv = (unsigned long long)v;
	lo += uv;
	hi += (long long)(lo < uv) + (v >> 63);
}
__device__ __forceinline__ void add128(unsigned long long &lo, long long &hi, unsigned long long lo2, long long hi2) {
	lo += lo2;
	hi += hi2 + (long long)(lo < lo2);
}

// ------------------------------------------------------------------- probe
constexpr int kProbeThreads = 256;
constexpr int kProbePairs = 4; // row-ID pairs per thread per iteration (8 independent gathers in flight)

__device__ __forceinline__ bool row_valid(const unsigned long long *valid, int64_t local) {
	return !valid || ((__ldg(valid + (local >> 6)) >> (local & 63)) & 1ull);
}

__device__ __forceinline__ void agg_row(const ProbeArgs &a, int64_t local, unsigned long long &lo, long long &hi,
                                        unsigned int &ovf, double &fsum, unsigned long long &nn) {
	if (!row_valid(a.agg_valid_a, local) || (a.agg_kind == 2 && !row_valid(a.agg_valid_b, local))) {
		return; // NULL input: not aggregated
	}
	nn++;
	if (a.agg_kind == 3) {
		fsum += __longlong_as_double(load_col(a.agg_a, local));
	} else if (a.agg_kind == 1) {
		add128(lo, hi, load_col(a.agg_a, local));
	} else if (a.agg_kind == 2) {
		const long long x = load_col(a.agg_a, local);
		const long long y = load_col(a.agg_b, local);
		const long long pr = x * y;
		if (__mul64hi(x, y) != (pr >> 63)) {
			ovf = 1;
		}
		add128(lo, hi, pr);
	}
}

__global__ void __launch_bounds__(kProbeThreads) cubit_probe_kernel(const __grid_constant__ ProbeArgs a) {
	const unsigned long long n = a.count_ptr ? *a.count_ptr : a.n;
	const unsigned long long n_pairs = n >> 1;
	unsigned long long lo = 0;
	long long hi = 0;
	unsigned int ovf = 0;
	double fsum = 0.0;
	unsigned long long nn = 0; // non-NULL aggregate inputs

	const unsigned long long stride = (unsigned long long)gridDim.x * kProbeThreads * kProbePairs;
	for (unsigned long long base = (unsigned long long)blockIdx.x * kProbeThreads * kProbePairs; base < n_pairs;
	     base += stride) {
		longlong2 id[kProbePairs];
		bool ok[kProbePairs];
#pragma unroll
		for (int u = 0; u < kProbePairs; u++) {
			const unsigned long long p = base + (unsigned long long)u * kProbeThreads + threadIdx.x;
			ok[u] = p < n_pairs;
			id[u] = ok[u] ? __ldg(reinterpret_cast<const longlong2 *>(a.ids) + p) : make_longlong2(0, 0);
		}
		for (int c = 0; c < a.n_cols; c++) {
			if (a.elem_bytes[c] == 8) {
				const long long *col = static_cast<const long long *>(a.col[c]);
				longlong2 v[kProbePairs];
#pragma unroll
				for (int u = 0; u < kProbePairs; u++) {
					if (ok[u]) {
						const int64_t l0 = id[u].x - a.row_base, l1 = id[u].y - a.row_base;
						if (!col) { // bit-packed column
							v[u].x = load_col(a.packed[c], l0);
							v[u].y = load_col(a.packed[c], l1);
						} else if (l1 == l0 + 1 && (l0 & 1) == 0) { // dense aligned run: one 128-bit load
							v[u] = __ldg(reinterpret_cast<const longlong2 *>(col + l0));
						} else {
							v[u].x = __ldg(col + l0);
							v[u].y = __ldg(col + l1);
						}
					}
				}
#pragma unroll
				for (int u = 0; u < kProbePairs; u++) {
					if (ok[u]) {
						const unsigned long long p = base + (unsigned long long)u * kProbeThreads + threadIdx.x;
						__stcs(reinterpret_cast<longlong2 *>(a.out[c]) + p, v[u]);
					}
				}
			} else {
				const int *col = static_cast<const int *>(a.col[c]);
				int2 v[kProbePairs];
#pragma unroll
				for (int u = 0; u < kProbePairs; u++) {
					if (ok[u]) {
						v[u].x = __ldg(col + (id[u].x - a.row_base));
						v[u].y = __ldg(col + (id[u].y - a.row_base));
					}
				}
#pragma unroll
				for (int u = 0; u < kProbePairs; u++) {
					if (ok[u]) {
						const unsigned long long p = base + (unsigned long long)u * kProbeThreads + threadIdx.x;
						__stcs(reinterpret_cast<int2 *>(a.out[c]) + p, v[u]);
					}
				}
			}
		}
		if (a.agg_kind) {
#pragma unroll
			for (int u = 0; u < kProbePairs; u++) {
				if (ok[u]) {
					agg_row(a, id[u].x - a.row_base, lo, hi, ovf, fsum, nn);
					agg_row(a, id[u].y - a.row_base, lo, hi, ovf, fsum, nn);
				}
			}
		}
	}
	// odd tail element
	if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
		const int64_t l = a.ids[n - 1] - a.row_base;
		for (int c = 0; c < a.n_cols; c++) {
			if (a.elem_bytes[c] == 8) {
				static_cast<long long *>(a.out[c])[n - 1] =
				    a.col[c] ? static_cast<const long long *>(a.col[c])[l] : load_col(a.packed[c], l);
			} else {
				static_cast<int *>(a.out[c])[n - 1] = static_cast<const int *>(a.col[c])[l];
			}
		}
		if (a.agg_kind) {
			agg_row(a, l, lo, hi, ovf, fsum, nn);
		}
	}
	if (!a.agg_kind) {
		return;
	}
	__shared__ BlockPartial red[kProbeThreads / 32];
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const unsigned long long olo = __shfl_xor_sync(0xffffffffu, lo, d);
		const long long ohi = __shfl_xor_sync(0xffffffffu, hi, d);
		add128(lo, hi, olo, ohi);
		ovf |= __shfl_xor_sync(0xffffffffu, ovf, d);
		fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
	}
	if ((threadIdx.x & 31) == 0 && fsum != 0.0) {
		atomicAdd(&a.hdr->sum_f64, fsum);
	}
	if (a.agg_valid_a || a.agg_valid_b) {
		nn = __reduce_add_sync(0xffffffffu, (unsigned)nn); // per-warp total stays far below 2^32 (rows / resident threads)
		if ((threadIdx.x & 31) == 0 && nn) {
			atomicAdd(&a.hdr->agg_rows, nn);
		}
	}
	if ((threadIdx.x & 31) == 0) {
		red[threadIdx.x >> 5].sum_lo = lo;
		red[threadIdx.x >> 5].sum_hi = hi;
		red[threadIdx.x >> 5].pad = ovf;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long tlo = 0, tov = 0;
		long long thi = 0;
		for (int w = 0; w < kProbeThreads / 32; w++) {
			add128(tlo, thi, red[w].sum_lo, red[w].sum_hi);
			tov |= red[w].pad;
		}
		if (tlo | (unsigned long long)thi) { // exact 128-bit accumulate (carry from the returned old value)
			const unsigned long long old = atomicAdd(&a.hdr->sum_lo, tlo);
			const long long carry = (old + tlo) < old ? 1 : 0;
			atomicAdd(reinterpret_cast<unsigned long long *>(&a.hdr->sum_hi), (unsigned long long)(thi + carry));
		}
		if (tov) {
			atomicOr(&a.hdr->overflow, 1u);
		}
	}
}

cudaError_t launch_probe(const ProbeArgs &args, int sm_count, cudaStream_t stream) {
	const int grid = sm_count * 8;
	cubit_probe_kernel<<<grid, kProbeThreads, 0, stream>>>(args);
	return cudaGetLastError();
}
int probe_grid(int sm_count) {
	return sm_count * 8;
}

// ------------------------------------------------------- validity at row IDs
__global__ void __launch_bounds__(256) cubit_validity_gather_kernel(const long long *__restrict__ ids,
                                                                    const unsigned long long *__restrict__ count_ptr,
                                                                    int64_t row_base,
                                                                    const unsigned long long *__restrict__ valid,
                                                                    uint32_t *__restrict__ out32) {
	const unsigned long long n = *count_ptr;
	const unsigned long long n32 = (n + 31) & ~31ull;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
		bool bit = false;
		if (i < n) {
			const int64_t local = __ldg(ids + i) - row_base;
			bit = (__ldg(valid + (local >> 6)) >> (local & 63)) & 1ull;
		}
		const unsigned m = __ballot_sync(0xffffffffu, bit);
		if ((threadIdx.x & 31) == 0) {
			out32[i >> 5] = m;
		}
	}
}

cudaError_t launch_validity_gather(const long long *ids, const unsigned long long *count_ptr, int64_t row_base,
                                   const unsigned long long *valid, uint32_t *out32, int sm_count,
                                   cudaStream_t stream) {
	cubit_validity_gather_kernel<<<sm_count * 8, 256, 0, stream>>>(ids, count_ptr, row_base, valid, out32);
	return cudaGetLastError();
}

// ------------------------------------------------------------- index build
// One CTA converts kBuildRows consecutive rows per iteration.  Each warp reads
// 32 consecutive rows (coalesced) and every lane ORs its row's bit into the
// 32-bit slice of its value's bitvector in a shared-memory tile
// [value][kBuildRows/32] (ATOMS.OR, bank-swizzled by the value), which is then
// flushed as contiguous runs of kBuildRows/8 bytes per value.
constexpr int kBuildThreads = 256;
constexpr int kBuildRows = 4096;
constexpr int kBuildSlots = kBuildRows / 32; // u32 slots per value per tile

template <typename T>
__global__ void __launch_bounds__(kBuildThreads)
    cubit_index_build_kernel(const T *__restrict__ col, const unsigned long long *__restrict__ valid, uint64_t row_begin,
                             uint64_t n_rows, int64_t base_value, uint32_t v_lo, uint32_t v_n,
                             uint64_t *__restrict__ bitvectors, uint64_t words_per_bv) {
	extern __shared__ uint32_t tile[]; // [v_n][kBuildSlots]
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint64_t n_tiles = (n_rows + kBuildRows - 1) / kBuildRows;
	constexpr int kIters = kBuildSlots / (kBuildThreads / 32); // 32-row slots per warp per tile
	const long long off = base_value + (long long)v_lo;
	// this thread's kIters column values of tile t as value ids relative to v_lo (0xffffffff: not indexed here)
	auto load_tile = [&](uint64_t t, uint32_t (&rel)[kIters]) {
		const uint64_t row0 = t * kBuildRows;
		if (row0 >= row_begin && row0 + kBuildRows <= n_rows && !valid) { // interior tile, no NULLs: no per-row checks
			const T *src = col + row0 + (uint64_t)warp * 32 + lane;
			T raw[kIters]; // all loads first, conversions after: kIters independent 128-byte requests per warp in flight
#pragma unroll
			for (int it = 0; it < kIters; it++) {
				raw[it] = __ldcs(src + it * kBuildThreads);
			}
#pragma unroll
			for (int it = 0; it < kIters; it++) {
				const unsigned long long d = (unsigned long long)((long long)raw[it] - off);
				rel[it] = d < (unsigned long long)v_n ? (uint32_t)d : 0xffffffffu;
			}
			return;
		}
#pragma unroll
		for (int it = 0; it < kIters; it++) {
			const uint64_t r = row0 + (uint64_t)(it * (kBuildThreads / 32) + warp) * 32 + lane;
			// rows below row_begin are already indexed (append path): their bits are kept as they are
			unsigned long long d = ~0ull;
			if (r >= row_begin && r < n_rows) {
				d = (unsigned long long)((long long)__ldcs(col + r) - off);
				if (valid) { // NULL keys are not indexed: one 32-bit slice of the validity mask per warp
					const uint32_t vm = __ldg(reinterpret_cast<const uint32_t *>(valid) + (r >> 5)); // r >> 5 is warp-uniform
					if (!((vm >> (r & 31u)) & 1u)) {
						d = ~0ull;
					}
				}
			}
			rel[it] = d < (unsigned long long)v_n ? (uint32_t)d : 0xffffffffu;
		}
	};
	uint32_t cur[kIters], nxt[kIters];
	uint64_t t = row_begin / kBuildRows + blockIdx.x;
	if (t < n_tiles) {
		load_tile(t, cur);
	}
	// the tile is zeroed once; the flush of every tile leaves it zeroed for the next one
	for (uint32_t i = threadIdx.x; i < v_n * (kBuildSlots / 4); i += kBuildThreads) {
		reinterpret_cast<uint4 *>(tile)[i] = make_uint4(0, 0, 0, 0);
	}
	__syncthreads();
	for (; t < n_tiles; t += gridDim.x) {
		const uint64_t row0 = t * kBuildRows;
		// the NEXT tile's column values are in flight while this one is scattered and flushed (the loads of a tile
		// right in front of its scatter were 56 % of the stall samples)
		if (t + gridDim.x < n_tiles) {
			load_tile(t + gridDim.x, nxt);
		}
		// every row sets ITS bit with one shared-memory atomic OR (fire and forget: nothing waits for it), instead of a
		// __match_any_sync per 32 rows whose result the next instruction needs — that dependency was 40 % of the
		// kernel's stall samples (profiles/r2_index_build.md).  The slot index is warp-uniform, so the tile is swizzled
		// by the value (the low five bits of the slot XOR the value's): lanes with different values hit different
		// banks; the flush undoes it (piece index XOR, then a permutation of the piece's four words).
#pragma unroll
		for (int it = 0; it < kIters; it++) {
			const uint32_t slot = (uint32_t)(it * (kBuildThreads / 32) + warp);
			const uint32_t v = cur[it];
			if (v != 0xffffffffu) {
				atomicOr(&tile[v * kBuildSlots + (slot ^ (v & 31u))], 1u << lane);
			}
		}
		__syncthreads();
		// flush (and re-zero): per value kBuildSlots u32 = kBuildRows/64 u64 words = 512 contiguous bytes of B_v — one
		// coalesced 512-byte store per (warp, value): lane l moves the 16-byte piece l
		const uint64_t word0 = row0 / 64;
		static_assert(kBuildRows / 128 == 32, "one 16-byte piece per lane");
		const bool interior = row0 >= row_begin && word0 + kBuildRows / 64 <= words_per_bv;
		for (uint32_t v = (uint32_t)warp; v < v_n; v += kBuildThreads / 32) {
			// un-swizzle: logical piece l sits at physical piece l ^ (m >> 2), its four words permuted by m & 3 (m = v mod 32)
			uint4 *src = reinterpret_cast<uint4 *>(tile) + v * (kBuildSlots / 4) + ((uint32_t)lane ^ ((v & 31u) >> 2));
			uint4 x = *src;
			*src = make_uint4(0, 0, 0, 0);
			if (v & 1u) {
				uint32_t tmp = x.x;
				x.x = x.y;
				x.y = tmp;
				tmp = x.z;
				x.z = x.w;
				x.w = tmp;
			}
			if (v & 2u) {
				uint32_t tmp = x.x;
				x.x = x.z;
				x.z = tmp;
				tmp = x.y;
				x.y = x.w;
				x.w = tmp;
			}
			uint4 *dst = reinterpret_cast<uint4 *>(bitvectors + (uint64_t)(v_lo + v) * words_per_bv + word0 + 2 * lane);
			if (!interior) {
				const uint64_t piece_row0 = (word0 + 2 * lane) * 64;
				if (word0 + 2 * lane >= words_per_bv || piece_row0 + 128 <= row_begin) {
					continue; // past the capacity / only rows indexed earlier
				}
				if (piece_row0 < row_begin) { // the boundary piece: old rows' bits stay (the tile holds zeros for them)
					const uint4 old = *dst;
					x.x |= old.x;
					x.y |= old.y;
					x.z |= old.z;
					x.w |= old.w;
				}
			}
			__stcs(dst, x);
		}
		__syncthreads();
#pragma unroll
		for (int it = 0; it < kIters; it++) {
			cur[it] = nxt[it];
		}
	}
}

cudaError_t launch_index_build(const void *col, uint32_t elem_bytes, const unsigned long long *valid, uint64_t row_begin,
                               uint64_t n_rows, int64_t base_value,
                               uint32_t cardinality, uint64_t *bitvectors, uint64_t words_per_bv, int sm_count,
                               cudaStream_t stream, int *n_launches) {
	// values per pass bounded by shared memory (≤ 200 KiB tile)
	const uint32_t max_v = (200u * 1024u) / (kBuildSlots * 4u);
	int launches = 0;
	static bool configured[64] = {}; // function attributes are per device
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (!configured[dev]) {
		cudaFuncSetAttribute(cubit_index_build_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		cudaFuncSetAttribute(cubit_index_build_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                     200 * 1024);
		configured[dev] = true;
	}
	const uint64_t n_tiles = (n_rows + kBuildRows - 1) / kBuildRows - row_begin / kBuildRows;
	for (uint32_t v_lo = 0; v_lo < cardinality; v_lo += max_v) {
		const uint32_t v_n = (cardinality - v_lo) < max_v ? (cardinality - v_lo) : max_v;
		const size_t smem = (size_t)v_n * kBuildSlots * 4;
		int per_sm = (int)((220u * 1024u) / (smem + 1024));
		if (per_sm < 1) {
			per_sm = 1;
		}
		if (per_sm > 8) {
			per_sm = 8;
		}
		uint64_t grid = (uint64_t)sm_count * per_sm;
		if (grid > n_tiles) {
			grid = n_tiles;
		}
		if (grid < 1) {
			grid = 1;
		}
		if (elem_bytes == 4) {
			cubit_index_build_kernel<int><<<(unsigned)grid, kBuildThreads, smem, stream>>>(
			    static_cast<const int *>(col), valid, row_begin, n_rows, base_value, v_lo, v_n, bitvectors, words_per_bv);
		} else {
			cubit_index_build_kernel<long long><<<(unsigned)grid, kBuildThreads, smem, stream>>>(
			    static_cast<const long long *>(col), valid, row_begin, n_rows, base_value, v_lo, v_n, bitvectors,
			    words_per_bv);
		}
		launches++;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) {
			return e;
		}
	}
	if (n_launches) {
		*n_launches = launches;
	}
	return cudaSuccess;
}

// ---------------------------------------------------------------- popcount
__global__ void cubit_popcount_many_kernel(const uint64_t *__restrict__ bitvectors, uint64_t words_per_bv,
                                           unsigned long long *__restrict__ out) {
	// grid = (blocks_per_bv, n_bv)
	const uint64_t *bv = bitvectors + (uint64_t)blockIdx.y * words_per_bv;
	unsigned long long c = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_bv;
	     i += (uint64_t)gridDim.x * blockDim.x) {
		c += __popcll(bv[i]);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		c += __shfl_xor_sync(0xffffffffu, c, d);
	}
	if ((threadIdx.x & 31) == 0 && c) {
		atomicAdd(out + blockIdx.y, c);
	}
}

cudaError_t launch_popcount_many(const uint64_t *bitvectors, uint64_t words_per_bv, uint32_t n_bv,
                                 unsigned long long *out, cudaStream_t stream) {
	cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned long long) * n_bv, stream);
	if (e != cudaSuccess) {
		return e;
	}
	uint64_t bx = (words_per_bv + 256 * 8 - 1) / (256 * 8);
	if (bx > 1024) {
		bx = 1024;
	}
	if (bx < 1) {
		bx = 1;
	}
	for (uint32_t v0 = 0; v0 < n_bv; v0 += 32768) {
		const uint32_t nv = (n_bv - v0) < 32768u ? (n_bv - v0) : 32768u;
		dim3 grid((unsigned)bx, nv);
		cubit_popcount_many_kernel<<<grid, 256, 0, stream>>>(bitvectors + (uint64_t)v0 * words_per_bv, words_per_bv,
		                                                      out + v0);
	}
	return cudaGetLastError();
}

cudaError_t launch_popcount(const uint64_t *words, uint64_t n_words, unsigned long long *out, int sm_count,
                            cudaStream_t stream) {
	(void)sm_count;
	return launch_popcount_many(words, n_words, 1, out, stream);
}

// ------------------------------------------------------- FOR bit-packing
// pass 1: per block of kPackBlock rows, min and the bit width of (max - min)
__global__ void __launch_bounds__(256) cubit_pack_widths_kernel(const long long *__restrict__ col, uint64_t n_rows,
                                                                long long *__restrict__ base_out,
                                                                uint32_t *__restrict__ width_out) {
	__shared__ long long smin[8], smax[8];
	const uint64_t blk = blockIdx.x;
	long long mn = LLONG_MAX, mx = LLONG_MIN;
	for (int i = threadIdx.x; i < kPackBlock; i += 256) {
		const uint64_t r = blk * kPackBlock + i;
		if (r < n_rows) {
			const long long v = col[r];
			mn = v < mn ? v : mn;
			mx = v > mx ? v : mx;
		}
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
		mn = a < mn ? a : mn;
		mx = b > mx ? b : mx;
	}
	if ((threadIdx.x & 31) == 0) {
		smin[threadIdx.x >> 5] = mn;
		smax[threadIdx.x >> 5] = mx;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++) {
			mn = smin[w] < mn ? smin[w] : mn;
			mx = smax[w] > mx ? smax[w] : mx;
		}
		const unsigned long long range = (unsigned long long)mx - (unsigned long long)mn; // exact in 64 bits
		base_out[blk] = mn;
		width_out[blk] = range ? 64u - (uint32_t)__clzll((long long)range) : 0u;
	}
}

// pass 2: one CTA packs one block; thread w assembles payload word w from the staged (value - base)
__global__ void __launch_bounds__(256) cubit_pack_blocks_kernel(const long long *__restrict__ col, uint64_t n_rows,
                                                                const PackHdr *__restrict__ hdr,
                                                                unsigned long long *__restrict__ words) {
	__shared__ unsigned long long rel[kPackBlock];
	const uint64_t blk = blockIdx.x;
	const PackHdr h = hdr[blk];
	for (int i = threadIdx.x; i < kPackBlock; i += 256) {
		const uint64_t r = blk * kPackBlock + i;
		rel[i] = r < n_rows ? (unsigned long long)col[r] - (unsigned long long)h.base : 0ull;
	}
	__syncthreads();
	const uint32_t width = h.width, n_words = 16u * width; // 1024 * width / 64
	for (uint32_t w = threadIdx.x; w < n_words; w += 256) {
		unsigned long long out = 0;
		const uint32_t bit0 = w * 64u;
		uint32_t i = bit0 / width; // first value overlapping this word
		while (i < (uint32_t)kPackBlock && i * width < bit0 + 64u) {
			const uint32_t vb = i * width;
			const unsigned long long v = rel[i];
			out |= vb >= bit0 ? v << (vb - bit0) : v >> (bit0 - vb);
			i++;
		}
		words[h.word_off + w] = out;
	}
}

cudaError_t launch_pack_widths(const long long *col, uint64_t n_rows, long long *base_out, uint32_t *width_out,
                               cudaStream_t stream) {
	const uint64_t blocks = (n_rows + kPackBlock - 1) / kPackBlock;
	cubit_pack_widths_kernel<<<(unsigned)blocks, 256, 0, stream>>>(col, n_rows, base_out, width_out);
	return cudaGetLastError();
}

cudaError_t launch_pack_blocks(const long long *col, uint64_t n_rows, const PackHdr *hdr, unsigned long long *words,
                               cudaStream_t stream) {
	const uint64_t blocks = (n_rows + kPackBlock - 1) / kPackBlock;
	cubit_pack_blocks_kernel<<<(unsigned)blocks, 256, 0, stream>>>(col, n_rows, hdr, words);
	return cudaGetLastError();
}

// ------------------------------------------------------ synthetic columns
// restated in oracle/cubit_oracle.c (oracle_synth_value)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

__global__ void cubit_synth_kernel(void *col, int kind, uint64_t n_rows, int64_t row_base, uint64_t seed,
                                   uint64_t threshold, uint32_t card, uint32_t hot_lo, uint32_t hot_n) {
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
	     r += (uint64_t)gridDim.x * blockDim.x) {
		if (kind == 0) {
			static_cast<long long *>(col)[r] = row_base + (long long)r;
		} else if (kind == 2) { // uniform int32 in [hot_lo, hot_lo + card)
			const uint64_t z = splitmix64(seed + (uint64_t)row_base + r);
			static_cast<int *>(col)[r] = (int)(hot_lo + (uint32_t)((z >> 7) % card));
		} else if (kind == 3) { // uniform int64 in [hot_lo, hot_lo + threshold)
			const uint64_t z = splitmix64(seed + (uint64_t)row_base + r);
			static_cast<long long *>(col)[r] = (long long)hot_lo + (long long)((z >> 7) % threshold);
		} else {
			const uint64_t z = splitmix64(seed + (uint64_t)row_base + r);
			const uint64_t y = z >> 7;
			uint32_t v;
			if (z < threshold) {
				v = hot_lo + (uint32_t)(y % hot_n);
			} else {
				v = (uint32_t)(y % (card - hot_n));
				if (v >= hot_lo) {
					v += hot_n;
				}
			}
			static_cast<int *>(col)[r] = (int)v;
		}
	}
}

cudaError_t launch_synth_column(void *col, int kind, uint64_t n_rows, int64_t row_base, uint64_t seed,
                                uint64_t threshold, uint32_t card, uint32_t hot_lo, uint32_t hot_n, int sm_count,
                                cudaStream_t stream) {
	cubit_synth_kernel<<<sm_count * 16, 256, 0, stream>>>(col, kind, n_rows, row_base, seed, threshold, card, hot_lo,
	                                                      hot_n);
	return cudaGetLastError();
}

} // namespace cubit
