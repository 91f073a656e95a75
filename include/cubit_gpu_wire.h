/* cubit_gpu_wire.h — the NARROW WIRE FORMAT of the DataChunk hand-off, and its host-side unpacker.
 *
 * Row-returning queries are PCIe-bound: 8 bytes of row ID + 8 bytes per projected value cross the bus for every
 * selected row although, inside one DataChunk (≤ 2048 consecutive result rows, STANDARD_VECTOR_SIZE,
 * src/include/duckdb/common/vector_size.hpp:16), sorted row IDs — and most column values — span a small range.
 * cubit_gpu_fetch_wire_async therefore ships every (stream, chunk) as a frame of reference: one int64 base
 * (the chunk's minimum) + unsigned deltas of 0 / 1 / 2 / 4 / 8 bytes, the narrowest width that holds
 * (max − min) — or, for the row-ID stream of a dense selection, as a BITMAP over [min, max] (the IDs are strictly
 * ascending, so bit j = "row min + j is selected"; at one row in two that is 2 bits per row ID instead of 16).
 * The GPU picks the form per chunk and writes the frames straight into the caller's page-locked
 * buffer; the worker that fills a DataChunk widens ONE chunk at a time into the chunk's own vectors
 * (cubit_wire_unpack_chunk below), i.e. into cache, right before the next operator reads it — the int64 arrays
 * never exist in host DRAM.  Lossless for every 4- and 8-byte type (arithmetic is modulo 2^64 on the bit pattern).
 *
 * A wire holds one WINDOW of a result: rows [offset, offset + n) as C = ceil(n / 2048) chunks of S streams
 * (stream 0 = row IDs when asked for, then the projected columns in query order):
 *
 *   [ cubit_wire_header (64 B) ][ directory: C·S × cubit_wire_dir (24 B), entry c·S + s ][ pad to 256 B ]
 *   [ frames, back to back in directory order (chunk-major: the order a consumer reads them), each padded to 16 B;
 *     frame c·S + s starts dir.offset bytes after header.data_offset ]
 *
 * Only the directory and the frames are written and cross the bus; the frames are compact and in consumption
 * order, so the host reads one sequential stream (hardware prefetch works; fixed 16 KiB slots cost a page walk and a
 * DRAM ramp per chunk: 3.4 ns per row and worker instead of 0.6).  The buffer must still be sized for the worst case
 * (every frame 8 bytes wide): cubit_wire_bytes.
 *
 * Mirrors: a DataChunk vector filled by a scan (src/function/table/table_scan.cpp:251-273) — same values, same
 * order; the frame-of-reference idea is the reference's own BitPacking FOR mode
 * (src/storage/compression/bitpacking.cpp:879, BitpackingFetchRow) applied to the hand-off instead of the disk. */
#ifndef CUBIT_GPU_WIRE_H
#define CUBIT_GPU_WIRE_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUBIT_WIRE_CHUNK 2048u            /* rows per frame = STANDARD_VECTOR_SIZE                       */
#define CUBIT_WIRE_SLOT_BYTES 16384u      /* CUBIT_WIRE_CHUNK × 8: the largest frame                     */
#define CUBIT_WIRE_MAGIC 0x45524957u      /* "WIRE"                                                      */

typedef struct cubit_wire_header {
	uint32_t magic;
	uint32_t n_streams;  /* S                                                   */
	uint64_t n_rows;     /* rows of the window                                  */
	uint64_t n_chunks;   /* C = ceil(n_rows / 2048)                             */
	uint64_t data_offset; /* byte offset of slot 0 from the start of the wire   */
	uint8_t elem[16];    /* bytes per value of every stream as the query returns them (8 for row IDs)    */
	uint64_t reserved[2];
} cubit_wire_header;

#define CUBIT_WIRE_BITMAP 255u            /* low byte of dir.width: the slot is a bitmap, dir.width >> 8 = its 64-bit words */

typedef struct cubit_wire_dir {
	int64_t base;   /* minimum of the chunk's values (as signed 64-bit)     */
	uint64_t offset; /* of the frame, in bytes from header.data_offset (a multiple of 16) */
	uint32_t width; /* bytes per delta: 0 (all values == base), 1, 2, 4, 8; or CUBIT_WIRE_BITMAP | words << 8:
	                   strictly ascending values, bit j of the slot's little-endian 64-bit words = base + j present */
	uint32_t n;     /* values in the chunk (2048 except in the last chunk)  */
} cubit_wire_dir;

static inline uint64_t cubit_wire_frame_bytes(cubit_wire_dir d) {
	return (d.width & 255u) == CUBIT_WIRE_BITMAP ? (uint64_t)(d.width >> 8) * 8 : (uint64_t)d.n * d.width;
}

static inline uint64_t cubit_wire_data_offset(uint64_t n_chunks, uint32_t n_streams) {
	return (sizeof(cubit_wire_header) + (uint64_t)n_streams * n_chunks * sizeof(cubit_wire_dir) + 255) & ~255ull;
}

/* bytes a wire buffer must have for a window of n_rows rows and n_streams streams (worst case: 8-byte frames) */
static inline uint64_t cubit_wire_bytes(uint64_t n_rows, uint32_t n_streams) {
	const uint64_t c = (n_rows + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK;
	return cubit_wire_data_offset(c, n_streams) + (uint64_t)n_streams * c * CUBIT_WIRE_SLOT_BYTES;
}

/* bytes of a wire that actually crossed the bus (directory + frames) */
static inline uint64_t cubit_wire_payload_bytes(const void *wire) {
	const cubit_wire_header *h = (const cubit_wire_header *)wire;
	const cubit_wire_dir *d = (const cubit_wire_dir *)((const char *)wire + sizeof(cubit_wire_header));
	uint64_t b = 0;
	for (uint64_t i = 0; i < (uint64_t)h->n_streams * h->n_chunks; i++) {
		b += sizeof(cubit_wire_dir) + ((cubit_wire_frame_bytes(d[i]) + 15) & ~15ull);
	}
	return b;
}

/* ---- widening loops.  With AVX2 (checked at run time) a worker widens ≈ 10 values per nanosecond, so 16 host
 * threads keep up with what PCIe delivers; the scalar loops are the fallback and the definition of the result. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(CUBIT_WIRE_NO_AVX2)
#include <immintrin.h>
#define CUBIT_WIRE_AVX2 1
__attribute__((target("avx2"))) static void cubit_wire_widen64_avx2(const unsigned char *src, uint32_t width, uint64_t base,
                                                                    uint64_t *o, uint32_t n) {
	const __m256i vb = _mm256_set1_epi64x((long long)base);
	uint32_t i = 0;
	if (width == 2) {
		const uint16_t *s = (const uint16_t *)src;
		for (; i + 8 <= n; i += 8) {
			const __m128i v = _mm_loadu_si128((const __m128i *)(s + i));
			_mm256_storeu_si256((__m256i *)(o + i), _mm256_add_epi64(vb, _mm256_cvtepu16_epi64(v)));
			_mm256_storeu_si256((__m256i *)(o + i + 4),
			                    _mm256_add_epi64(vb, _mm256_cvtepu16_epi64(_mm_srli_si128(v, 8))));
		}
		for (; i < n; i++) {
			o[i] = base + s[i];
		}
	} else if (width == 1) {
		for (; i + 8 <= n; i += 8) {
			const __m128i v = _mm_loadl_epi64((const __m128i *)(src + i));
			_mm256_storeu_si256((__m256i *)(o + i), _mm256_add_epi64(vb, _mm256_cvtepu8_epi64(v)));
			_mm256_storeu_si256((__m256i *)(o + i + 4),
			                    _mm256_add_epi64(vb, _mm256_cvtepu8_epi64(_mm_srli_si128(v, 4))));
		}
		for (; i < n; i++) {
			o[i] = base + src[i];
		}
	} else if (width == 4) {
		const uint32_t *s = (const uint32_t *)src;
		for (; i + 8 <= n; i += 8) {
			const __m256i v = _mm256_loadu_si256((const __m256i *)(s + i));
			_mm256_storeu_si256((__m256i *)(o + i),
			                    _mm256_add_epi64(vb, _mm256_cvtepu32_epi64(_mm256_castsi256_si128(v))));
			_mm256_storeu_si256((__m256i *)(o + i + 4),
			                    _mm256_add_epi64(vb, _mm256_cvtepu32_epi64(_mm256_extracti128_si256(v, 1))));
		}
		for (; i < n; i++) {
			o[i] = base + s[i];
		}
	} else if (width == 8) {
		const uint64_t *s = (const uint64_t *)src;
		for (; i + 4 <= n; i += 4) {
			_mm256_storeu_si256((__m256i *)(o + i),
			                    _mm256_add_epi64(vb, _mm256_loadu_si256((const __m256i *)(s + i))));
		}
		for (; i < n; i++) {
			o[i] = base + s[i];
		}
	} else {
		for (; i + 4 <= n; i += 4) {
			_mm256_storeu_si256((__m256i *)(o + i), vb);
		}
		for (; i < n; i++) {
			o[i] = base;
		}
	}
}
static inline int cubit_wire_has_avx2(void) {
	static int known = -1; /* benign race: every thread computes the same answer */
	if (known < 0) {
		known = (__builtin_cpu_supports("avx2") && !getenv("CUBIT_WIRE_SCALAR")) ? 1 : 0; /* (env: tests) */
	}
	return known;
}
#endif

/* bitmap frame → values: base + position of every set bit, ascending; stops after n values */
static inline uint32_t cubit_wire_bitmap_decode(const unsigned char *src, uint32_t words, uint64_t base, uint32_t n,
                                                void *out, uint32_t out_elem) {
	uint32_t k = 0;
	for (uint32_t w = 0; w < words && k < n; w++) {
		uint64_t bits;
		memcpy(&bits, src + (size_t)w * 8, 8);
		const uint64_t b0 = base + (uint64_t)w * 64;
		if (out_elem == 8) {
			uint64_t *o = (uint64_t *)out;
			while (bits && k < n) {
				o[k++] = b0 + (uint64_t)__builtin_ctzll(bits);
				bits &= bits - 1;
			}
		} else {
			uint32_t *o = (uint32_t *)out;
			while (bits && k < n) {
				o[k++] = (uint32_t)(b0 + (uint64_t)__builtin_ctzll(bits));
				bits &= bits - 1;
			}
		}
	}
	return k;
}

/* Widen chunk `chunk` of stream `stream` into `out` (out_elem = 8: int64 / double bit patterns, 4: int32 / float).
 * Returns the number of values written (0 for a chunk past the window), or -1 for a malformed wire. */
static inline int cubit_wire_unpack_chunk(const void *wire, uint32_t stream, uint64_t chunk, void *out,
                                          uint32_t out_elem) {
	const cubit_wire_header *h = (const cubit_wire_header *)wire;
	if (h->magic != CUBIT_WIRE_MAGIC || stream >= h->n_streams || (out_elem != 4 && out_elem != 8)) {
		return -1;
	}
	if (chunk >= h->n_chunks) {
		return 0;
	}
	const uint64_t slot = chunk * h->n_streams + stream;
	const cubit_wire_dir d = ((const cubit_wire_dir *)((const char *)wire + sizeof(cubit_wire_header)))[slot];
	if (d.n > CUBIT_WIRE_CHUNK || d.offset > h->n_chunks * h->n_streams * (uint64_t)CUBIT_WIRE_SLOT_BYTES) {
		return -1;
	}
	const unsigned char *src = (const unsigned char *)wire + h->data_offset + d.offset;
	const uint64_t base = (uint64_t)d.base;
	const uint32_t n = d.n;
	if ((d.width & 255u) == CUBIT_WIRE_BITMAP) {
		if ((uint64_t)(d.width >> 8) * 8 > CUBIT_WIRE_SLOT_BYTES) {
			return -1;
		}
		return cubit_wire_bitmap_decode(src, d.width >> 8, base, n, out, out_elem) == n ? (int)n : -1;
	}
	if (d.width > 8 || (d.width & (d.width - 1))) {
		return -1;
	}
	if (out_elem == 8) {
		uint64_t *o = (uint64_t *)out;
#ifdef CUBIT_WIRE_AVX2
		if (cubit_wire_has_avx2()) {
			cubit_wire_widen64_avx2(src, d.width, base, o, n);
			return (int)n;
		}
#endif
		switch (d.width) {
		case 0:
			for (uint32_t i = 0; i < n; i++) {
				o[i] = base;
			}
			break;
		case 1:
			for (uint32_t i = 0; i < n; i++) {
				o[i] = base + src[i];
			}
			break;
		case 2: {
			const uint16_t *s = (const uint16_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = base + s[i];
			}
			break;
		}
		case 4: {
			const uint32_t *s = (const uint32_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = base + s[i];
			}
			break;
		}
		case 8: {
			const uint64_t *s = (const uint64_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = base + s[i];
			}
			break;
		}
		default:
			return -1;
		}
	} else {
		uint32_t *o = (uint32_t *)out;
		const uint32_t b32 = (uint32_t)base;
		switch (d.width) {
		case 0:
			for (uint32_t i = 0; i < n; i++) {
				o[i] = b32;
			}
			break;
		case 1:
			for (uint32_t i = 0; i < n; i++) {
				o[i] = b32 + src[i];
			}
			break;
		case 2: {
			const uint16_t *s = (const uint16_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = b32 + s[i];
			}
			break;
		}
		case 4: {
			const uint32_t *s = (const uint32_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = b32 + s[i];
			}
			break;
		}
		case 8: {
			const uint64_t *s = (const uint64_t *)src;
			for (uint32_t i = 0; i < n; i++) {
				o[i] = (uint32_t)(base + s[i]);
			}
			break;
		}
		default:
			return -1;
		}
	}
	return (int)n;
}

#ifdef __cplusplus
}
#endif
#endif /* CUBIT_GPU_WIRE_H */
