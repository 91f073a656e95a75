/* cubit_gpu_wire.h — the NARROW WIRE FORMAT of the DataChunk hand-off, and its host-side unpacker.
 *
 * Row-returning queries are PCIe-bound: 8 bytes of row ID + 8 bytes per projected value cross the bus for every
 * selected row although, inside one DataChunk (≤ 2048 consecutive result rows, STANDARD_VECTOR_SIZE,
 * src/include/duckdb/common/vector_size.hpp:16), sorted row IDs — and most column values — span a small range.
 * cubit_gpu_fetch_wire_async therefore ships every (stream, chunk) as a frame of reference: one int64 base
 * (the chunk's minimum) + unsigned deltas of 0 / 1 / 2 / 4 / 8 bytes, the narrowest width that holds
 * (max − min).  The GPU picks the width per chunk and writes the frames straight into the caller's page-locked
 * buffer; the worker that fills a DataChunk widens ONE chunk at a time into the chunk's own vectors
 * (cubit_wire_unpack_chunk below), i.e. into cache, right before the next operator reads it — the int64 arrays
 * never exist in host DRAM.  Lossless for every 4- and 8-byte type (arithmetic is modulo 2^64 on the bit pattern).
 *
 * A wire holds one WINDOW of a result: rows [offset, offset + n) as C = ceil(n / 2048) chunks of S streams
 * (stream 0 = row IDs when asked for, then the projected columns in query order):
 *
 *   [ cubit_wire_header (64 B) ][ directory: S·C × cubit_wire_dir (16 B), entry s·C + c ][ pad to 256 B ]
 *   [ slots: S·C × 16 KiB, slot s·C + c holds dir.n deltas of dir.width bytes each, little endian ]
 *
 * Slots sit at fixed positions (so the GPU needs no pass to place them); only dir.n × dir.width bytes of each are
 * written and cross the bus.
 *
 * Mirrors: a DataChunk vector filled by a scan (src/function/table/table_scan.cpp:251-273) — same values, same
 * order; the frame-of-reference idea is the reference's own BitPacking FOR mode
 * (src/storage/compression/bitpacking.cpp:879, BitpackingFetchRow) applied to the hand-off instead of the disk. */
#ifndef CUBIT_GPU_WIRE_H
#define CUBIT_GPU_WIRE_H

#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUBIT_WIRE_CHUNK 2048u            /* rows per frame = STANDARD_VECTOR_SIZE                       */
#define CUBIT_WIRE_SLOT_BYTES 16384u      /* CUBIT_WIRE_CHUNK × 8                                        */
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

typedef struct cubit_wire_dir {
	int64_t base;   /* minimum of the chunk's values (as signed 64-bit)     */
	uint32_t width; /* bytes per delta: 0 (all values == base), 1, 2, 4, 8  */
	uint32_t n;     /* values in the chunk (2048 except in the last chunk)  */
} cubit_wire_dir;

/* bytes a wire buffer must have for a window of n_rows rows and n_streams streams */
static inline uint64_t cubit_wire_bytes(uint64_t n_rows, uint32_t n_streams) {
	const uint64_t c = (n_rows + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK;
	const uint64_t dir = sizeof(cubit_wire_header) + (uint64_t)n_streams * c * sizeof(cubit_wire_dir);
	return ((dir + 255) & ~255ull) + (uint64_t)n_streams * c * CUBIT_WIRE_SLOT_BYTES;
}

/* bytes of a wire that actually crossed the bus (directory + the written part of every slot) */
static inline uint64_t cubit_wire_payload_bytes(const void *wire) {
	const cubit_wire_header *h = (const cubit_wire_header *)wire;
	const cubit_wire_dir *d = (const cubit_wire_dir *)((const char *)wire + sizeof(cubit_wire_header));
	uint64_t b = 0;
	for (uint64_t i = 0; i < (uint64_t)h->n_streams * h->n_chunks; i++) {
		b += sizeof(cubit_wire_dir) + (((uint64_t)d[i].n * d[i].width + 15) & ~15ull);
	}
	return b;
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
	const uint64_t slot = (uint64_t)stream * h->n_chunks + chunk;
	const cubit_wire_dir d = ((const cubit_wire_dir *)((const char *)wire + sizeof(cubit_wire_header)))[slot];
	if (d.n > CUBIT_WIRE_CHUNK) {
		return -1;
	}
	const unsigned char *src = (const unsigned char *)wire + h->data_offset + slot * CUBIT_WIRE_SLOT_BYTES;
	const uint64_t base = (uint64_t)d.base;
	const uint32_t n = d.n;
	if (out_elem == 8) {
		uint64_t *o = (uint64_t *)out;
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
