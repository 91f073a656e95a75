/* rle_oracle.c — CPU restatement of DuckDB's RLE column-segment format — TEST INFRASTRUCTURE ONLY
 * (same status and rules as cubit_oracle.c: only tests/, smoke() and bench.py's CPU legs may use it).
 *
 * Reference: src/storage/compression/rle.cpp
 *   layout (RLECompressState::FlushSegment, :190-205; RLEScanState ctor, :248-258):
 *     [u64 rle_count_offset][T values[n_runs]] pad to 8 [u16 counts[n_runs]]
 *     rle_count_offset = AlignValue(8 + sizeof(T) * n_runs); n_runs itself is not stored — a scan consumes
 *     runs until the segment's row count is produced (RLEScanPartialInternal, :338-364)
 *   run building (RLEState::Update, :40-80): a run ends when the value changes or its length reaches
 *     65535 (rle_count_t = uint16_t, :13,71); NULLs extend the current run (:44-60) and live in the validity
 *     segment, not here.
 * Parity: pinned byte-for-byte (encoder) and value-for-value (decoder) to segments written by the reference
 * itself, tests/golden/rle_segments.npz (tests/golden/make_rle_golden.py). */
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* decode `count` rows of one RLE segment; 0 = ok, -1 = malformed */
ORACLE_API int oracle_rle_decode(const uint8_t *seg, uint64_t seg_bytes, uint32_t elem_bytes, uint64_t count, void *out,
                                 uint64_t *n_runs_out) {
	if (seg_bytes < 8 || (elem_bytes != 4 && elem_bytes != 8)) {
		return -1;
	}
	uint64_t off;
	memcpy(&off, seg, 8);
	if (off < 8 || (off & 7) || off > seg_bytes) {
		return -1;
	}
	const uint64_t max_runs = (off - 8) / elem_bytes;
	uint64_t produced = 0, run = 0;
	while (produced < count) {
		if (run >= max_runs || off + 2 * (run + 1) > seg_bytes) {
			return -1;
		}
		uint16_t c;
		memcpy(&c, seg + off + 2 * run, 2);
		if (c == 0) {
			return -1;
		}
		uint64_t take = c;
		if (take > count - produced) {
			take = count - produced;
		}
		if (elem_bytes == 8) {
			uint64_t v;
			memcpy(&v, seg + 8 + 8 * run, 8);
			for (uint64_t i = 0; i < take; i++) {
				((uint64_t *)out)[produced + i] = v;
			}
		} else {
			uint32_t v;
			memcpy(&v, seg + 8 + 4 * run, 4);
			for (uint64_t i = 0; i < take; i++) {
				((uint32_t *)out)[produced + i] = v;
			}
		}
		produced += take;
		run++;
	}
	if (n_runs_out) {
		*n_runs_out = run;
	}
	return 0;
}

/* encode n values as ONE RLE segment (no NULLs); returns the segment size or -1 when cap is too small */
ORACLE_API int64_t oracle_rle_encode(const void *values, uint64_t n, uint32_t elem_bytes, uint8_t *out, uint64_t cap) {
	/* pass 1: count runs */
	uint64_t n_runs = 0, len = 0;
	for (uint64_t i = 0; i < n; i++) {
		int same = i > 0 && memcmp((const uint8_t *)values + i * elem_bytes, (const uint8_t *)values + (i - 1) * elem_bytes,
		                           elem_bytes) == 0;
		if (i == 0 || !same || len == 65535) {
			n_runs++;
			len = 0;
		}
		len++;
	}
	const uint64_t off = (8 + (uint64_t)elem_bytes * n_runs + 7) & ~7ull;
	const uint64_t total = off + 2 * n_runs;
	if (total > cap) {
		return -1;
	}
	memset(out, 0, total);
	memcpy(out, &off, 8);
	uint64_t run = 0;
	len = 0;
	for (uint64_t i = 0; i < n; i++) {
		int same = i > 0 && memcmp((const uint8_t *)values + i * elem_bytes, (const uint8_t *)values + (i - 1) * elem_bytes,
		                           elem_bytes) == 0;
		if (i == 0 || !same || len == 65535) {
			if (i > 0) {
				uint16_t c = (uint16_t)len;
				memcpy(out + off + 2 * run, &c, 2);
				run++;
			}
			memcpy(out + 8 + (uint64_t)elem_bytes * run, (const uint8_t *)values + i * elem_bytes, elem_bytes);
			len = 0;
		}
		len++;
	}
	if (n > 0) {
		uint16_t c = (uint16_t)len;
		memcpy(out + off + 2 * run, &c, 2);
	}
	return (int64_t)total;
}
