/*
 * bitpacking_oracle.c — CPU restatement of the reference's BitPacking column-segment decoder.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT (same rule as cubit_oracle.c): only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * PARITY STATUS: PINNED.  Unlike the bitmap arithmetic, this code path EXISTS in /root/reference:
 *   segment layout     src/storage/compression/bitpacking.cpp:22-75 (metadata word), :524-544 (FlushSegment:
 *                      data grows up from byte 8, one u32 metadata word per 2048-value group grows down,
 *                      bytes 0..7 = offset of the END of the metadata = segment size)
 *   group headers      BitpackingScanState::LoadNextGroup, bitpacking.cpp:660-709
 *   decode             BitpackingScanPartial, bitpacking.cpp:776-860 (CONSTANT, CONSTANT_DELTA, FOR, DELTA_FOR;
 *                      sign extension is always skipped, :786)
 *   delta decode       DeltaDecode, bitpacking.cpp:596-620;  ApplyFrameOfReference, :583-593
 *   32-value unpack    duckdb_fastpforlib::fastunpack(const uint32_t *in, uint32_t or uint64_t *out, bit),
 *                      third_party/fastpforlib/bitpacking.cpp:173-226 (Unroller): value i of a group is the
 *                      `bit` bits at bit position i*bit of the group's little-endian 32-bit word stream
 * It is checked (tests/test_bitpacking.py) against column segments the reference binary itself wrote
 * (tests/golden/bitpacking_segments.npz, lifted from a checkpointed database file by
 * tests/golden/make_bitpacking_golden.py) and against the reference's own fastunpack/fastpack compiled from
 * /root/reference/third_party/fastpforlib into oracle/_ref/ (oracle/Makefile target `ref`).
 */
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { BP_INVALID = 0, BP_AUTO = 1, BP_CONSTANT = 2, BP_CONSTANT_DELTA = 3, BP_DELTA_FOR = 4, BP_FOR = 5 };
#define BP_META_GROUP 2048u /* BITPACKING_METADATA_GROUP_SIZE, bitpacking.cpp:22 */
#define BP_ALGO_GROUP 32u   /* BITPACKING_ALGORITHM_GROUP_SIZE, bitpacking.hpp:30 */

static inline uint32_t ld32(const uint8_t *p) {
	uint32_t v;
	memcpy(&v, p, 4);
	return v;
}
static inline uint64_t ld64(const uint8_t *p) {
	uint64_t v;
	memcpy(&v, p, 8);
	return v;
}

/* fastunpack(const uint32_t *in, uint64_t *out, bit): 32 values of `bit` bits, LSB first, from 32-bit words */
ORACLE_API void oracle_bp_unpack_group64(const uint8_t *in, uint64_t *out, uint32_t bit) {
	for (uint32_t i = 0; i < BP_ALGO_GROUP; i++) {
		uint64_t pos = (uint64_t)i * bit;
		uint64_t v = 0;
		uint32_t got = 0;
		while (got < bit) { /* walk the 32-bit words this value touches (up to three) */
			const uint32_t w = ld32(in + 4 * ((pos + got) >> 5));
			const uint32_t sh = (uint32_t)((pos + got) & 31u);
			const uint32_t take = (32u - sh) < (bit - got) ? (32u - sh) : (bit - got);
			const uint64_t piece = ((uint64_t)w >> sh) & (take == 32 ? 0xffffffffull : ((1ull << take) - 1));
			v |= piece << got;
			got += take;
		}
		out[i] = v;
	}
}

/* fastunpack(const uint32_t *in, uint32_t *out, bit) */
ORACLE_API void oracle_bp_unpack_group32(const uint8_t *in, uint32_t *out, uint32_t bit) {
	for (uint32_t i = 0; i < BP_ALGO_GROUP; i++) {
		uint64_t pos = (uint64_t)i * bit;
		uint64_t two = ld32(in + 4 * (pos >> 5));
		if (((pos & 31u) + bit) > 32u) {
			two |= (uint64_t)ld32(in + 4 * (pos >> 5) + 4) << 32;
		}
		out[i] = bit == 0 ? 0u : (uint32_t)((two >> (pos & 31u)) & (bit >= 32 ? 0xffffffffull : ((1ull << bit) - 1)));
	}
}

/* Decode `count` values of one BitPacking segment.
 *   seg, seg_bytes : the segment as stored in the block (starts with the u64 metadata-end offset)
 *   elem_bytes     : 8 (BIGINT) or 4 (INTEGER / DATE)
 *   out            : count elements of elem_bytes
 *   mode_hist[6]   : optional, += number of metadata groups per BitpackingMode
 * returns 0, or a negative code for a malformed segment */
ORACLE_API int oracle_bitpacking_decode(const uint8_t *seg, uint64_t seg_bytes, uint32_t elem_bytes, uint64_t count,
                                        void *out, uint64_t *mode_hist) {
	if (seg_bytes < 12 || (elem_bytes != 8 && elem_bytes != 4)) {
		return -1;
	}
	const uint64_t meta_end = ld64(seg); /* BitpackingScanState ctor, bitpacking.cpp:633-636 */
	if (meta_end > seg_bytes || meta_end < 12) {
		return -2;
	}
	const uint8_t *meta = seg + meta_end - 4; /* first group's metadata word; following groups at lower addresses */
	uint64_t done = 0;
	while (done < count) {
		if (meta < seg + 8) {
			return -3;
		}
		const uint32_t enc = ld32(meta); /* DecodeMeta, bitpacking.cpp:68-73 */
		meta -= 4;
		const uint32_t mode = enc >> 24, off = enc & 0x00ffffffu;
		const uint64_t n = (count - done) < BP_META_GROUP ? (count - done) : BP_META_GROUP;
		if (mode_hist && mode < 6) {
			mode_hist[mode]++;
		}
		const uint8_t *p = seg + off;
		if (elem_bytes == 8) {
			uint64_t *o = (uint64_t *)out + done;
			if (mode == BP_CONSTANT) {
				const uint64_t c = ld64(p);
				for (uint64_t i = 0; i < n; i++) {
					o[i] = c;
				}
			} else if (mode == BP_CONSTANT_DELTA) {
				const uint64_t frame = ld64(p), delta = ld64(p + 8); /* bitpacking.cpp:815-826 */
				for (uint64_t i = 0; i < n; i++) {
					o[i] = delta * i + frame;
				}
			} else if (mode == BP_FOR || mode == BP_DELTA_FOR) {
				const uint64_t frame = ld64(p);
				const uint32_t width = (uint32_t)(ld64(p + 8) & 0xffu);
				uint64_t prev = 0;
				p += 16;
				if (mode == BP_DELTA_FOR) {
					prev = ld64(p);
					p += 8;
				}
				if (width > 64) {
					return -4;
				}
				for (uint64_t g = 0; g < n; g += BP_ALGO_GROUP) {
					uint64_t tmp[BP_ALGO_GROUP];
					oracle_bp_unpack_group64(p + g * width / 8, tmp, width);
					const uint64_t m = (n - g) < BP_ALGO_GROUP ? (n - g) : BP_ALGO_GROUP;
					for (uint64_t i = 0; i < m; i++) {
						uint64_t v = tmp[i] + frame; /* ApplyFrameOfReference (wraps) */
						if (mode == BP_DELTA_FOR) {  /* DeltaDecode: running sum seeded with delta_offset */
							v += prev;
							prev = v;
						}
						o[g + i] = v;
					}
				}
			} else {
				return -5;
			}
		} else {
			uint32_t *o = (uint32_t *)out + done;
			if (mode == BP_CONSTANT) {
				const uint32_t c = ld32(p);
				for (uint64_t i = 0; i < n; i++) {
					o[i] = c;
				}
			} else if (mode == BP_CONSTANT_DELTA) {
				const uint32_t frame = ld32(p), delta = ld32(p + 4);
				for (uint64_t i = 0; i < n; i++) {
					o[i] = delta * (uint32_t)i + frame;
				}
			} else if (mode == BP_FOR || mode == BP_DELTA_FOR) {
				const uint32_t frame = ld32(p);
				const uint32_t width = ld32(p + 4) & 0xffu;
				uint32_t prev = 0;
				p += 8;
				if (mode == BP_DELTA_FOR) {
					prev = ld32(p);
					p += 4;
				}
				if (width > 32) {
					return -4;
				}
				for (uint64_t g = 0; g < n; g += BP_ALGO_GROUP) {
					uint32_t tmp[BP_ALGO_GROUP];
					oracle_bp_unpack_group32(p + g * width / 8, tmp, width);
					const uint64_t m = (n - g) < BP_ALGO_GROUP ? (n - g) : BP_ALGO_GROUP;
					for (uint64_t i = 0; i < m; i++) {
						uint32_t v = tmp[i] + frame;
						if (mode == BP_DELTA_FOR) {
							v += prev;
							prev = v;
						}
						o[g + i] = v;
					}
				}
			} else {
				return -5;
			}
		}
		done += n;
	}
	return 0;
}

/* ------------------------------------------------------------------ encoder (test data at scale)
 * Restates the reference's WRITER so that tests and the bench can make BitPacking segments of any size where
 * the reference binary is not available (the GPU box): BitpackingState::Flush (bitpacking.cpp:229-289) with
 * CalculateFORStats / CalculateDeltaStats (:148-215), the group writers (:392-446), FlushSegment (:524-544)
 * and BitpackingPrimitives::MinimumBitWidth / GetEffectiveWidth (bitpacking.hpp:84-87,139-172,201-209).
 * Pinned byte-for-byte against the segments the reference wrote for the same rows (tests/test_bitpacking.py).
 * The reference packs a ragged last 32-value group from an uninitialised stack buffer (PackBuffer,
 * bitpacking.hpp:43-58); those padding values are written as zero here.
 * One call = one segment holding all `count` values (the caller chooses the row ranges).
 *   values     : count elements of elem_bytes (8 or 4), signed
 *   force_mode : BP_AUTO, or BP_CONSTANT / BP_CONSTANT_DELTA / BP_DELTA_FOR / BP_FOR as PRAGMA force_bitpacking_mode
 * returns the segment size in bytes, or a negative code (-1 bad args, -2 capacity, -3 group not encodable) */
typedef __int128 i128;

static int fits_t(i128 v, int tbits) {
	const i128 hi = ((i128)1 << (tbits - 1)) - 1, lo = -hi - 1;
	return v >= lo && v <= hi;
}
static uint32_t effective_width(uint32_t w, int tbits) { /* GetEffectiveWidth */
	return (w + (uint32_t)tbits / 8u > (uint32_t)tbits) ? (uint32_t)tbits : w;
}
static uint32_t width_unsigned(uint64_t v, int tbits) { /* FindMinimumBitWidth<T, false> */
	uint32_t w = 0;
	if (v == 0) {
		return 0;
	}
	while (v) {
		w++;
		v >>= 1;
	}
	return effective_width(w, tbits);
}
static uint32_t width_signed(int64_t v, int tbits) { /* FindMinimumBitWidth<T, true>(v, v), v >= 0 here */
	uint32_t w = 1;
	if (v == 0) {
		return 0;
	}
	while (v) {
		w++;
		v >>= 1;
	}
	return effective_width(w, tbits);
}
static void st_t(uint8_t *p, uint64_t v, uint32_t elem_bytes) {
	memcpy(p, &v, elem_bytes); /* little endian */
}
/* fastpack: value i at bit i*width of the little-endian 32-bit word stream; n values, padded to 32 */
static void pack_run(uint8_t *dst, const uint64_t *vals, uint32_t n, uint32_t width) {
	const uint32_t n_pad = (n + 31u) & ~31u;
	memset(dst, 0, (size_t)n_pad * width / 8);
	for (uint32_t i = 0; i < n; i++) {
		uint64_t pos = (uint64_t)i * width;
		uint32_t put = 0;
		while (put < width) {
			const uint32_t sh = (uint32_t)((pos + put) & 31u);
			const uint32_t take = (32u - sh) < (width - put) ? (32u - sh) : (width - put);
			const uint64_t piece = (vals[i] >> put) & (take == 32 ? 0xffffffffull : ((1ull << take) - 1));
			uint32_t w = ld32(dst + 4 * ((pos + put) >> 5));
			w |= (uint32_t)(piece << sh);
			memcpy(dst + 4 * ((pos + put) >> 5), &w, 4);
			put += take;
		}
	}
}

ORACLE_API int64_t oracle_bitpacking_encode(const void *values, uint64_t count, uint32_t elem_bytes,
                                            uint32_t force_mode, uint8_t *out, uint64_t cap) {
	if ((elem_bytes != 8 && elem_bytes != 4) || count == 0) {
		return -1;
	}
	const int tbits = (int)elem_bytes * 8;
	const uint64_t tmask = elem_bytes == 8 ? ~0ull : 0xffffffffull;
	const uint64_t n_grp = (count + BP_META_GROUP - 1) / BP_META_GROUP;
	if (cap < 8 + n_grp * (3 * 8 + BP_META_GROUP * 8 + 4) + 8) {
		return -2; /* conservative bound: the caller allocates generously */
	}
	uint64_t data_ptr = 8;
	static __thread uint32_t meta[1 << 16];
	if (n_grp > (1u << 16)) {
		return -1;
	}
	for (uint64_t gi = 0; gi < n_grp; gi++) {
		const uint32_t n = (uint32_t)((count - gi * BP_META_GROUP) < BP_META_GROUP ? (count - gi * BP_META_GROUP) : BP_META_GROUP);
		int64_t v[BP_META_GROUP] = {0};
		uint64_t rel[BP_META_GROUP];
		for (uint32_t i = 0; i < n; i++) {
			v[i] = elem_bytes == 8 ? ((const int64_t *)values)[gi * BP_META_GROUP + i]
			                       : (int64_t)((const int32_t *)values)[gi * BP_META_GROUP + i];
		}
		int64_t mn = v[0], mx = v[0];
		for (uint32_t i = 1; i < n; i++) {
			mn = v[i] < mn ? v[i] : mn;
			mx = v[i] > mx ? v[i] : mx;
		}
		uint32_t mode;
		if (mx == mn && (force_mode == BP_AUTO || force_mode == BP_CONSTANT)) {
			mode = BP_CONSTANT;
			meta[gi] = (mode << 24) | (uint32_t)data_ptr;
			st_t(out + data_ptr, (uint64_t)mx, elem_bytes);
			data_ptr += elem_bytes;
			continue;
		}
		const int can_do_for = fits_t((i128)mx - mn, tbits);
		const int64_t min_max_diff = can_do_for ? (int64_t)((i128)mx - mn) : 0;
		/* CalculateDeltaStats */
		int can_do_delta = 0;
		int64_t delta[BP_META_GROUP], min_d = INT64_MAX, max_d = INT64_MIN, min_max_delta_diff = 0, delta_offset = 0;
		if (n >= 2) {
			int ok = 1;
			for (uint32_t i = 1; i < n && ok; i++) {
				const i128 d = (i128)v[i] - v[i - 1];
				ok = fits_t(d, tbits);
				delta[i] = (int64_t)d;
			}
			if (ok) {
				for (uint32_t i = 1; i < n; i++) {
					max_d = delta[i] > max_d ? delta[i] : max_d;
					min_d = delta[i] < min_d ? delta[i] : min_d;
				}
				delta[0] = min_d;
				can_do_delta = fits_t((i128)max_d - min_d, tbits) && fits_t((i128)v[0] - min_d, tbits);
				if (can_do_delta) {
					min_max_delta_diff = (int64_t)((i128)max_d - min_d);
					delta_offset = (int64_t)((i128)v[0] - min_d);
				}
			}
		}
		if (can_do_delta) {
			if (max_d == min_d && force_mode != BP_FOR && force_mode != BP_DELTA_FOR) {
				mode = BP_CONSTANT_DELTA;
				meta[gi] = (mode << 24) | (uint32_t)data_ptr;
				st_t(out + data_ptr, (uint64_t)v[0], elem_bytes);
				st_t(out + data_ptr + elem_bytes, (uint64_t)max_d, elem_bytes);
				data_ptr += 2 * elem_bytes;
				continue;
			}
			const uint32_t delta_w = width_unsigned((uint64_t)min_max_delta_diff & tmask, tbits);
			const uint32_t regular_w = width_signed(min_max_diff, tbits);
			if (delta_w < regular_w && force_mode != BP_FOR) {
				mode = BP_DELTA_FOR;
				meta[gi] = (mode << 24) | (uint32_t)data_ptr;
				for (uint32_t i = 0; i < n; i++) {
					rel[i] = ((uint64_t)delta[i] - (uint64_t)min_d) & tmask;
				}
				st_t(out + data_ptr, (uint64_t)min_d, elem_bytes);
				st_t(out + data_ptr + elem_bytes, (uint64_t)delta_w, elem_bytes);
				st_t(out + data_ptr + 2 * elem_bytes, (uint64_t)delta_offset, elem_bytes);
				data_ptr += 3 * elem_bytes;
				pack_run(out + data_ptr, rel, n, delta_w);
				data_ptr += (uint64_t)((n + 31u) & ~31u) * delta_w / 8;
				continue;
			}
		}
		if (!can_do_for) {
			return -3;
		}
		{
			const uint32_t width = width_unsigned((uint64_t)min_max_diff & tmask, tbits);
			mode = BP_FOR;
			meta[gi] = (mode << 24) | (uint32_t)data_ptr;
			for (uint32_t i = 0; i < n; i++) {
				rel[i] = ((uint64_t)v[i] - (uint64_t)mn) & tmask;
			}
			st_t(out + data_ptr, (uint64_t)mn, elem_bytes);
			st_t(out + data_ptr + elem_bytes, (uint64_t)width, elem_bytes);
			data_ptr += 2 * elem_bytes;
			pack_run(out + data_ptr, rel, n, width);
			data_ptr += (uint64_t)((n + 31u) & ~31u) * width / 8;
		}
		if (data_ptr > 0x00ffffffull) {
			return -2; /* metadata offsets are 24 bits: the caller must use smaller segments */
		}
	}
	/* FlushSegment: metadata right after the 8-byte aligned data, first group at the highest address */
	const uint64_t meta_off = (data_ptr + 7) & ~7ull;
	memset(out + data_ptr, 0, meta_off - data_ptr);
	for (uint64_t gi = 0; gi < n_grp; gi++) {
		memcpy(out + meta_off + 4 * (n_grp - 1 - gi), &meta[gi], 4);
	}
	const uint64_t total = meta_off + 4 * n_grp;
	memcpy(out, &total, 8);
	return (int64_t)total;
}
