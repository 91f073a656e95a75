/*
 * wah_oracle.c — CPU restatement of the WAH (word-aligned hybrid) bitvector code of FastBit's
 * ibis::bitvector, the compressed form in which the upstream CUBIT library keeps its value bitvectors.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT (same rule as cubit_oracle.c).
 *
 * PARITY STATUS: UNPINNED against source — FastBit / junchangwang/CUBIT are not in /root/reference and not
 * vendored.  The format is restated from its published description (K. Wu, E. Otoo,
 * A. Shoshani, "Optimizing bitmap indices with efficient compression", ACM TODS 31(1), 2006, §3 / Fig. 1;
 * 32-bit words):
 *   literal word   MSB = 0, the low 31 bits hold 31 consecutive bits of the bitmap, FIRST bit in the MOST
 *                  significant of the 31 positions
 *   fill word      MSB = 1, bit 30 = the fill bit, low 30 bits = number of 31-bit groups the fill spans
 *   active word    the last (bitmap length mod 31) bits, kept apart: value in the low `nbits` bits, first bit
 *                  most significant of those
 * The writer follows FastBit's append rule (a lone all-zero / all-one group is written as a literal, fills start at
 * two groups).  Known answers: the paper's 128-bit example, 1·0^20·1^3·0^79·1^25 → 40000380 80000002 001FFFFF +
 * active 0000000F (4 bits), and the 40 vectors of tests/golden/wah_vectors.json, written by an independent
 * pure-Python encoder (tests/golden/make_wah_golden.py) — all checked in tests/test_wah.py.
 */
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))
#define WAH_GROUP 31u
#define WAH_ALLONES 0x7fffffffu
#define WAH_MAXCNT 0x3fffffffu

static inline uint32_t bit_at(const uint64_t *words, uint64_t r) {
	return (uint32_t)((words[r >> 6] >> (r & 63)) & 1u);
}

/* bitmap (DuckDB bit order: row r = bit r%64 of word r/64) of n_bits rows → WAH.
 * out: capacity cap words.  Returns the number of WAH words (or -1 if cap is too small);
 * *active_val / *active_nbits receive the trailing n_bits % 31 bits. */
ORACLE_API int64_t oracle_wah_encode(const uint64_t *words, uint64_t n_bits, uint32_t *out, uint64_t cap,
                                     uint32_t *active_val, uint32_t *active_nbits) {
	uint64_t n_out = 0;
	const uint64_t n_groups = n_bits / WAH_GROUP;
	for (uint64_t g = 0; g < n_groups; g++) {
		uint32_t lit = 0;
		for (uint32_t b = 0; b < WAH_GROUP; b++) {
			lit = (lit << 1) | bit_at(words, g * WAH_GROUP + b); /* first bit ends up most significant */
		}
		/* append rule of the FastBit writer (ibis::bitvector::append_active, bitvector.h): a single all-zero /
		 * all-one group goes out as a LITERAL word; a second one in a row turns that literal into a fill of 2;
		 * further ones increment the fill */
		if (lit == 0 || lit == WAH_ALLONES) {
			const uint32_t fill = 0x80000000u | (lit ? 0x40000000u : 0u);
			if (n_out && out[n_out - 1] == lit) {
				out[n_out - 1] = fill | 2u;
				continue;
			}
			if (n_out && (out[n_out - 1] & 0xc0000000u) == fill && (out[n_out - 1] & WAH_MAXCNT) < WAH_MAXCNT) {
				out[n_out - 1]++;
				continue;
			}
		}
		if (n_out >= cap) {
			return -1;
		}
		out[n_out++] = lit;
	}
	uint32_t av = 0;
	const uint32_t an = (uint32_t)(n_bits % WAH_GROUP);
	for (uint32_t b = 0; b < an; b++) {
		av = (av << 1) | bit_at(words, n_groups * WAH_GROUP + b);
	}
	*active_val = av;
	*active_nbits = an;
	return (int64_t)n_out;
}

/* number of bitmap bits a WAH vector describes, or -1 if malformed (zero-length fill, active_nbits > 30) */
ORACLE_API int64_t oracle_wah_bits(const uint32_t *wah, uint64_t n_wah, uint32_t active_nbits) {
	uint64_t groups = 0;
	if (active_nbits >= WAH_GROUP) {
		return -1;
	}
	for (uint64_t i = 0; i < n_wah; i++) {
		if (wah[i] & 0x80000000u) {
			if ((wah[i] & WAH_MAXCNT) == 0) {
				return -1;
			}
			groups += wah[i] & WAH_MAXCNT;
		} else {
			groups++;
		}
	}
	return (int64_t)(groups * WAH_GROUP + active_nbits);
}

/* WAH → bitmap words (n_words zero-initialised by this function; bits past the described length stay 0).
 * Returns the number of bits described, or -1 if malformed / longer than n_words * 64. */
ORACLE_API int64_t oracle_wah_decode(const uint32_t *wah, uint64_t n_wah, uint32_t active_val, uint32_t active_nbits,
                                     uint64_t *words, uint64_t n_words) {
	const int64_t total = oracle_wah_bits(wah, n_wah, active_nbits);
	if (total < 0 || (uint64_t)total > n_words * 64) {
		return -1;
	}
	memset(words, 0, n_words * 8);
	uint64_t pos = 0;
	for (uint64_t i = 0; i < n_wah; i++) {
		if (wah[i] & 0x80000000u) {
			const uint64_t n = (uint64_t)(wah[i] & WAH_MAXCNT) * WAH_GROUP;
			if (wah[i] & 0x40000000u) {
				for (uint64_t r = pos; r < pos + n; r++) {
					words[r >> 6] |= 1ull << (r & 63);
				}
			}
			pos += n;
		} else {
			for (uint32_t b = 0; b < WAH_GROUP; b++) {
				if ((wah[i] >> (WAH_GROUP - 1 - b)) & 1u) {
					words[(pos + b) >> 6] |= 1ull << ((pos + b) & 63);
				}
			}
			pos += WAH_GROUP;
		}
	}
	for (uint32_t b = 0; b < active_nbits; b++) {
		if ((active_val >> (active_nbits - 1 - b)) & 1u) {
			words[(pos + b) >> 6] |= 1ull << ((pos + b) & 63);
		}
	}
	return total;
}
