// col_ref.cuh — device-side read of one row of an 8-byte column, raw or FOR-bit-packed (kernels.h: ColRef)
#pragma once
#include "kernels.h"

namespace cubit {

__device__ __forceinline__ long long load_col(const ColRef &c, long long local) {
	if (c.raw) {
		return __ldg(c.raw + local);
	}
	// 16-byte block header, shared by 1024 consecutive rows (stays in L1/L2)
	const uint4 hraw = __ldg(reinterpret_cast<const uint4 *>(c.hdr + (local >> 10)));
	const long long base = (long long)(((unsigned long long)hraw.y << 32) | hraw.x);
	const uint32_t word_off = hraw.z, width = hraw.w;
	if (width == 0) {
		return base; // constant block
	}
	const unsigned long long bit = (unsigned long long)(local & (kPackBlock - 1)) * width;
	const unsigned long long *w = c.words + word_off + (bit >> 6);
	const unsigned sh = (unsigned)(bit & 63);
	unsigned long long v = __ldg(w) >> sh;
	if (sh + width > 64) {
		v |= __ldg(w + 1) << (64 - sh);
	}
	if (width < 64) {
		v &= (1ull << width) - 1;
	}
	return base + (long long)v;
}

// ---- warp-cooperative variant for position-ordered probes ---------------------------------------
// The rows a warp probes in one write-out come from at most 16 consecutive pack blocks starting at a block
// boundary (spans and slots are multiples of 1024 rows, ≤ 16384 rows).  Lanes 0..15 park those headers in the
// warp's shared-memory scratch ONCE (one LDS.128 per value afterwards, no shuffles); the per-value work is then
// an index multiply, two 32-bit loads that hit L1, one funnel shift and a mask (widths ≤ 32, the common case).
constexpr int kHdrSlots = 16; // pack-block headers staged per warp and column

__device__ __forceinline__ void stage_hdrs(const ColRef &c, long long local0, int lane, uint4 *hs) {
	if (!c.raw && lane < kHdrSlots) { // the header array is padded by 16 entries, so this never leaves it
		hs[lane] = __ldg(reinterpret_cast<const uint4 *>(c.hdr + (local0 >> 10) + lane));
	}
}

// `active` lanes get the value of row local0 + rel; hs = this warp's staged headers of the column
__device__ __forceinline__ long long load_col_staged(const ColRef &c, const uint4 *hs, long long local0, uint32_t rel,
                                                     bool active) {
	if (c.raw) {
		return active ? __ldg(c.raw + local0 + rel) : 0;
	}
	if (!active) {
		return 0;
	}
	const uint4 h = hs[rel >> 10];
	const long long base = (long long)(((unsigned long long)h.y << 32) | h.x);
	const uint32_t width = h.w;
	const uint32_t bit = (rel & (kPackBlock - 1)) * width;
	if (width <= 32u) { // branch-free down to width 0 (mask 0 → base); one spare word follows the payload
		const uint32_t *w32 = reinterpret_cast<const uint32_t *>(c.words + h.z) + (bit >> 5);
		const uint32_t lo = __ldg(w32), hi = __ldg(w32 + 1);
		const uint32_t mask = width >= 32u ? 0xffffffffu : (1u << width) - 1u;
		return base + (long long)(__funnelshift_r(lo, hi, bit & 31u) & mask);
	}
	const unsigned long long *w = c.words + h.z + (bit >> 6);
	const unsigned sh = bit & 63u;
	unsigned long long v = __ldg(w) >> sh;
	if (sh + width > 64) {
		v |= __ldg(w + 1) << (64 - sh);
	}
	if (width < 64) {
		v &= (1ull << width) - 1;
	}
	return base + (long long)v;
}

} // namespace cubit
