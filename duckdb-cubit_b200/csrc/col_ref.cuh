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
// boundary (spans and slots are multiples of 1024 rows, ≤ 16384 rows).  Lanes 0..15 fetch those headers ONCE; every value
// then gets its header by shuffle, so the only global loads left per value are its 1-2 payload words.
struct HdrRegs {
	uint32_t blo, bhi, off, wid;
};

__device__ __forceinline__ HdrRegs load_hdrs(const ColRef &c, long long local0, int lane) {
	HdrRegs h = {0, 0, 0, 0};
	if (!c.raw && lane < 16) { // the header array is padded by 16 entries, so this never leaves it
		const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(c.hdr + (local0 >> 10) + lane));
		h.blo = raw.x;
		h.bhi = raw.y;
		h.off = raw.z;
		h.wid = raw.w;
	}
	return h;
}

// all lanes must call this (shuffles); `active` lanes get the value of row local0 + rel
__device__ __forceinline__ long long load_col_hoisted(const ColRef &c, const HdrRegs &h, long long local0, uint32_t rel,
                                                      bool active) {
	if (c.raw) {
		return active ? __ldg(c.raw + local0 + rel) : 0;
	}
	const int blk = active ? (int)(rel >> 10) : 0;
	const uint32_t blo = __shfl_sync(0xffffffffu, h.blo, blk), bhi = __shfl_sync(0xffffffffu, h.bhi, blk);
	const uint32_t off = __shfl_sync(0xffffffffu, h.off, blk), width = __shfl_sync(0xffffffffu, h.wid, blk);
	if (!active) {
		return 0;
	}
	const long long base = (long long)(((unsigned long long)bhi << 32) | blo);
	if (width == 0) {
		return base;
	}
	const uint32_t bit = (rel & (kPackBlock - 1)) * width;
	const unsigned long long *w = c.words + off + (bit >> 6);
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
