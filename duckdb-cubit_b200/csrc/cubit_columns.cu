#include "table.h"

#include <algorithm>
#include <cstring>

using namespace cubit;

// ------------------------------------------------------------------ columns
extern "C" int cubit_gpu_upload_column(cubit_gpu_table *t, int32_t col_id, const void *data, uint32_t elem_bytes,
                                       uint64_t n) {
	ABI_BEGIN
	if (!t || !data) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (elem_bytes != 4 && elem_bytes != 8) {
		return fail(CUBIT_EINVAL, "elem_bytes must be 4 or 8");
	}
	if (n != t->n_rows) {
		return fail(CUBIT_EINVAL, "column has %llu rows, table has %llu", (unsigned long long)n,
		            (unsigned long long)t->n_rows);
	}
	if (t->sharded()) {
		for (size_t s = 0; s < t->shards.size(); s++) {
			int rc = cubit_gpu_upload_column(t->shards[s], col_id, static_cast<const uint8_t *>(data) + t->shard_row0[s] * elem_bytes,
			                                 elem_bytes, t->shards[s]->n_rows);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Column &c = t->columns[col_id];
	if (c.packed() || (c.d && (c.elem != elem_bytes || c.n != n))) {
		CU_TRY(cudaStreamSynchronize(t->stream));
		free_column(c);
	}
	if (!c.d) {
		// + 16 bytes so a 128-bit load of the last aligned pair never leaves the allocation
		CU_TRY(cudaMalloc(&c.d, (size_t)n * elem_bytes + 16));
		c.cap = n;
	}
	c.elem = elem_bytes;
	c.n = n;
	CU_TRY(cudaMemcpyAsync(c.d, data, (size_t)n * elem_bytes, cudaMemcpyHostToDevice, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	if (c.d_valid) { // new contents: all valid until a mask is uploaded again
		cudaFree(c.d_valid);
		c.d_valid = nullptr;
		c.valid_cap_words = 0;
	}
	return CUBIT_OK;
	ABI_END
}

// NULLs of a column: its validity mask in the reference's layout (ValidityMask, validity_mask.hpp:50,163-168 —
// what a validity_uncompressed segment stores, validity_uncompressed.cpp:381).  The probe reports the validity
// of every projected value (cubit_gpu_fetch_validity) and aggregates skip NULL inputs.
extern "C" int cubit_gpu_upload_column_validity(cubit_gpu_table *t, int32_t col_id, const uint64_t *words,
                                                uint64_t n_words) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) { // shard boundaries are multiples of 64 rows
		if (words && n_words != t->n_words) {
			return fail(CUBIT_EINVAL, "validity mask has %llu words, table needs %llu", (unsigned long long)n_words,
			            (unsigned long long)t->n_words);
		}
		for (size_t s = 0; s < t->shards.size(); s++) {
			int rc = cubit_gpu_upload_column_validity(t->shards[s], col_id, words ? words + t->shard_row0[s] / 64 : nullptr,
			                                          words ? t->shards[s]->n_words : 0);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	Column &c = it->second;
	CU_TRY(cudaStreamSynchronize(t->stream));
	if (!words) { // drop the mask: every row valid
		if (c.d_valid) {
			cudaFree(c.d_valid);
		}
		c.d_valid = nullptr;
		c.valid_cap_words = 0;
		return CUBIT_OK;
	}
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "validity mask has %llu words, table needs %llu", (unsigned long long)n_words,
		            (unsigned long long)t->n_words);
	}
	if (c.valid_cap_words < n_words) {
		if (c.d_valid) {
			cudaFree(c.d_valid);
			c.d_valid = nullptr;
		}
		CU_TRY(cudaMalloc((void **)&c.d_valid, (n_words + 2) * 8));
		c.valid_cap_words = n_words;
	}
	CU_TRY(cudaMemcpyAsync(c.d_valid, words, n_words * 8, cudaMemcpyHostToDevice, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	return CUBIT_OK;
	ABI_END
}

// Decode the reference's on-disk column segments on the GPU (column_decode.cu).  Everything the kernel will
// dereference is bounds-checked here, on the host copy of the segment, so a malformed segment is an error
// return and never an out-of-bounds device access.
extern "C" int cubit_gpu_upload_column_segments(cubit_gpu_table *t, int32_t col_id, uint32_t elem_bytes,
                                                const cubit_column_segment *segs, uint32_t n_segs,
                                                cubit_decode_info *info) {
	ABI_BEGIN
	if (!t || (!segs && n_segs)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (elem_bytes != 4 && elem_bytes != 8) {
		return fail(CUBIT_EINVAL, "elem_bytes must be 4 or 8");
	}
	if (t->sharded()) {
		// on-disk segments do not end on shard boundaries: the caller decodes per shard or uploads decoded rows
		return fail(CUBIT_ESTATE, "segment upload addresses one shard (upload decoded rows to a sharded table)");
	}
	auto ld32 = [](const uint8_t *p) {
		uint32_t v;
		memcpy(&v, p, 4);
		return v;
	};
	auto ld64 = [](const uint8_t *p) {
		uint64_t v;
		memcpy(&v, p, 8);
		return v;
	};
	// ---- plan: validate, lay the segments out in one blob, build the group directory
	std::vector<BpGroup> groups;
	std::vector<RleTile> tiles;
	std::vector<uint64_t> seg_off(n_segs, 0), seg_used(n_segs, 0); // blob offset / bytes staged per segment
	auto ld16 = [](const uint8_t *p) {
		uint16_t v;
		memcpy(&v, p, 2);
		return v;
	};
	cubit_decode_info di;
	memset(&di, 0, sizeof(di));
	uint64_t blob_bytes = 0, next_row = 0;
	for (uint32_t si = 0; si < n_segs; si++) {
		const cubit_column_segment &sg = segs[si];
		if (sg.row_start != next_row || sg.count == 0 || !sg.data) {
			return fail(CUBIT_EINVAL, "segment %u: segments must tile the rows in order (row_start %llu, expected %llu)", si,
			            (unsigned long long)sg.row_start, (unsigned long long)next_row);
		}
		next_row += sg.count;
		const uint8_t *p = static_cast<const uint8_t *>(sg.data);
		if (sg.kind == CUBIT_SEG_UNCOMPRESSED) {
			if (sg.bytes < sg.count * elem_bytes) {
				return fail(CUBIT_EINVAL, "segment %u: %llu bytes for %llu uncompressed rows", si,
				            (unsigned long long)sg.bytes, (unsigned long long)sg.count);
			}
			continue; // copied straight into the column
		}
		seg_off[si] = blob_bytes;
		if (sg.kind == CUBIT_SEG_CONSTANT) {
			if (sg.bytes < elem_bytes || sg.count > 0xffffffffull) {
				return fail(CUBIT_EINVAL, "segment %u: bad constant segment", si);
			}
			groups.push_back(BpGroup {blob_bytes, sg.row_start, (uint32_t)sg.count, BP_CONSTANT});
			di.mode_groups[BP_CONSTANT]++;
			blob_bytes += 8;
			continue;
		}
		if (sg.kind == CUBIT_SEG_RLE) {
			// [u64 offset of the run lengths][values][pad][u16 run lengths] (rle.cpp:190-205).  The run count is not
			// stored: walk the lengths until the segment's rows are covered (what RLEScanPartialInternal does,
			// :338-364), cutting tiles of ≤ kRleTileRuns runs / ~128 K rows as we go.
			if (sg.bytes < 8) {
				return fail(CUBIT_EINVAL, "segment %u: bad RLE segment size %llu", si, (unsigned long long)sg.bytes);
			}
			const uint64_t off = ld64(p);
			if (off < 8 || (off & 7) || off > sg.bytes) {
				return fail(CUBIT_EINVAL, "segment %u: RLE run-length offset %llu outside the segment", si,
				            (unsigned long long)off);
			}
			const uint64_t max_runs = (off - 8) / elem_bytes;
			uint64_t produced = 0, run = 0;
			RleTile tl {blob_bytes + 8, blob_bytes + off, sg.row_start, 0, 0};
			while (produced < sg.count) {
				if (run >= max_runs || off + 2 * (run + 1) > sg.bytes) {
					return fail(CUBIT_EINVAL, "segment %u: RLE runs end after %llu of %llu rows", si,
					            (unsigned long long)produced, (unsigned long long)sg.count);
				}
				const uint64_t len = ld16(p + off + 2 * run);
				if (len == 0) {
					return fail(CUBIT_EINVAL, "segment %u: RLE run %llu has length 0", si, (unsigned long long)run);
				}
				const uint64_t take = std::min<uint64_t>(len, sg.count - produced);
				if (tl.n_runs == (uint32_t)kRleTileRuns || (tl.n_runs && tl.n_rows + take > 131072)) {
					tiles.push_back(tl);
					tl = RleTile {blob_bytes + 8 + run * elem_bytes, blob_bytes + off + 2 * run, sg.row_start + produced, 0, 0};
				}
				tl.n_runs++;
				tl.n_rows += (uint32_t)take;
				produced += take;
				run++;
			}
			tiles.push_back(tl);
			di.rle_runs += run;
			seg_used[si] = off + 2 * run;
			blob_bytes += (seg_used[si] + 7) & ~7ull;
			continue;
		}
		if (sg.kind != CUBIT_SEG_BITPACKING) {
			return fail(CUBIT_EINVAL, "segment %u: unknown kind %u", si, sg.kind);
		}
		const uint64_t n_grp = (sg.count + 2047) / 2048;
		if (sg.bytes < 12 || (sg.bytes & 3)) {
			return fail(CUBIT_EINVAL, "segment %u: bad size %llu", si, (unsigned long long)sg.bytes);
		}
		const uint64_t meta_end = ld64(p); // BitpackingScanState ctor, bitpacking.cpp:633-636
		if (meta_end > sg.bytes || (meta_end & 3) || meta_end < 8 + 4 * n_grp) {
			return fail(CUBIT_EINVAL, "segment %u: metadata end %llu outside the segment (%llu bytes, %llu groups)", si,
			            (unsigned long long)meta_end, (unsigned long long)sg.bytes, (unsigned long long)n_grp);
		}
		const uint64_t data_end = meta_end - 4 * n_grp; // group data lives in [8, data_end)
		for (uint64_t gi = 0; gi < n_grp; gi++) {
			const uint32_t enc = ld32(p + meta_end - 4 * (gi + 1)); // DecodeMeta, bitpacking.cpp:68-73
			const uint32_t mode = enc >> 24, off = enc & 0x00ffffffu;
			const uint32_t n = (uint32_t)std::min<uint64_t>(2048, sg.count - gi * 2048);
			uint64_t need; // bytes of the group at `off`
			if (mode == BP_CONSTANT) {
				need = elem_bytes;
			} else if (mode == BP_CONSTANT_DELTA) {
				need = 2 * elem_bytes;
			} else if (mode == BP_FOR || mode == BP_DELTA_FOR) {
				need = (mode == BP_FOR ? 2 : 3) * (uint64_t)elem_bytes;
				if (off < 8 || (off & 3) || off + need > data_end) {
					return fail(CUBIT_EINVAL, "segment %u group %llu: header outside the segment", si, (unsigned long long)gi);
				}
				const uint32_t width = (uint32_t)(elem_bytes == 8 ? ld64(p + off + 8) : ld32(p + off + 4)) & 0xffu;
				if (width > elem_bytes * 8) {
					return fail(CUBIT_EINVAL, "segment %u group %llu: bit width %u", si, (unsigned long long)gi, width);
				}
				need += (uint64_t)((n + 31) / 32) * width * 4; // GetRequiredSize, bitpacking.hpp:103-106
			} else {
				return fail(CUBIT_EINVAL, "segment %u group %llu: invalid bitpacking mode %u", si, (unsigned long long)gi, mode);
			}
			if (off < 8 || (off & 3) || off + need > data_end) {
				return fail(CUBIT_EINVAL, "segment %u group %llu: data [%u, +%llu) outside the segment", si,
				            (unsigned long long)gi, off, (unsigned long long)need);
			}
			groups.push_back(BpGroup {blob_bytes + off, sg.row_start + gi * 2048, n, mode});
			di.mode_groups[mode]++;
		}
		blob_bytes += (sg.bytes + 7) & ~7ull;
	}
	if (next_row != t->n_rows) {
		return fail(CUBIT_EINVAL, "segments cover %llu rows, table has %llu", (unsigned long long)next_row,
		            (unsigned long long)t->n_rows);
	}
	if (groups.size() > 0x7fffffffull || tiles.size() > 0x7fffffffull) {
		return fail(CUBIT_EINVAL, "too many metadata groups");
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Column &c = t->columns[col_id];
	if (c.packed() || (c.d && (c.elem != elem_bytes || c.n != t->n_rows))) {
		CU_TRY(cudaStreamSynchronize(t->stream));
		free_column(c);
	}
	if (!c.d) {
		CU_TRY(cudaMalloc(&c.d, (size_t)t->n_rows * elem_bytes + 16));
		c.cap = t->n_rows;
	}
	c.elem = elem_bytes;
	c.n = t->n_rows;
	// ---- compressed bytes host → device as stored
	uint8_t *d_blob = nullptr;
	BpGroup *d_groups = nullptr;
	RleTile *d_tiles = nullptr;
	auto cleanup = [&]() {
		if (d_blob) {
			cudaFree(d_blob);
		}
		if (d_groups) {
			cudaFree(d_groups);
		}
		if (d_tiles) {
			cudaFree(d_tiles);
		}
	};
#define CU_TRY_CLEAN(expr)                                                                                             \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			cleanup();                                                                                                 \
			return fail(_e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "%s: %s (%s:%d)", #expr,         \
			            cudaGetErrorString(_e), __FILE__, __LINE__);                                                   \
		}                                                                                                              \
	} while (0)
	CU_TRY_CLEAN(cudaMalloc(&d_blob, blob_bytes + 16)); // + 16: the kernel never reads past a group, this is slack
	CU_TRY_CLEAN(cudaMalloc(&d_groups, (groups.size() + 1) * sizeof(BpGroup)));
	CU_TRY_CLEAN(cudaMalloc(&d_tiles, (tiles.size() + 1) * sizeof(RleTile)));
	// Thousands of sub-megabyte segments: gather them into two pinned staging chunks on the host and move each
	// chunk with ONE async copy (the next chunk is being filled while the previous one is on the wire).
	const uint64_t chunk = kStageChunk;
	cudaError_t pe = cudaSuccess;
	if (ensure_stage(t) != CUBIT_OK) {
		cleanup();
		return CUBIT_ECUDA;
	}
	uint8_t *const *stage = t->h_stage;
	cudaEvent_t *staged = t->stage_ev;
	auto cleanup_stage = []() {};
	uint64_t chunk_base = 0; // blob offset of the chunk being filled
	int cur = 0;
	bool used[2] = {false, false};
	auto flush_chunk = [&](uint64_t upto) -> cudaError_t { // send blob bytes [chunk_base, upto)
		cudaError_t e = cudaSuccess;
		if (upto > chunk_base) {
			e = cudaMemcpyAsync(d_blob + chunk_base, stage[cur], upto - chunk_base, cudaMemcpyHostToDevice, t->stream);
			if (e == cudaSuccess) {
				e = cudaEventRecord(staged[cur], t->stream);
			}
			used[cur] = true;
			cur ^= 1;
			if (e == cudaSuccess && used[cur]) {
				e = cudaEventSynchronize(staged[cur]); // the other chunk must have left the host before it is refilled
			}
			chunk_base = upto;
		}
		return e;
	};
	for (uint32_t si = 0; si < n_segs && pe == cudaSuccess; si++) {
		const cubit_column_segment &sg = segs[si];
		if (sg.kind == CUBIT_SEG_UNCOMPRESSED) {
			pe = cudaMemcpyAsync(static_cast<uint8_t *>(c.d) + sg.row_start * elem_bytes, sg.data, sg.count * elem_bytes,
			                     cudaMemcpyHostToDevice, t->stream);
			di.h2d_bytes += sg.count * elem_bytes;
			continue;
		}
		const uint64_t nb = sg.kind == CUBIT_SEG_CONSTANT ? elem_bytes : (sg.kind == CUBIT_SEG_RLE ? seg_used[si] : sg.bytes);
		const uint8_t *src = static_cast<const uint8_t *>(sg.data);
		uint64_t done = 0;
		while (done < nb && pe == cudaSuccess) { // a segment may straddle chunks
			const uint64_t at = seg_off[si] + done;
			if (at >= chunk_base + chunk) {
				pe = flush_chunk(chunk_base + chunk);
				continue;
			}
			const uint64_t take = std::min<uint64_t>(nb - done, chunk_base + chunk - at);
			memcpy(stage[cur] + (at - chunk_base), src + done, take);
			done += take;
		}
		di.h2d_bytes += nb;
	}
	if (pe == cudaSuccess) {
		pe = flush_chunk(blob_bytes);
	}
	if (pe == cudaSuccess) {
		pe = cudaStreamSynchronize(t->stream); // staging buffers are freed below
	}
	cleanup_stage();
	CU_TRY_CLEAN(pe);
	CU_TRY_CLEAN(cudaMemcpyAsync(d_groups, groups.data(), groups.size() * sizeof(BpGroup), cudaMemcpyHostToDevice,
	                             t->stream));
	CU_TRY_CLEAN(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(RleTile), cudaMemcpyHostToDevice,
	                             t->stream));
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	CU_TRY_CLEAN(cudaEventCreate(&e0));
	CU_TRY_CLEAN(cudaEventCreate(&e1));
	cudaEventRecord(e0, t->stream);
	cudaError_t le = launch_bp_decode(d_blob, d_groups, (uint32_t)groups.size(), c.d, elem_bytes, t->stream);
	if (le == cudaSuccess) {
		le = launch_rle_decode(d_blob, d_tiles, (uint32_t)tiles.size(), c.d, elem_bytes, t->stream);
	}
	cudaEventRecord(e1, t->stream);
	cudaError_t se = cudaStreamSynchronize(t->stream);
	if (le == cudaSuccess && se == cudaSuccess) {
		cudaEventElapsedTime(&di.ms_decode, e0, e1);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	CU_TRY_CLEAN(le);
	CU_TRY_CLEAN(se);
#undef CU_TRY_CLEAN
	cleanup();
	di.n_launches = (groups.empty() ? 0u : 1u) + (tiles.empty() ? 0u : 1u);
	t->launches += di.n_launches;
	di.n_groups = groups.size();
	if (info) {
		*info = di;
	}
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_download_column(cubit_gpu_table *t, int32_t col_id, void *data, uint32_t elem_bytes,
                                         uint64_t n) {
	ABI_BEGIN
	if (!t || !data) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		if (n > t->n_rows) {
			return fail(CUBIT_EINVAL, "column %d shape mismatch", col_id);
		}
		for (size_t s = 0; s < t->shards.size() && t->shard_row0[s] < n; s++) {
			const uint64_t take = std::min<uint64_t>(t->shards[s]->n_rows, n - t->shard_row0[s]);
			int rc = cubit_gpu_download_column(t->shards[s], col_id, static_cast<uint8_t *>(data) + t->shard_row0[s] * elem_bytes,
			                                   elem_bytes, take);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	if (it->second.elem != elem_bytes || n > it->second.n) {
		return fail(CUBIT_EINVAL, "column %d shape mismatch", col_id);
	}
	const Column &c = it->second;
	if (c.d) {
		CU_TRY(cudaMemcpyAsync(data, c.d, (size_t)n * elem_bytes, cudaMemcpyDeviceToHost, t->stream));
		CU_TRY(cudaStreamSynchronize(t->stream));
		return CUBIT_OK;
	}
	// only the packed form is resident: fetch it and decode on the host (diagnostic path)
	const uint64_t n_blk = (c.n + kPackBlock - 1) / kPackBlock;
	std::vector<PackHdr> hdr(n_blk);
	std::vector<unsigned long long> words(c.packed_bytes / 8);
	CU_TRY(cudaMemcpyAsync(hdr.data(), c.d_hdr, n_blk * sizeof(PackHdr), cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaMemcpyAsync(words.data(), c.d_words, c.packed_bytes, cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	long long *out = static_cast<long long *>(data);
	for (uint64_t r = 0; r < n; r++) {
		const PackHdr &h = hdr[r / kPackBlock];
		unsigned long long v = 0;
		if (h.width) {
			const uint64_t bit = (r % kPackBlock) * h.width;
			const unsigned sh = (unsigned)(bit & 63);
			v = words[h.word_off + (bit >> 6)] >> sh;
			if (sh + h.width > 64) {
				v |= words[h.word_off + (bit >> 6) + 1] << (64 - sh);
			}
			if (h.width < 64) {
				v &= (1ull << h.width) - 1;
			}
		}
		out[r] = h.base + (long long)v;
	}
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_synth_column(cubit_gpu_table *t, int32_t col_id, int32_t kind, uint64_t seed,
                                      uint64_t threshold, uint32_t card, uint32_t hot_lo, uint32_t hot_n) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (kind < 0 || kind > 3) {
		return fail(CUBIT_EINVAL, "kind must be 0..3");
	}
	if ((kind == 2 && card == 0) || (kind == 3 && threshold == 0)) {
		return fail(CUBIT_EINVAL, "empty value range");
	}
	if (kind == 1 && (hot_n == 0 || hot_n >= card || hot_lo + hot_n > card)) {
		return fail(CUBIT_EINVAL, "bad hot range");
	}
	if (t->sharded()) { // the generators are seeded by global row id (row_base + r): any shard count yields the same table
		for (auto *s : t->shards) {
			int rc = cubit_gpu_synth_column(s, col_id, kind, seed, threshold, card, hot_lo, hot_n);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	const uint32_t elem = (kind == 0 || kind == 3) ? 8 : 4;
	Column &c = t->columns[col_id];
	if (c.packed() || (c.d && (c.elem != elem || c.n != t->n_rows))) {
		CU_TRY(cudaStreamSynchronize(t->stream));
		free_column(c);
	}
	if (!c.d) {
		CU_TRY(cudaMalloc(&c.d, (size_t)t->n_rows * elem + 16));
		c.cap = t->n_rows;
	}
	c.elem = elem;
	c.n = t->n_rows;
	CU_TRY(launch_synth_column(c.d, kind, t->n_rows, t->row_base, seed, threshold, card, hot_lo, hot_n, t->sm_count,
	                           t->stream));
	t->launches++;
	CU_TRY(cudaStreamSynchronize(t->stream));
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_pack_column(cubit_gpu_table *t, int32_t col_id, int keep_raw, uint64_t *packed_bytes) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		uint64_t total = 0;
		for (auto *s : t->shards) {
			uint64_t b = 0;
			int rc = cubit_gpu_pack_column(s, col_id, keep_raw, &b);
			if (rc) {
				return rc;
			}
			total += b;
		}
		if (packed_bytes) {
			*packed_bytes = total;
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	Column &c = it->second;
	if (c.elem != 8) {
		return fail(CUBIT_EINVAL, "only 8-byte columns can be bit-packed");
	}
	if (c.packed()) {
		if (packed_bytes) {
			*packed_bytes = c.packed_bytes;
		}
		return CUBIT_OK;
	}
	const uint64_t n_blk = (c.n + kPackBlock - 1) / kPackBlock;
	long long *d_base = nullptr;
	uint32_t *d_width = nullptr;
	CU_TRY(cudaMalloc(&d_base, n_blk * sizeof(long long)));
	CU_TRY(cudaMalloc(&d_width, n_blk * sizeof(uint32_t)));
	CU_TRY(launch_pack_widths(static_cast<const long long *>(c.d), c.n, d_base, d_width, t->stream));
	t->launches++;
	std::vector<long long> base(n_blk);
	std::vector<uint32_t> width(n_blk);
	CU_TRY(cudaMemcpyAsync(base.data(), d_base, n_blk * sizeof(long long), cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaMemcpyAsync(width.data(), d_width, n_blk * sizeof(uint32_t), cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	cudaFree(d_base);
	cudaFree(d_width);
	std::vector<PackHdr> hdr(n_blk);
	uint64_t off = 0;
	uint32_t max_width = 0;
	for (uint64_t b = 0; b < n_blk; b++) {
		hdr[b].base = base[b];
		hdr[b].width = width[b];
		max_width = std::max(max_width, width[b]);
		if (off > 0xffffffffull) {
			return fail(CUBIT_EINVAL, "packed column exceeds 32 GiB");
		}
		hdr[b].word_off = (uint32_t)off;
		off += 16ull * width[b];
	}
	if (keep_raw == 2 && max_width > 32) { // "both forms if the dense probe can use the packed one": it cannot
		if (packed_bytes) {
			*packed_bytes = 0;
		}
		return CUBIT_OK;
	}
	const uint64_t bytes = (off + 2) * 8; // + spare words: the decoder may read one word past a value
	CU_TRY(cudaMalloc(&c.d_hdr, (n_blk + 16) * sizeof(PackHdr))); // + 16: load_hdrs reads a whole span's headers
	CU_TRY(cudaMemsetAsync(c.d_hdr + n_blk, 0, 16 * sizeof(PackHdr), t->stream));
	CU_TRY(cudaMalloc(&c.d_words, bytes));
	CU_TRY(cudaMemsetAsync(c.d_words + off, 0, 16, t->stream));
	CU_TRY(cudaMemcpyAsync(c.d_hdr, hdr.data(), n_blk * sizeof(PackHdr), cudaMemcpyHostToDevice, t->stream));
	CU_TRY(launch_pack_blocks(static_cast<const long long *>(c.d), c.n, c.d_hdr, c.d_words, t->stream));
	t->launches++;
	CU_TRY(cudaStreamSynchronize(t->stream));
	c.packed_bytes = bytes;
	c.pack_max_width = max_width;
	c.pack_avg_width = n_blk ? (double)off * 64.0 / ((double)n_blk * kPackBlock) : 0.0;
	if (!keep_raw) {
		cudaFree(c.d);
		c.d = nullptr;
	}
	if (packed_bytes) {
		*packed_bytes = bytes + n_blk * sizeof(PackHdr);
	}
	return CUBIT_OK;
	ABI_END
}

// Append path (INSERT: new rows take the next row ids — DataTable::Append / BoundIndex::Append,
// src/include/duckdb/execution/index/bound_index.hpp:71-75; rowids are dense positions, row_group.cpp:511-514).
// Bitvectors are padded to whole segments, so appending inside the last segment touches no allocation; past it
// every index is re-strided once (capacity grows by half).  Indexes built from a column are extended on the GPU
// by the index-build kernel over the new rows only (compressed indexes: the touched tail is rebuilt from the
// source column).  Every new allocation is made BEFORE anything is swapped in, so a failed append leaves the
// table exactly as it was.
extern "C" int cubit_gpu_append_rows(cubit_gpu_table *t, uint64_t n_new, const cubit_append_column *cols,
                                     uint32_t n_cols) {
	ABI_BEGIN
	if (!t || (!cols && n_cols)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n_new == 0) {
		return fail(CUBIT_EINVAL, "n_new must be > 0");
	}
	if (t->sharded()) { // new rows take the next row ids: they belong to the LAST shard
		cubit_gpu_table *last = t->shards.back();
		int rc = cubit_gpu_append_rows(last, n_new, cols, n_cols);
		if (rc) {
			return rc;
		}
		t->n_rows += n_new;
		t->n_words = (t->n_rows + 63) / 64;
		t->shard_row0.back() = t->n_rows;
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	const uint64_t old_n = t->n_rows, new_n = old_n + n_new;
	const uint64_t new_n_seg = (new_n + t->seg_bits - 1) / t->seg_bits;
	if (new_n_seg > 0x7fffffffull) {
		return fail(CUBIT_EINVAL, "too many segments");
	}
	if (n_cols != t->columns.size()) {
		return fail(CUBIT_EINVAL, "append must supply all %zu resident columns (got %u)", t->columns.size(), n_cols);
	}
	for (uint32_t i = 0; i < n_cols; i++) {
		auto it = t->columns.find(cols[i].col_id);
		if (it == t->columns.end() || !cols[i].data) {
			return fail(CUBIT_EINVAL, "append: no resident column %d (or NULL data)", cols[i].col_id);
		}
		for (uint32_t j = 0; j < i; j++) {
			if (cols[j].col_id == cols[i].col_id) {
				return fail(CUBIT_EINVAL, "append: column %d listed twice", cols[i].col_id);
			}
		}
		if (it->second.elem != cols[i].elem_bytes) {
			return fail(CUBIT_EINVAL, "append: column %d is %u bytes wide", cols[i].col_id, it->second.elem);
		}
		if (it->second.packed() && !it->second.d) {
			return fail(CUBIT_ESTATE, "append: column %d is resident only bit-packed; appends need the raw form",
			            cols[i].col_id);
		}
	}
	for (Index *ix : t->indexes) {
		if ((uint64_t)ix->card * new_n_seg > (1ull << 30) && ix->delta.n_ent) {
			return fail(CUBIT_ESTATE, "append: merge the pending deltas first (cardinality * segments > 2^30)");
		}
	}
	// ---- phase 1: every allocation the append needs (nothing of the table is touched yet)
	struct NewBuf {
		void *p = nullptr;
		uint64_t cap = 0;
	};
	const bool restride = new_n_seg * t->seg_words > t->words_per_bv;
	const uint64_t cap_seg = restride ? ((std::max<uint64_t>(new_n_seg, (uint64_t)(t->words_per_bv / t->seg_words) * 3 / 2 + 1) + 3) & ~3ull)
	                                  : t->words_per_bv / t->seg_words; // (a multiple of 4 segments, as at creation)
	const uint64_t new_stride = cap_seg * t->seg_words;
	std::vector<NewBuf> nbits(t->indexes.size()), ncol(n_cols), nvalid(n_cols);
	auto release_new = [&]() {
		for (auto *v : {&nbits, &ncol, &nvalid}) {
			for (auto &b : *v) {
				if (b.p) {
					cudaFree(b.p);
				}
			}
		}
	};
	cudaError_t e = cudaSuccess;
	for (size_t i = 0; restride && i < t->indexes.size() && e == cudaSuccess; i++) {
		Index *ix = t->indexes[i];
		const size_t bytes = ix->compressed ? (size_t)ix->card * cap_seg * 8 : (size_t)ix->card * new_stride * 8;
		e = cudaMalloc(&nbits[i].p, bytes);
		nbits[i].cap = bytes;
	}
	for (uint32_t i = 0; i < n_cols && e == cudaSuccess; i++) {
		Column &c = t->columns[cols[i].col_id];
		if (new_n > c.cap) {
			ncol[i].cap = std::max<uint64_t>(new_n, c.cap + c.cap / 2);
			e = cudaMalloc(&ncol[i].p, (size_t)ncol[i].cap * c.elem + 16);
		}
		const uint64_t new_w = (new_n + 63) / 64;
		if (e == cudaSuccess && c.d_valid && new_w > c.valid_cap_words) {
			nvalid[i].cap = std::max<uint64_t>(new_w, c.valid_cap_words + c.valid_cap_words / 2);
			e = cudaMalloc(&nvalid[i].p, (nvalid[i].cap + 2) * 8);
		}
	}
	if (e != cudaSuccess) {
		release_new();
		return fail(e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "append: %s", cudaGetErrorString(e));
	}
	// ---- phase 2: fill the new buffers (device-to-device), still without touching the table
	cudaStream_t st = t->stream;
	for (size_t i = 0; restride && i < t->indexes.size() && e == cudaSuccess; i++) {
		Index *ix = t->indexes[i];
		e = cudaMemsetAsync(nbits[i].p, 0, nbits[i].cap, st);
		if (e != cudaSuccess) {
			break;
		}
		if (ix->compressed) {
			e = cudaMemcpy2DAsync(nbits[i].p, cap_seg * 8, ix->cs.d_dir, ix->cs.n_seg_cap * 8, ix->cs.n_seg_cap * 8, ix->card,
			                      cudaMemcpyDeviceToDevice, st);
		} else {
			e = cudaMemcpy2DAsync(nbits[i].p, new_stride * 8, ix->d_bits, t->words_per_bv * 8, t->words_per_bv * 8, ix->card,
			                      cudaMemcpyDeviceToDevice, st);
		}
	}
	const uint64_t old_w = (old_n + 63) / 64, new_w = (new_n + 63) / 64;
	for (uint32_t i = 0; i < n_cols && e == cudaSuccess; i++) {
		Column &c = t->columns[cols[i].col_id];
		if (ncol[i].p) {
			e = cudaMemcpyAsync(ncol[i].p, c.d, (size_t)old_n * c.elem, cudaMemcpyDeviceToDevice, st);
		}
		if (e == cudaSuccess && nvalid[i].p) {
			e = cudaMemcpyAsync(nvalid[i].p, c.d_valid, old_w * 8, cudaMemcpyDeviceToDevice, st);
		}
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(st); // ... which also drains every scan that still reads the old buffers
	}
	if (e != cudaSuccess) {
		release_new();
		return fail(CUBIT_ECUDA, "append: %s", cudaGetErrorString(e));
	}
	// ---- phase 3: swap in (no failure path from here on except the row copies themselves)
	if (restride) {
		for (size_t i = 0; i < t->indexes.size(); i++) {
			Index *ix = t->indexes[i];
			if (ix->compressed) {
				cudaFree(ix->cs.d_dir);
				ix->cs.d_dir = static_cast<unsigned long long *>(nbits[i].p);
				ix->cs.n_seg_cap = cap_seg;
			} else {
				cudaFree(ix->d_bits);
				ix->d_bits = static_cast<uint64_t *>(nbits[i].p);
			}
			nbits[i].p = nullptr;
		}
		t->words_per_bv = new_stride;
	}
	for (uint32_t i = 0; i < n_cols; i++) {
		Column &c = t->columns[cols[i].col_id];
		if (c.packed()) {
			// both forms were resident: the packed one does not cover the new rows — drop it (cudaFree waits for every
			// kernel that may still read it); cubit_gpu_pack_column brings it back whenever the caller wants
			cudaFree(c.d_words);
			cudaFree(c.d_hdr);
			c.d_words = nullptr;
			c.d_hdr = nullptr;
			c.packed_bytes = 0;
			c.pack_max_width = 0;
			c.pack_avg_width = 0;
		}
		if (ncol[i].p) {
			cudaFree(c.d);
			c.d = ncol[i].p;
			c.cap = ncol[i].cap;
		}
		if (nvalid[i].p) {
			cudaFree(c.d_valid);
			c.d_valid = static_cast<unsigned long long *>(nvalid[i].p);
			c.valid_cap_words = nvalid[i].cap;
		}
	}
	// ---- the new rows host → device behind the old ones; appended rows are valid until a new mask is uploaded
	for (uint32_t i = 0; i < n_cols; i++) {
		Column &c = t->columns[cols[i].col_id];
		CU_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(c.d) + (size_t)old_n * c.elem, cols[i].data, (size_t)n_new * c.elem,
		                       cudaMemcpyHostToDevice, st));
		c.n = new_n;
		if (c.d_valid) {
			if (old_n & 63) {
				unsigned long long last = 0;
				CU_TRY(cudaMemcpyAsync(&last, c.d_valid + old_w - 1, 8, cudaMemcpyDeviceToHost, st));
				CU_TRY(cudaStreamSynchronize(st));
				last |= ~0ull << (old_n & 63);
				CU_TRY(cudaMemcpyAsync(c.d_valid + old_w - 1, &last, 8, cudaMemcpyHostToDevice, st));
				CU_TRY(cudaStreamSynchronize(st));
			}
			if (new_w > old_w) {
				CU_TRY(cudaMemsetAsync(c.d_valid + old_w, 0xff, (new_w - old_w) * 8, st));
			}
		}
	}
	const uint32_t old_n_seg = t->n_seg;
	t->n_rows = new_n;
	t->n_seg = (uint32_t)new_n_seg;
	t->n_words = (new_n + 63) / 64;
	// ---- pending-delta CSRs are keyed by value * n_seg + segment: re-key them when the segment count changed
	if (t->n_seg != old_n_seg) {
		for (Index *ix : t->indexes) {
			int rc = delta_restride_locked(t, ix, t->n_seg);
			if (rc) {
				return rc;
			}
		}
	}
	// ---- indexes built from a column: index the new rows on the GPU
	for (Index *ix : t->indexes) {
		ix->counts_valid = false;
		if (ix->src_col < 0) {
			continue; // uploaded bitvectors: the new rows' bits are 0 until the caller uploads them
		}
		auto it = t->columns.find(ix->src_col);
		if (it == t->columns.end() || !it->second.d) {
			continue;
		}
		int launches = 0;
		if (!ix->compressed) {
			CU_TRY(launch_index_build(it->second.d, it->second.elem, it->second.d_valid, old_n, new_n, ix->src_base, ix->card,
			                          ix->d_bits, t->words_per_bv, t->sm_count, st, &launches));
			t->launches += launches;
			continue;
		}
		// compressed: containers are immutable, so the values are rebuilt from the source column in batches
		const uint64_t bv_bytes = t->words_per_bv * 8;
		const uint32_t nb = (uint32_t)std::min<uint64_t>(ix->card, std::max<uint64_t>(1, (2ull << 30) / bv_bytes));
		uint64_t *tmp = nullptr;
		CU_TRY(cudaMalloc(&tmp, (size_t)nb * bv_bytes));
		int rc = CUBIT_OK;
		for (uint32_t v0 = 0; v0 < ix->card && rc == CUBIT_OK; v0 += nb) {
			const uint32_t nv = std::min<uint32_t>(nb, ix->card - v0);
			cudaError_t be = cudaMemsetAsync(tmp, 0, (size_t)nv * bv_bytes, st);
			if (be == cudaSuccess) {
				be = launch_index_build(it->second.d, it->second.elem, it->second.d_valid, 0, new_n, ix->src_base + v0, nv, tmp,
				                        t->words_per_bv, t->sm_count, st, &launches);
			}
			t->launches += launches;
			rc = be == cudaSuccess ? compress_value_locked(t, ix, v0, nv, tmp)
			                       : fail(CUBIT_ECUDA, "append: index rebuild: %s", cudaGetErrorString(be));
		}
		cudaStreamSynchronize(st);
		cudaFree(tmp);
		if (rc) {
			return rc;
		}
	}
	CU_TRY(cudaStreamSynchronize(st));
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_drop_column(cubit_gpu_table *t, int32_t col_id) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		for (auto *s : t->shards) {
			int rc = cubit_gpu_drop_column(s, col_id);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	free_column(it->second);
	t->columns.erase(it);
	return CUBIT_OK;
	ABI_END
}
