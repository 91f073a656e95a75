// cubit_sharded.cu — one table over several B200s behind the same C-ABI handle.
//
// Reference analog of the unit being replaced: the row-group ranges RowGroupCollection::NextParallelScan hands to
// the scan threads (src/storage/table/row_group_collection.cpp:174-224) and TableFunction's MaxThreads
// (src/include/duckdb/function/table_function.hpp:45-67).  Here the parallel unit is a contiguous ROW RANGE,
// aligned to the segment size, owned by one GPU: bit r of every bitvector and row r of every column depend only
// on r, so every shard is an ordinary single-device table with row_base = its first global row, and nothing is
// exchanged while a query runs.  cubit_gpu_create_sharded returns a parent handle; every entry point fans out:
//   * uploads / synthetic columns / index builds / deltas are cut (or routed) by row range,
//   * cubit_gpu_query enqueues the query on EVERY shard before waiting for any (all GPUs run concurrently),
//   * COUNT / SUM come back as one pinned header per shard and are added exactly (128-bit) — the final "gather" of
//     north_star is 8 headers; row-ID lists stay on their GPUs and cubit_gpu_fetch walks the shards in row order
//     (shard order IS row order, so the concatenation is globally sorted), copying from several GPUs at once.
// One process drives all devices (what a DuckDB process does); bench.py's torchrun ranks use one single-device
// table each and reduce with cubit_gpu_result_add_limbs + NCCL instead.
#include "table.h"

#include <algorithm>
#include <cstring>

using namespace cubit;

extern "C" int cubit_gpu_create_sharded(const int *devices, uint32_t n_devices, uint64_t n_rows, int64_t row_base,
                                        uint32_t seg_bits, cubit_gpu_table **out) {
	ABI_BEGIN
	if (!out || !devices || n_devices == 0) {
		return fail(CUBIT_EINVAL, "NULL argument / no devices");
	}
	*out = nullptr;
	if (seg_bits != 32768 && seg_bits != 65536 && seg_bits != 131072) {
		return fail(CUBIT_EINVAL, "seg_bits must be 32768, 65536 or 131072 (got %u)", seg_bits);
	}
	if (n_rows == 0) {
		return fail(CUBIT_EINVAL, "n_rows must be > 0");
	}
	if (row_base < 0 || (uint64_t)row_base % seg_bits != 0) {
		return fail(CUBIT_EINVAL, "row_base of a sharded table must be a non-negative multiple of seg_bits");
	}
	// contiguous row ranges of whole segments, as even as the segment count allows
	const uint64_t n_seg = (n_rows + seg_bits - 1) / seg_bits;
	const uint64_t n_shards = std::min<uint64_t>(n_devices, n_seg);
	cubit_gpu_table *p = new cubit_gpu_table();
	p->device = devices[0];
	p->n_rows = n_rows;
	p->row_base = row_base;
	p->seg_bits = seg_bits;
	p->seg_words = seg_bits / 64;
	p->n_seg = (uint32_t)n_seg;
	p->n_words = (n_rows + 63) / 64;
	p->words_per_bv = n_seg * p->seg_words;
	uint64_t row0 = 0;
	for (uint64_t s = 0; s < n_shards; s++) {
		const uint64_t segs = n_seg / n_shards + (s < n_seg % n_shards ? 1 : 0);
		const uint64_t rows = std::min<uint64_t>(segs * seg_bits, n_rows - row0);
		cubit_gpu_table *c = nullptr;
		int rc = cubit_gpu_create(devices[s], rows, row_base + (int64_t)row0, seg_bits, &c);
		if (rc) {
			const std::string why = last_error_cstr();
			sharded_destroy(p);
			return fail(rc, "shard %llu on device %d: %s", (unsigned long long)s, devices[s], why.c_str());
		}
		p->shards.push_back(c);
		p->shard_row0.push_back(row0);
		row0 += rows;
	}
	p->shard_row0.push_back(n_rows);
	p->sm_count = p->shards[0]->sm_count;
	*out = p;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_shard_count(const cubit_gpu_table *t, uint32_t *n_shards) {
	if (!t || !n_shards) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*n_shards = t->sharded() ? (uint32_t)t->shards.size() : 1u;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_shard_info(const cubit_gpu_table *t, uint32_t shard, int *device, uint64_t *first_row,
                                    uint64_t *n_rows) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	const uint32_t n = t->sharded() ? (uint32_t)t->shards.size() : 1u;
	if (shard >= n) {
		return fail(CUBIT_EINVAL, "shard %u out of range (%u shards)", shard, n);
	}
	const cubit_gpu_table *c = t->sharded() ? t->shards[shard] : t;
	if (device) {
		*device = c->device;
	}
	if (first_row) {
		*first_row = t->sharded() ? t->shard_row0[shard] : 0;
	}
	if (n_rows) {
		*n_rows = c->n_rows;
	}
	return CUBIT_OK;
}

namespace cubit {

int sharded_destroy(cubit_gpu_table *t) {
	for (auto *s : t->shards) {
		cubit_gpu_destroy(s);
	}
	t->shards.clear();
	wire_pool_free(t);
	delete t;
	return CUBIT_OK;
}

int sharded_free_result(cubit_gpu_result *r) {
	for (auto *p : r->parts) {
		cubit_gpu_free_result(p);
	}
	delete r;
	return CUBIT_OK;
}

int sharded_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out) {
	cubit_gpu_result *r = new cubit_gpu_result();
	r->t = t;
	r->flags = q->flags;
	r->agg_kind = q->agg_kind;
	r->n_cols = (q->flags & CUBIT_Q_VALUES) ? q->n_cols : 0;
	cubit_query cq = *q;
	cq.flags |= CUBIT_Q_ASYNC; // enqueue on every GPU first, wait afterwards
	for (auto *s : t->shards) {
		cubit_gpu_result *part = nullptr;
		int rc = cubit_gpu_query(s, &cq, &part);
		if (rc) {
			const std::string why = last_error_cstr();
			sharded_free_result(r);
			return fail(rc, "%s", why.c_str());
		}
		r->parts.push_back(part);
	}
	if (!(q->flags & CUBIT_Q_ASYNC)) {
		int rc = sharded_result_finish(r);
		if (rc) {
			const std::string why = last_error_cstr();
			sharded_free_result(r);
			return fail(rc, "%s", why.c_str());
		}
	}
	*out = r;
	return CUBIT_OK;
}

int sharded_result_finish(cubit_gpu_result *r) {
	if (r->finished.load(std::memory_order_acquire)) {
		return r->fin_rc ? fail(r->fin_rc, "%s", r->fin_err.c_str()) : CUBIT_OK;
	}
	std::lock_guard<std::mutex> lk(r->fin_mu);
	if (r->finished.load(std::memory_order_acquire)) {
		return r->fin_rc ? fail(r->fin_rc, "%s", r->fin_err.c_str()) : CUBIT_OK;
	}
	cubit_result_info tot;
	memset(&tot, 0, sizeof(tot));
	int rc = CUBIT_OK;
	r->count_prefix.assign(r->parts.size() + 1, 0);
	for (size_t i = 0; i < r->parts.size(); i++) {
		cubit_result_info pi;
		const int prc = cubit_gpu_result_get(r->parts[i], &pi);
		if (prc) {
			if (rc == CUBIT_OK) {
				rc = prc;
				r->fin_err = last_error_cstr();
			}
			continue; // every shard is waited for, whatever the first error was
		}
		r->count_prefix[i + 1] = r->count_prefix[i] + pi.count;
		tot.count += pi.count;
		const uint64_t lo = tot.sum_lo + pi.sum_lo; // exact 128-bit add of the shard sums
		tot.sum_hi += pi.sum_hi + (lo < tot.sum_lo ? 1 : 0);
		tot.sum_lo = lo;
		tot.sum_f64 += pi.sum_f64;
		tot.agg_rows += pi.agg_rows;
		tot.capacity += pi.capacity;
		tot.n_streams = pi.n_streams;
		tot.n_launches += pi.n_launches;
		tot.delta_entries += pi.delta_entries;
		tot.algo_bytes_scan += pi.algo_bytes_scan;
		tot.algo_bytes_probe += pi.algo_bytes_probe;
		tot.ms_scan = std::max(tot.ms_scan, pi.ms_scan); // the shards run concurrently: the slowest one is the query
		tot.ms_probe = std::max(tot.ms_probe, pi.ms_probe);
		tot.ms_total = std::max(tot.ms_total, pi.ms_total);
		tot.fused = pi.fused;
		tot.probe_path = pi.probe_path;
		tot.scan_path = pi.scan_path;
	}
	r->info = tot; // device pointers stay NULL: the rows live on several devices (cubit_gpu_fetch walks them)
	r->fin_rc = rc;
	r->finished.store(true, std::memory_order_release);
	return rc ? fail(rc, "%s", r->fin_err.c_str()) : CUBIT_OK;
}

int sharded_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                  void *const *host_cols, cubit_gpu_fetch_ticket **ticket) {
	if (offset > r->info.count || n > r->info.count - offset) {
		return fail(CUBIT_EINVAL, "fetch range [%llu, +%llu) outside result of %llu rows", (unsigned long long)offset,
		            (unsigned long long)n, (unsigned long long)r->info.count);
	}
	if (n_cols > r->n_cols || n_cols > CUBIT_MAX_PROBE_COLS) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, %u requested", r->n_cols, n_cols);
	}
	// every overlapping shard copies its piece concurrently (one copy engine per GPU); the caller waits for all
	cubit_gpu_fetch_ticket *tk = new cubit_gpu_fetch_ticket();
	int rc = CUBIT_OK;
	for (size_t i = 0; i < r->parts.size() && rc == CUBIT_OK && n; i++) {
		const uint64_t p0 = r->count_prefix[i], p1 = r->count_prefix[i + 1];
		const uint64_t lo = std::max(offset, p0), hi = std::min(offset + n, p1);
		if (lo >= hi) {
			continue;
		}
		void *cols[CUBIT_MAX_PROBE_COLS] = {};
		for (uint32_t c = 0; c < n_cols; c++) {
			if (host_cols && host_cols[c]) {
				cols[c] = static_cast<char *>(host_cols[c]) + (lo - offset) * r->parts[i]->val_elem[c];
			}
		}
		cubit_gpu_fetch_ticket *pt = nullptr;
		rc = cubit_gpu_fetch_async(r->parts[i], lo - p0, hi - lo, host_rowids ? host_rowids + (lo - offset) : nullptr, n_cols,
		                           cols, &pt);
		if (rc == CUBIT_OK) {
			tk->parts.push_back(pt);
		}
	}
	if (rc || !ticket) {
		const std::string why = rc ? last_error_cstr() : "";
		const int wrc = cubit_gpu_fetch_wait(tk);
		return rc ? fail(rc, "%s", why.c_str()) : wrc;
	}
	*ticket = tk;
	return CUBIT_OK;
}

int sharded_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n, uint64_t *host_words,
                           int *all_valid) {
	if (col >= r->n_cols) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, column %u requested", r->n_cols, col);
	}
	if (offset > r->info.count || n > r->info.count - offset) {
		return fail(CUBIT_EINVAL, "fetch range [%llu, +%llu) outside result of %llu rows", (unsigned long long)offset,
		            (unsigned long long)n, (unsigned long long)r->info.count);
	}
	const uint64_t out_words = (n + 63) / 64;
	for (uint64_t w = 0; host_words && w < out_words; w++) {
		host_words[w] = 0;
	}
	bool all = true;
	std::vector<uint64_t> tmp;
	for (size_t i = 0; i < r->parts.size() && n; i++) {
		const uint64_t p0 = r->count_prefix[i], p1 = r->count_prefix[i + 1];
		const uint64_t lo = std::max(offset, p0), hi = std::min(offset + n, p1);
		if (lo >= hi) {
			continue;
		}
		const uint64_t m = hi - lo;
		tmp.assign((m + 63) / 64 + 1, 0);
		int av = 1;
		int rc = cubit_gpu_fetch_validity(r->parts[i], col, lo - p0, m, tmp.data(), &av);
		if (rc) {
			return rc;
		}
		all &= av != 0;
		if (host_words) { // append m bits at bit position (lo - offset)
			const uint64_t at = lo - offset, w0 = at / 64, sh = at % 64;
			for (uint64_t w = 0; w < (m + 63) / 64; w++) {
				host_words[w0 + w] |= tmp[w] << sh;
				if (sh && w0 + w + 1 < out_words) {
					host_words[w0 + w + 1] |= tmp[w] >> (64 - sh);
				}
			}
		}
	}
	if (all_valid) {
		*all_valid = all ? 1 : 0;
	}
	return CUBIT_OK;
}

int sharded_fetch_bitvector(cubit_gpu_result *r, uint64_t *host_words, uint64_t n_words) {
	cubit_gpu_table *t = r->t;
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "n_words mismatch");
	}
	for (size_t i = 0; i < r->parts.size(); i++) {
		int rc = cubit_gpu_fetch_bitvector(r->parts[i], host_words + t->shard_row0[i] / 64, t->shards[i]->n_words);
		if (rc) {
			return rc;
		}
	}
	return CUBIT_OK;
}

int sharded_probe(cubit_gpu_table *t, int32_t col_id, const int64_t *host_rowids, uint64_t n, void *host_out,
                  uint64_t *sum_lo, int64_t *sum_hi) {
	// row IDs are sorted (the contract of the scan), so every shard's share is one contiguous run
	uint64_t lo_sum = 0;
	int64_t hi_sum = 0;
	uint64_t at = 0;
	uint32_t elem = 0;
	{
		cubit_gpu_table *s0 = t->shards[0];
		std::shared_lock<std::shared_mutex> lk(s0->mu);
		auto it = s0->columns.find(col_id);
		if (it == s0->columns.end()) {
			return fail(CUBIT_EINVAL, "no column %d", col_id);
		}
		elem = it->second.elem;
	}
	for (size_t s = 0; s < t->shards.size(); s++) {
		const int64_t end = t->row_base + (int64_t)t->shard_row0[s + 1];
		uint64_t e = at;
		while (e < n && host_rowids[e] < end) {
			if (e > at && host_rowids[e] < host_rowids[e - 1]) {
				return fail(CUBIT_EINVAL, "row ids must be ascending");
			}
			e++;
		}
		if (e > at) {
			uint64_t l = 0;
			int64_t h = 0;
			int rc = cubit_gpu_probe(t->shards[s], col_id, host_rowids + at, e - at,
			                         host_out ? static_cast<char *>(host_out) + at * elem : nullptr, (sum_lo || sum_hi) ? &l : nullptr,
			                         (sum_lo || sum_hi) ? &h : nullptr);
			if (rc) {
				return rc;
			}
			const uint64_t nl = lo_sum + l;
			hi_sum += h + (nl < lo_sum ? 1 : 0);
			lo_sum = nl;
		}
		at = e;
	}
	if (at != n) {
		return fail(CUBIT_EINVAL, "row id %lld outside this table", (long long)host_rowids[at]);
	}
	if (sum_lo) {
		*sum_lo = lo_sum;
	}
	if (sum_hi) {
		*sum_hi = hi_sum;
	}
	return CUBIT_OK;
}

} // namespace cubit
