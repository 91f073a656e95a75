// cubit_delta.cu — pending update/delete deltas of a CUBIT index: incremental, device-side ingestion
// (cubit_gpu_add_delta / _add_delta_pairs / _set_delta), threshold-driven and manual merge-back.
//
// Reference analog: BoundIndex::Append / Delete / Insert / MergeIndexes
// (src/include/duckdb/execution/index/bound_index.hpp:71-97) — the calls DuckDB's DML path makes on an index.
// An UPDATE v→w of row r flips r in D_v and D_w, a DELETE of a row holding v flips r in D_v;
// scans XOR the lists into the staged segments at query time.  Nothing here sorts, deduplicates or waits for the
// GPU: the flipped (value, row) pairs are copied to the device and merged into the index's delta CSR by the
// kernels of delta_kernels.cu on the table's kernel stream.
#include "table.h"

#include <algorithm>
#include <cstring>

using namespace cubit;

namespace {

uint32_t seg_shift_of(const cubit_gpu_table *t) {
	return t->seg_bits == 32768 ? 15u : (t->seg_bits == 65536 ? 16u : 17u);
}

// merge `n` (value, row) pairs into the index's CSR; drop_value: forget that value's old entries first
int ingest_locked(cubit_gpu_table *t, Index *ix, const uint32_t *values, uint32_t one_value, const int64_t *rows, uint64_t n,
                  uint32_t drop_value) {
	DeltaSet &d = ix->delta;
	{
		int src = delta_settle_locked(t, ix); // counts of an earlier add_delta_pairs (and its merge-back decision)
		if (src) {
			return src;
		}
	}
	const uint64_t n_keys = (uint64_t)ix->card * t->n_seg;
	if (n_keys > (1ull << 30)) {
		return fail(CUBIT_ESTATE, "pending deltas need cardinality * segments <= 2^30 (have %llu)", (unsigned long long)n_keys);
	}
	if (d.d_off && d.n_seg != t->n_seg) {
		int rc = delta_restride_locked(t, ix, t->n_seg);
		if (rc) {
			return rc;
		}
	}
	const uint64_t dropped = drop_value != 0xffffffffu ? d.rows[drop_value] : 0;
	const uint64_t n_old = d.d_off ? d.n_ent : 0;
	const uint64_t n_total = n_old - dropped + n;
	if (n_total > 0xfffffff0ull) {
		return fail(CUBIT_EINVAL, "too many pending delta entries: merge them first");
	}
	if (n_total == 0) { // nothing pending any more
		CU_TRY(cudaStreamSynchronize(t->stream));
		free_delta(d);
		return CUBIT_OK;
	}
	cudaStream_t st = t->stream;
	long long *d_rows = nullptr;
	uint32_t *d_vals = nullptr, *d_cnt = nullptr, *d_bsum = nullptr, *new_off = nullptr;
	DeltaEnt *new_ent = nullptr;
	const uint64_t n_blocks = (n_keys + 4095) / 4096;
	cudaError_t e = cudaSuccess;
	auto try_ = [&](cudaError_t x) {
		if (e == cudaSuccess) {
			e = x;
		}
	};
	if (n) {
		try_(cudaMallocAsync((void **)&d_rows, n * 8, st));
		if (values) {
			try_(cudaMallocAsync((void **)&d_vals, n * 4, st));
		}
		try_(cudaMallocAsync((void **)&d_cnt, n_keys * 4, st));
	}
	try_(cudaMallocAsync((void **)&d_bsum, (n_blocks + 1) * 4, st));
	try_(cudaMallocAsync((void **)&new_off, (n_keys + 1) * 4, st));
	try_(cudaMallocAsync((void **)&new_ent, n_total * sizeof(DeltaEnt), st));
	if (e == cudaSuccess && n) {
		// pageable source: the copy is staged by the driver and the caller may reuse its arrays on return
		try_(cudaMemcpyAsync(d_rows, rows, n * 8, cudaMemcpyHostToDevice, st));
		if (values) {
			try_(cudaMemcpyAsync(d_vals, values, n * 4, cudaMemcpyHostToDevice, st));
		}
		try_(cudaMemsetAsync(d_cnt, 0, n_keys * 4, st));
	}
	if (e == cudaSuccess) {
		DeltaIngest a;
		memset(&a, 0, sizeof(a));
		a.rows = d_rows;
		a.values = d_vals;
		a.one_value = one_value;
		a.n_new = n;
		a.n_seg = t->n_seg;
		a.seg_shift = seg_shift_of(t);
		a.n_keys = n_keys;
		a.old_off = d.d_off;
		a.old_ent = d.d_ent;
		a.n_old = n_old;
		a.drop_value = drop_value;
		a.cnt = d_cnt;
		a.block_sum = d_bsum;
		a.new_off = new_off;
		a.new_ent = new_ent;
		int launches = 0;
		try_(launch_delta_ingest(a, t->sm_count, st, &launches));
		t->launches += launches;
	}
	// scratch and the OLD lists go back to the pool in stream order: scans enqueued earlier still read them
	for (void *p : {(void *)d_rows, (void *)d_vals, (void *)d_cnt, (void *)d_bsum}) {
		if (p) {
			cudaFreeAsync(p, st);
		}
	}
	if (e != cudaSuccess) {
		if (new_off) {
			cudaFreeAsync(new_off, st);
		}
		if (new_ent) {
			cudaFreeAsync(new_ent, st);
		}
		return fail(e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "delta ingestion: %s", cudaGetErrorString(e));
	}
	if (d.d_off) {
		cudaFreeAsync(d.d_off, st);
	}
	if (d.d_ent) {
		cudaFreeAsync(d.d_ent, st);
	}
	d.d_off = new_off;
	d.d_ent = new_ent;
	d.n_ent = n_total;
	d.cap_ent = n_total;
	d.n_seg = t->n_seg;
	if (values) { // per-value counts come back from the device (no 10 M-iteration host loop): see delta_settle_locked
		if (!d.h_voff) {
			CU_TRY(cudaMallocHost((void **)&d.h_voff, ((size_t)ix->card + 1) * 4));
			CU_TRY(cudaEventCreateWithFlags(&d.ev_voff, cudaEventDisableTiming));
		}
		uint32_t *d_voff = nullptr;
		CU_TRY(cudaMallocAsync((void **)&d_voff, ((size_t)ix->card + 1) * 4, st));
		CU_TRY(launch_delta_value_offsets(new_off, t->n_seg, ix->card, d_voff, st));
		t->launches++;
		CU_TRY(cudaMemcpyAsync(d.h_voff, d_voff, ((size_t)ix->card + 1) * 4, cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaEventRecord(d.ev_voff, st));
		cudaFreeAsync(d_voff, st);
		d.voff_pending = true;
	}
	return CUBIT_OK;
}

// the merge-back rule ("compaction of D_i into B_i past a threshold"): once the pending entries of
// any touched value outweigh merge_fraction of its bitvector, scans would read more delta than data — fold them in
int maybe_auto_merge(cubit_gpu_table *t, Index *ix, const std::vector<uint32_t> &touched) {
	if (ix->merge_fraction <= 0 || ix->delta.n_ent == 0) {
		return CUBIT_OK;
	}
	const double limit = ix->merge_fraction * (double)t->n_words * 8.0;
	const uint32_t n_check = touched.empty() ? ix->card : (uint32_t)touched.size();
	for (uint32_t i = 0; i < n_check; i++) {
		const uint32_t v = touched.empty() ? i : touched[i];
		if ((double)ix->delta.rows[v] * sizeof(DeltaEnt) > limit) {
			bool any = false;
			int rc = merge_deltas_locked(t, ix, &any);
			if (rc == CUBIT_OK && any) {
				ix->auto_merges++;
			}
			return rc;
		}
	}
	return CUBIT_OK;
}

int check_rows(const cubit_gpu_table *t, const int64_t *rows, uint64_t n) {
	uint64_t bad = 0;
	const uint64_t lim = t->n_rows;
	for (uint64_t i = 0; i < n; i++) {
		bad |= (uint64_t)rows[i] >= lim; // negative rows wrap to huge values
	}
	if (bad) {
		return fail(CUBIT_EINVAL, "delta row out of range [0, %llu)", (unsigned long long)t->n_rows);
	}
	return CUBIT_OK;
}

// sharded parent: route the (value, row) pairs to the shard owning each row, rows re-based to the shard
int sharded_add(cubit_gpu_table *t, int32_t index_id, const uint32_t *values, uint32_t one_value, const int64_t *rows,
                uint64_t n, bool replace) {
	const size_t ns = t->shards.size();
	std::vector<std::vector<int64_t>> srows(ns);
	std::vector<std::vector<uint32_t>> svals(ns);
	for (uint64_t i = 0; i < n; i++) {
		const uint64_t r = (uint64_t)rows[i];
		if (r >= t->n_rows) {
			return fail(CUBIT_EINVAL, "delta row out of range [0, %llu)", (unsigned long long)t->n_rows);
		}
		size_t s = std::upper_bound(t->shard_row0.begin(), t->shard_row0.end(), r) - t->shard_row0.begin() - 1;
		srows[s].push_back((int64_t)(r - t->shard_row0[s]));
		if (values) {
			svals[s].push_back(values[i]);
		}
	}
	for (size_t s = 0; s < ns; s++) {
		int rc;
		if (replace) {
			rc = cubit_gpu_set_delta(t->shards[s], index_id, one_value, srows[s].data(), srows[s].size());
		} else if (srows[s].empty()) {
			continue;
		} else if (values) {
			rc = cubit_gpu_add_delta_pairs(t->shards[s], index_id, svals[s].data(), srows[s].data(), srows[s].size());
		} else {
			rc = cubit_gpu_add_delta(t->shards[s], index_id, one_value, srows[s].data(), srows[s].size());
		}
		if (rc) {
			return rc;
		}
	}
	return CUBIT_OK;
}

} // namespace

namespace cubit {

int delta_settle_locked(cubit_gpu_table *t, Index *ix) {
	DeltaSet &d = ix->delta;
	if (!d.voff_pending) {
		return CUBIT_OK;
	}
	CU_TRY(cudaEventSynchronize(d.ev_voff));
	d.voff_pending = false;
	for (uint32_t v = 0; v < ix->card; v++) {
		d.rows[v] = d.h_voff[v + 1] - d.h_voff[v];
	}
	return maybe_auto_merge(t, ix, {});
}

int merge_deltas_locked(cubit_gpu_table *t, Index *ix, bool *any) {
	DeltaSet &d = ix->delta;
	if (any) {
		*any = false;
	}
	if (d.voff_pending) { // (the counts are about to be zeroed: only the event has to be retired)
		CU_TRY(cudaEventSynchronize(d.ev_voff));
		d.voff_pending = false;
	}
	if (!d.d_off || d.n_ent == 0) {
		return CUBIT_OK;
	}
	if (d.n_seg != t->n_seg) {
		int rc = delta_restride_locked(t, ix, t->n_seg);
		if (rc) {
			return rc;
		}
	}
	if (!ix->compressed) {
		CU_TRY(launch_delta_apply(d.d_ent, 0, d.n_ent, t->n_seg, t->seg_words, ix->d_bits, t->words_per_bv, 0, t->sm_count,
		                          t->stream));
		t->launches++;
	} else {
		// containers are immutable: expand every value that has pending rows, XOR its entries in, store it again
		std::vector<uint32_t> off(ix->card + 1);
		std::vector<uint32_t> all((size_t)ix->card * t->n_seg + 1);
		CU_TRY(cudaMemcpyAsync(all.data(), d.d_off, all.size() * 4, cudaMemcpyDeviceToHost, t->stream));
		CU_TRY(cudaStreamSynchronize(t->stream));
		uint64_t *tmp = nullptr;
		CU_TRY(cudaMalloc(&tmp, t->words_per_bv * 8));
		int rc = CUBIT_OK;
		for (uint32_t v = 0; v < ix->card && rc == CUBIT_OK; v++) {
			const uint64_t e0 = all[(size_t)v * t->n_seg], e1 = all[(size_t)(v + 1) * t->n_seg];
			if (e1 == e0) {
				continue;
			}
			rc = expand_value_locked(t, ix, v, tmp);
			if (rc == CUBIT_OK) {
				cudaError_t e = launch_delta_apply(d.d_ent, e0, e1, t->n_seg, t->seg_words, tmp, t->words_per_bv, v, t->sm_count,
				                                   t->stream);
				t->launches++;
				rc = e == cudaSuccess ? compress_value_locked(t, ix, v, 1, tmp)
				                      : fail(CUBIT_ECUDA, "delta apply: %s", cudaGetErrorString(e));
			}
		}
		cudaStreamSynchronize(t->stream);
		cudaFree(tmp);
		if (rc) {
			return rc;
		}
	}
	// the lists are released in stream order: the apply kernel (and earlier scans) still read them
	cudaFreeAsync(d.d_off, t->stream);
	cudaFreeAsync(d.d_ent, t->stream);
	d.d_off = nullptr;
	d.d_ent = nullptr;
	d.n_ent = d.cap_ent = 0;
	std::fill(d.rows.begin(), d.rows.end(), 0);
	ix->counts_valid = false;
	if (any) {
		*any = true;
	}
	return CUBIT_OK;
}

int delta_rows_locked(cubit_gpu_table *t, Index *ix, uint32_t v, std::vector<int64_t> &rows) {
	rows.clear();
	DeltaSet &d = ix->delta;
	int src = delta_settle_locked(t, ix);
	if (src) {
		return src;
	}
	if (!d.d_off || d.rows[v] == 0) {
		return CUBIT_OK;
	}
	if (d.n_seg != t->n_seg) {
		int rc = delta_restride_locked(t, ix, t->n_seg);
		if (rc) {
			return rc;
		}
	}
	std::vector<uint32_t> off((size_t)t->n_seg + 1);
	CU_TRY(cudaMemcpyAsync(off.data(), d.d_off + (size_t)v * t->n_seg, off.size() * 4, cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	const uint64_t e0 = off.front(), e1 = off.back();
	std::vector<DeltaEnt> ent(e1 - e0);
	if (e1 > e0) {
		CU_TRY(cudaMemcpyAsync(ent.data(), d.d_ent + e0, (e1 - e0) * sizeof(DeltaEnt), cudaMemcpyDeviceToHost, t->stream));
		CU_TRY(cudaStreamSynchronize(t->stream));
	}
	for (uint32_t sgm = 0; sgm < t->n_seg; sgm++) {
		for (uint64_t e = off[sgm]; e < off[sgm + 1]; e++) {
			const DeltaEnt &x = ent[e - e0];
			const uint64_t row0 = ((uint64_t)sgm * t->seg_words + x.word) * 64;
			for (uint64_t m = x.mask; m; m &= m - 1) {
				rows.push_back((int64_t)(row0 + (uint64_t)__builtin_ctzll(m)));
			}
		}
	}
	// the list may name a row twice (it cancels): report the net flips, ascending
	std::sort(rows.begin(), rows.end());
	std::vector<int64_t> net;
	for (size_t i = 0; i < rows.size();) {
		size_t j = i;
		while (j < rows.size() && rows[j] == rows[i]) {
			j++;
		}
		if ((j - i) & 1) {
			net.push_back(rows[i]);
		}
		i = j;
	}
	rows.swap(net);
	return CUBIT_OK;
}

int delta_restride_locked(cubit_gpu_table *t, Index *ix, uint32_t new_n_seg) {
	DeltaSet &d = ix->delta;
	if (!d.d_off || d.n_seg == new_n_seg) {
		d.n_seg = new_n_seg;
		return CUBIT_OK;
	}
	if ((uint64_t)ix->card * new_n_seg > (1ull << 30)) {
		return fail(CUBIT_ESTATE, "pending deltas need cardinality * segments <= 2^30");
	}
	uint32_t *new_off = nullptr;
	CU_TRY(cudaMallocAsync((void **)&new_off, ((uint64_t)ix->card * new_n_seg + 1) * 4, t->stream));
	CU_TRY(launch_delta_restride(d.d_off, new_off, d.d_ent, d.n_ent, ix->card, d.n_seg, new_n_seg, t->sm_count, t->stream));
	t->launches += 2;
	cudaFreeAsync(d.d_off, t->stream);
	d.d_off = new_off;
	d.n_seg = new_n_seg;
	return CUBIT_OK;
}

} // namespace cubit

extern "C" int cubit_gpu_add_delta(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const int64_t *rows,
                                   uint64_t n) {
	ABI_BEGIN
	if (!t || (n && !rows)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n == 0) {
		return CUBIT_OK;
	}
	if (t->sharded()) {
		return sharded_add(t, index_id, nullptr, value_id, rows, n, false);
	}
	int rc = check_rows(t, rows, n);
	if (rc) {
		return rc;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	rc = ingest_locked(t, ix, nullptr, value_id, rows, n, 0xffffffffu);
	if (rc) {
		return rc;
	}
	ix->delta.rows[value_id] += n;
	return maybe_auto_merge(t, ix, {value_id});
	ABI_END
}

extern "C" int cubit_gpu_add_delta_pairs(cubit_gpu_table *t, int32_t index_id, const uint32_t *value_ids,
                                         const int64_t *rows, uint64_t n) {
	ABI_BEGIN
	if (!t || (n && (!rows || !value_ids))) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n == 0) {
		return CUBIT_OK;
	}
	if (t->sharded()) {
		return sharded_add(t, index_id, value_ids, 0, rows, n, false);
	}
	int rc = check_rows(t, rows, n);
	if (rc) {
		return rc;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	uint32_t vmax = 0;
	for (uint64_t i = 0; i < n; i++) {
		vmax = std::max(vmax, value_ids[i]);
	}
	if (vmax >= ix->card) {
		return fail(CUBIT_EINVAL, "a pair names value %u outside the index (cardinality %u)", vmax, ix->card);
	}
	// the per-value counts (planning bounds, merge-back rule) come back from the device: delta_settle_locked
	return ingest_locked(t, ix, value_ids, 0, rows, n, 0xffffffffu);
	ABI_END
}

// replace the pending delta of one bitvector (a row listed twice cancels); never triggers the automatic merge-back —
// the caller states exactly what stays pending
extern "C" int cubit_gpu_set_delta(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const int64_t *rows,
                                   uint64_t n) {
	ABI_BEGIN
	if (!t || (n && !rows)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		return sharded_add(t, index_id, nullptr, value_id, rows, n, true);
	}
	int rc = check_rows(t, rows, n);
	if (rc) {
		return rc;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	rc = delta_settle_locked(t, ix);
	if (rc) {
		return rc;
	}
	if (n == 0 && ix->delta.rows[value_id] == 0) {
		return CUBIT_OK;
	}
	rc = ingest_locked(t, ix, nullptr, value_id, rows, n, value_id);
	if (rc) {
		return rc;
	}
	ix->delta.rows[value_id] = n;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_merge_deltas(cubit_gpu_table *t, int32_t index_id) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		for (auto *s : t->shards) {
			int rc = cubit_gpu_merge_deltas(s, index_id);
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
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	return merge_deltas_locked(t, ix, nullptr);
	ABI_END
}

extern "C" int cubit_gpu_set_merge_threshold(cubit_gpu_table *t, int32_t index_id, double fraction) {
	ABI_BEGIN
	if (!t || !(fraction >= 0)) {
		return fail(CUBIT_EINVAL, "bad argument");
	}
	if (t->sharded()) {
		for (auto *s : t->shards) {
			int rc = cubit_gpu_set_merge_threshold(s, index_id, fraction);
			if (rc) {
				return rc;
			}
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	ix->merge_fraction = fraction;
	return CUBIT_OK;
	ABI_END
}
