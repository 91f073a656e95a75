// table.h — internal object model shared by the host-side translation units of libcubit_gpu
// (cubit_gpu.cu, cubit_columns.cu, cubit_delta.cu, cubit_persist.cu, cubit_query.cu, cubit_sharded.cu).
// Not part of the public ABI (include/cubit_gpu.h).
//
// Concurrency model of one table shard
//   * `stream` is the shard's in-order KERNEL stream: every kernel that reads or writes table state (value
//     bitvectors, pending-delta lists, columns) is enqueued on it, so maintenance (add_delta, merge-back, append)
//     is ordered against the scans around it by the stream itself and needs no host synchronisation.
//   * `mu` guards the host-side metadata and is held only while a call plans and ENQUEUES its work — never while
//     it waits for the GPU.  Several host threads can therefore have queries in flight on one table; their kernels
//     run back to back on the kernel stream while the callers overlap planning, waiting and DataChunk hand-off.
//     (The scan kernel is a persistent, co-resident grid — two such grids interleaved on one GPU could starve each
//     other's look-back — so kernels of one shard are deliberately NOT spread over several streams.)
//   * result rows leave the device on the shard's COPY streams (cubit_gpu_fetch / _fetch_async): a copy waits for
//     its query's completion event, not for the kernel stream, so window i+1 of one result, or the rows of another
//     query, cross PCIe while the kernel stream is already running the next scan.
//   * anything that frees or re-allocates device memory a kernel may still read synchronises the kernel stream
//     first (re-stride on append, drop/replace column, destroy).
#pragma once
#include "../../include/cubit_gpu.h"
#include "kernels.h"

#include <atomic>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

namespace cubit {

// ------------------------------------------------------------------- errors
int fail(int code, const char *fmt, ...); // sets the thread-local message, returns `code`
const char *last_error_cstr();

#define CU_TRY(expr)                                                                                                   \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			return ::cubit::fail(_e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "%s: %s (%s:%d)",       \
			                     #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
		}                                                                                                              \
	} while (0)

// No exception crosses the ABI (include/cubit_gpu.h): every extern "C" body that can allocate on the host sits
// between these two.
#define ABI_BEGIN try {
#define ABI_END                                                                                                        \
	}                                                                                                                  \
	catch (const std::bad_alloc &) {                                                                                   \
		return ::cubit::fail(CUBIT_ENOMEM, "host allocation failed");                                                  \
	}                                                                                                                  \
	catch (const std::exception &ex) {                                                                                 \
		return ::cubit::fail(CUBIT_EINVAL, "internal error: %s", ex.what());                                           \
	}                                                                                                                  \
	catch (...) {                                                                                                      \
		return ::cubit::fail(CUBIT_EINVAL, "internal error");                                                          \
	}

// ------------------------------------------------------------------ objects
// Pending update/delete deltas of ONE index: a CSR over keys (value_id * n_seg + segment) of 16-byte entries
// (word-in-segment, key, 64-bit mask).  Entries of one key are contiguous but unordered and MAY repeat a word:
// the scan kernel XORs them into the staged segment with shared-memory atomics, so a row listed twice cancels
// by itself and ingestion never has to sort or deduplicate.  The lists are rebuilt ON THE DEVICE by
// cubit_gpu_add_delta (histogram → scan → move old → scatter new, all on the kernel stream, no host sync) into
// the other half of a double buffer.
struct DeltaSet {
	uint32_t *d_off = nullptr; // [card * n_seg + 1] of the live list
	DeltaEnt *d_ent = nullptr; // live entries
	uint64_t n_ent = 0;        // live entries (host-tracked: ingestion is exact, nothing is dropped on the device)
	uint64_t cap_ent = 0;      // entries allocated in d_ent
	uint32_t n_seg = 0;        // the segment count the keys were computed with
	std::vector<uint64_t> rows; // [card] pending flipped rows per value (upper bound of the rows they can add)
	// add_delta_pairs does not count its pairs per value on the host: the per-value entry offsets of the new CSR are
	// copied back (card + 1 words, page-locked) behind the ingestion and folded into `rows` by delta_settle_locked
	uint32_t *h_voff = nullptr;
	cudaEvent_t ev_voff = nullptr;
	bool voff_pending = false;
};

// Roaring-style compressed storage of an index: one container per (value, segment) —
// EMPTY, FULL, ARRAY (sorted 16-bit positions, staged next to the ring and expanded in shared memory by the scan
// kernel) or BITMAP (the verbatim segment, bulk-copied as usual).  dir[value * n_seg_cap + segment] =
// type (2 bits) | count (14 bits: ARRAY entries) | pool offset in 16-byte units (48 bits).
struct CompressedStore {
	unsigned long long *d_dir = nullptr;
	uint8_t *d_pool = nullptr;
	uint64_t pool_bytes = 0, pool_cap = 0;
	uint64_t n_seg_cap = 0; // directory stride (segments per value allocated)
	uint64_t garbage_bytes = 0; // pool bytes no directory entry points at any more (re-compressed values)
};

struct Index {
	uint32_t card = 0;
	uint64_t *d_bits = nullptr;   // [card][words_per_bv] verbatim, or nullptr when the index is compressed
	bool compressed = false;
	CompressedStore cs;
	std::vector<uint64_t> counts; // popcount of every B_v as stored
	bool counts_valid = false;
	int32_t src_col = -1;         // column the index was built from (cubit_gpu_index_build), -1 = uploaded
	int64_t src_base = 0;
	DeltaSet delta;
	double merge_fraction = 0.25; // auto merge-back once a value's pending delta bytes exceed this share of its bitvector
	uint64_t auto_merges = 0;
};

struct Column {
	void *d = nullptr; // raw array (may be dropped once packed)
	uint32_t elem = 0;
	uint64_t n = 0;
	uint64_t cap = 0; // rows allocated (>= n; grows on append)
	// FOR-bit-packed form of an 8-byte column (kernels.h: ColRef)
	unsigned long long *d_words = nullptr;
	PackHdr *d_hdr = nullptr;
	uint64_t packed_bytes = 0;
	uint32_t pack_max_width = 0; // widest block of the packed form (bits)
	double pack_avg_width = 0;   // mean bits per value of the packed form
	// validity mask (ValidityMask layout, bit = 1: valid), nullptr = the column holds no NULLs
	unsigned long long *d_valid = nullptr;
	uint64_t valid_cap_words = 0;
	bool packed() const {
		return d_words != nullptr;
	}
};

constexpr int kCopyStreams = 4; // hand-off streams: the gaps between one window's kernels / copies are filled by another's
constexpr int kAggStreams = 4; // streams for queries whose kernels have no inter-CTA dependency (see plan_and_launch)
constexpr uint64_t kStageChunk = 16ull << 20;

} // namespace cubit

struct cubit_gpu_table {
	int device = 0;
	int sm_count = 0;
	uint64_t n_rows = 0;
	int64_t row_base = 0;
	uint32_t seg_bits = 0, seg_words = 0, n_seg = 0;
	uint64_t n_words = 0;      // ceil(n_rows / 64)
	uint64_t words_per_bv = 0; // segments allocated per bitvector * seg_words
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr; // kernel stream (own_stream unless cubit_gpu_set_stream)
	cudaStream_t copy_stream[cubit::kCopyStreams] = {};
	std::atomic<uint32_t> next_copy {0};
	// Aggregate-only / bitvector-only queries (no row positions → no look-back between CTAs, so their grids need not
	// be co-resident) run on a small pool of streams and overlap ON THE GPU: the config-1 query is a 92-CTA kernel.
	// Ordering against maintenance on the kernel stream: a pool stream waits for mut_event (recorded at the end of
	// every mutating section), and a mutating section first orders the kernel stream behind agg_last[] (TableLock).
	cudaStream_t agg_stream[cubit::kAggStreams] = {};
	cudaEvent_t agg_last[cubit::kAggStreams] = {};
	bool agg_used[cubit::kAggStreams] = {};
	cudaEvent_t mut_event = nullptr;
	bool mut_recorded = false;
	std::atomic<uint32_t> next_agg {0};
	std::vector<cudaEvent_t> ev_pool; // completion events, recycled across queries
	std::vector<cubit::Index *> indexes;
	std::map<int32_t, cubit::Column> columns;
	// queries plan and enqueue under a SHARED lock (several host threads at once: the CUDA runtime is thread-safe and
	// their work lands on different streams or interleaves harmlessly on one); everything that may change table state
	// takes it exclusively (TableLock).  meta_mu guards the small host-side pools and the lazily refreshed counts.
	std::shared_mutex mu;
	std::mutex meta_mu;
	std::atomic<uint64_t> launches {0};
	unsigned long long *d_scratch = nullptr; // popcount scratch
	uint64_t scratch_n = 0;
	std::vector<cubit::ResultHeader *> hdr_pool; // pinned result headers, recycled across queries
	std::vector<std::pair<void *, uint64_t>> wire_pool; // page-locked narrow-wire windows of cubit_gpu_drain, recycled
	uint8_t *h_stage[2] = {nullptr, nullptr};     // pinned staging chunks of the segment / delta upload (lazy, kept)
	cudaEvent_t stage_ev[2] = {nullptr, nullptr};
	// sharded parent (cubit_gpu_create_sharded): no device state of its own, every call fans out to the children
	std::vector<cubit_gpu_table *> shards;
	std::vector<uint64_t> shard_row0; // first parent-local row of every shard, + one end entry
	bool sharded() const {
		return !shards.empty();
	}
};

struct cubit_gpu_fetch_ticket {
	cudaEvent_t ev = nullptr;
	std::vector<cubit_gpu_fetch_ticket *> parts; // sharded result: one per shard touched
};

struct cubit_gpu_result {
	cubit_gpu_table *t = nullptr;
	cudaStream_t stream = nullptr;
	unsigned char *d_block = nullptr; // hdr | ctrl | partials | probe done ctr
	cubit::ResultHeader *d_hdr = nullptr;
	cubit::ResultHeader *h_hdr = nullptr; // pinned
	long long *d_ids = nullptr;
	uint64_t *d_q = nullptr;
	uint64_t *d_q_tmp = nullptr;
	void *d_vals[CUBIT_MAX_PROBE_COLS] = {};
	uint32_t *d_valid[CUBIT_MAX_PROBE_COLS] = {}; // validity of the projected values (bit j = result row j), or nullptr
	bool agg_nulls = false;                       // an aggregate input has a validity mask
	uint32_t val_elem[CUBIT_MAX_PROBE_COLS] = {};
	uint32_t n_cols = 0;
	uint32_t flags = 0;
	int agg_kind = 0;
	cudaEvent_t ev[4] = {};
	cudaEvent_t ev_done = nullptr;
	bool timing = false, probe_timed = false;
	uint64_t probe_fixed_bytes = 0;
	uint64_t probe_widths = 0; // bytes per selected row the probe needs (distinct columns, + row-ID re-read)
	std::mutex fin_mu;         // finish_result may be raced by several consumer threads of one result
	std::atomic<bool> finished {false};
	int fin_rc = 0;
	std::string fin_err;
	cubit_result_info info = {};
	std::atomic<uint32_t> copies_in_flight {0};
	// narrow-wire hand-off (cubit_wire.cu): a ring of per-frame forms between its two kernels, allocated on first use
	std::atomic<uint4 *> d_wire_stats {nullptr};
	std::atomic<uint64_t> wire_stats_cursor {0};
	// sharded result: one child per shard, in row order; count_prefix[i] = rows of the children before i
	std::vector<cubit_gpu_result *> parts;
	std::vector<uint64_t> count_prefix;
};

namespace cubit {

// Exclusive section that may change table state (everything except the query path).  Taking it orders the kernel
// stream behind every pool-stream query in flight; leaving it records the point pool streams must wait for.
struct TableLock {
	cubit_gpu_table *t;
	std::unique_lock<std::shared_mutex> lk;
	explicit TableLock(cubit_gpu_table *t);
	~TableLock();
};

int use_device(const cubit_gpu_table *t);
Index *get_index(cubit_gpu_table *t, int32_t index_id);
void free_column(Column &c);
ColRef col_ref(const Column *c, bool prefer_raw = false);
int refresh_counts(cubit_gpu_table *t, Index *ix); // caller holds t->mu
int ensure_stage(cubit_gpu_table *t);              // the two pinned staging chunks (caller holds t->mu)
void free_delta(DeltaSet &d);
void free_compressed(CompressedStore &cs);
// device pointer of B_v's verbatim words (verbatim index) — compressed indexes have none
inline uint64_t *bv_ptr(const cubit_gpu_table *t, const Index *ix, uint32_t v) {
	return ix->d_bits + (uint64_t)v * t->words_per_bv;
}
// merge-back of one index's pending deltas (caller holds t->mu); `any` reports whether something was merged
int merge_deltas_locked(cubit_gpu_table *t, Index *ix, bool *any);
// expand B_v of a compressed index into `dst` (words_per_bv words, on the kernel stream)
int expand_value_locked(cubit_gpu_table *t, Index *ix, uint32_t v, uint64_t *dst);
// replace the stored form of B_v of a compressed index by the compression of `src` (words_per_bv verbatim words)
int compress_value_locked(cubit_gpu_table *t, Index *ix, uint32_t v0, uint32_t nv, const uint64_t *src);
// pending-delta rows of one value as a host list (persistence)
int delta_rows_locked(cubit_gpu_table *t, Index *ix, uint32_t v, std::vector<int64_t> &rows);
// pending per-value counts of the last add_delta_pairs → rows[], then the automatic merge-back rule (caller holds t->mu)
int delta_settle_locked(cubit_gpu_table *t, Index *ix);
// re-key the delta CSR after the segment count changed (append)
int delta_restride_locked(cubit_gpu_table *t, Index *ix, uint32_t new_n_seg);

void wire_pool_free(cubit_gpu_table *t); // cubit_wire.cu

// sharded fan-out (cubit_sharded.cu)
int sharded_destroy(cubit_gpu_table *t);
int sharded_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out);
int sharded_result_finish(cubit_gpu_result *r);
int sharded_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                  void *const *host_cols, cubit_gpu_fetch_ticket **ticket);
int sharded_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n, uint64_t *host_words,
                           int *all_valid);
int sharded_fetch_bitvector(cubit_gpu_result *r, uint64_t *host_words, uint64_t n_words);
int sharded_free_result(cubit_gpu_result *r);
int sharded_probe(cubit_gpu_table *t, int32_t col_id, const int64_t *host_rowids, uint64_t n, void *host_out,
                  uint64_t *sum_lo, int64_t *sum_hi);

} // namespace cubit
