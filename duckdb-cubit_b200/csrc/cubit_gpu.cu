// cubit_gpu.cu — table shards and CUBIT indexes: the part of the C-ABI (include/cubit_gpu.h) that owns the
// HBM-resident state of one table shard (value bitvectors padded to whole segments — verbatim or as roaring-style
// containers —, the pending-delta lists, decoded column slices).  The other translation units of the host side:
//   cubit_columns.cu  column upload / decode of on-disk segments / append
//   cubit_delta.cu    device-side ingestion of pending deltas, merge-back
//   cubit_persist.cu  index images
//   cubit_query.cu    query planning, kernel launches, result hand-off
//   cubit_sharded.cu  one table over several devices
// Object model and locking: table.h.  There is no CPU fallback anywhere: every compute entry point launches CUDA
// work or fails.
#include "table.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace cubit;

// ------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

namespace cubit {

int fail(int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	try {
		g_last_error = buf;
	} catch (...) {
	}
	return code;
}

const char *last_error_cstr() {
	return g_last_error.c_str();
}

TableLock::TableLock(cubit_gpu_table *tp) : t(tp), lk(tp->mu) {
	if (t->sharded() || !t->stream) {
		return;
	}
	cudaSetDevice(t->device);
	for (int i = 0; i < kAggStreams; i++) {
		if (t->agg_used[i]) {
			cudaStreamWaitEvent(t->stream, t->agg_last[i], 0);
			t->agg_used[i] = false;
		}
	}
}

TableLock::~TableLock() {
	if (t->sharded() || !t->mut_event || !t->stream) {
		return;
	}
	cudaSetDevice(t->device);
	if (cudaEventRecord(t->mut_event, t->stream) == cudaSuccess) {
		t->mut_recorded = true;
	}
}

int use_device(const cubit_gpu_table *t) {
	CU_TRY(cudaSetDevice(t->device));
	return CUBIT_OK;
}

Index *get_index(cubit_gpu_table *t, int32_t index_id) {
	if (index_id < 0 || (size_t)index_id >= t->indexes.size()) {
		return nullptr;
	}
	return t->indexes[index_id];
}

// Which form of the column a probe reads when both are resident (cubit_gpu_pack_column keep_raw = 1): see
// cubit_query.cu (the choice is made per query from the measured crossover).
ColRef col_ref(const Column *c, bool prefer_raw) {
	ColRef r;
	r.raw = nullptr;
	r.words = nullptr;
	r.hdr = nullptr;
	if (c) {
		if (c->packed() && !(prefer_raw && c->d)) {
			r.words = c->d_words;
			r.hdr = c->d_hdr;
		} else {
			r.raw = static_cast<const long long *>(c->d);
		}
	}
	return r;
}

void free_column(Column &c) {
	if (c.d) {
		cudaFree(c.d);
	}
	if (c.d_words) {
		cudaFree(c.d_words);
	}
	if (c.d_hdr) {
		cudaFree(c.d_hdr);
	}
	if (c.d_valid) {
		cudaFree(c.d_valid);
	}
	c = Column();
}

void free_delta(DeltaSet &d) {
	if (d.d_off) {
		cudaFree(d.d_off);
	}
	if (d.d_ent) {
		cudaFree(d.d_ent);
	}
	d.d_off = nullptr;
	d.d_ent = nullptr;
	d.n_ent = d.cap_ent = 0;
	d.n_seg = 0;
	std::fill(d.rows.begin(), d.rows.end(), 0);
	if (d.ev_voff) {
		cudaEventSynchronize(d.ev_voff);
	}
	d.voff_pending = false;
}

void free_compressed(CompressedStore &cs) {
	if (cs.d_dir) {
		cudaFree(cs.d_dir);
	}
	if (cs.d_pool) {
		cudaFree(cs.d_pool);
	}
	cs = CompressedStore();
}

int ensure_stage(cubit_gpu_table *t) {
	for (int b = 0; b < 2; b++) {
		if (!t->h_stage[b]) {
			CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&t->h_stage[b]), kStageChunk));
			CU_TRY(cudaEventCreateWithFlags(&t->stage_ev[b], cudaEventDisableTiming));
		}
	}
	return CUBIT_OK;
}

} // namespace cubit

// ------------------------------------------------------------------ library
extern "C" int cubit_gpu_abi_version(void) {
	return CUBIT_GPU_ABI_VERSION;
}

extern "C" const char *cubit_gpu_last_error(void) {
	return g_last_error.c_str();
}

extern "C" int cubit_gpu_device_count(int *count) {
	if (!count) {
		return fail(CUBIT_EINVAL, "count is NULL");
	}
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		*count = 0;
		return fail(CUBIT_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
	}
	*count = n;
	return CUBIT_OK;
}

// -------------------------------------------------------------------- table
extern "C" int cubit_gpu_create(int device, uint64_t n_rows, int64_t row_base, uint32_t seg_bits,
                                cubit_gpu_table **out) {
	ABI_BEGIN
	if (!out) {
		return fail(CUBIT_EINVAL, "out is NULL");
	}
	*out = nullptr;
	if (seg_bits != 32768 && seg_bits != 65536 && seg_bits != 131072) {
		return fail(CUBIT_EINVAL, "seg_bits must be 32768, 65536 or 131072 (got %u)", seg_bits);
	}
	if (n_rows == 0) {
		return fail(CUBIT_EINVAL, "n_rows must be > 0");
	}
	if (row_base < 0 || (uint64_t)row_base % 64 != 0) {
		return fail(CUBIT_EINVAL, "row_base must be a non-negative multiple of 64");
	}
	if ((n_rows + seg_bits - 1) / seg_bits > 0x7fffffffull) {
		return fail(CUBIT_EINVAL, "too many segments");
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		return fail(CUBIT_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
	}
	if (device < 0 || device >= ndev) {
		return fail(CUBIT_EINVAL, "device %d out of range (have %d)", device, ndev);
	}
	CU_TRY(cudaSetDevice(device));
	if (const char *g = getenv("CUBIT_L2_FETCH_BYTES")) { // experiment knob (profiles/): L2→DRAM fetch granularity
		cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
	}
	cubit_gpu_table *t = new cubit_gpu_table();
	t->device = device;
	t->n_rows = n_rows;
	t->row_base = row_base;
	t->seg_bits = seg_bits;
	t->seg_words = seg_bits / 64;
	t->n_seg = (uint32_t)((n_rows + seg_bits - 1) / seg_bits);
	t->n_words = (n_rows + 63) / 64;
	// capacity in whole segments, rounded up to a multiple of 4 (zero-filled like every pad bit): short queries scan
	// 2 or 4 consecutive segments as one tile (cubit_query.cu, tile_mult), and the last tile must not leave the bitvector
	t->words_per_bv = (((uint64_t)t->n_seg + 3) & ~3ull) * t->seg_words;
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) {
		delete t;
		return fail(CUBIT_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
	}
	if (prop.major < 10) {
		delete t;
		return fail(CUBIT_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
		            prop.major, prop.minor);
	}
	t->sm_count = prop.multiProcessorCount;
	e = cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking);
	for (int c = 0; c < kCopyStreams && e == cudaSuccess; c++) {
		e = cudaStreamCreateWithFlags(&t->copy_stream[c], cudaStreamNonBlocking);
	}
	for (int c = 0; c < kAggStreams && e == cudaSuccess; c++) {
		e = cudaStreamCreateWithFlags(&t->agg_stream[c], cudaStreamNonBlocking);
		if (e == cudaSuccess) {
			e = cudaEventCreateWithFlags(&t->agg_last[c], cudaEventDisableTiming);
		}
	}
	if (e == cudaSuccess) {
		e = cudaEventCreateWithFlags(&t->mut_event, cudaEventDisableTiming);
	}
	if (e != cudaSuccess) {
		cubit_gpu_destroy(t);
		return fail(CUBIT_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
	}
	t->stream = t->own_stream;
	// keep freed result buffers in the stream-ordered pool (no cudaMalloc per query)
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t thresh = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
	}
	*out = t;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_destroy(cubit_gpu_table *t) {
	if (!t) {
		return CUBIT_OK;
	}
	if (t->sharded()) {
		return sharded_destroy(t);
	}
	cudaSetDevice(t->device);
	if (t->stream) {
		cudaStreamSynchronize(t->stream);
	}
	wire_pool_free(t);
	for (auto &cs : t->copy_stream) {
		if (cs) {
			cudaStreamSynchronize(cs);
			cudaStreamDestroy(cs);
		}
	}
	for (int c = 0; c < kAggStreams; c++) {
		if (t->agg_stream[c]) {
			cudaStreamSynchronize(t->agg_stream[c]);
			cudaStreamDestroy(t->agg_stream[c]);
		}
		if (t->agg_last[c]) {
			cudaEventDestroy(t->agg_last[c]);
		}
	}
	if (t->mut_event) {
		cudaEventDestroy(t->mut_event);
	}
	for (auto ev : t->ev_pool) {
		cudaEventDestroy(ev);
	}
	for (Index *ix : t->indexes) {
		if (!ix) {
			continue;
		}
		free_delta(ix->delta);
		if (ix->delta.h_voff) {
			cudaFreeHost(ix->delta.h_voff);
		}
		if (ix->delta.ev_voff) {
			cudaEventDestroy(ix->delta.ev_voff);
		}
		free_compressed(ix->cs);
		if (ix->d_bits) {
			cudaFree(ix->d_bits);
		}
		delete ix;
	}
	for (auto &kv : t->columns) {
		free_column(kv.second);
	}
	for (int b = 0; b < 2; b++) {
		if (t->h_stage[b]) {
			cudaFreeHost(t->h_stage[b]);
		}
		if (t->stage_ev[b]) {
			cudaEventDestroy(t->stage_ev[b]);
		}
	}
	if (t->d_scratch) {
		cudaFree(t->d_scratch);
	}
	for (ResultHeader *h : t->hdr_pool) {
		cudaFreeHost(h);
	}
	if (t->own_stream) {
		cudaStreamDestroy(t->own_stream);
	}
	delete t;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_set_stream(cubit_gpu_table *t, void *cuda_stream) {
	if (!t) {
		return fail(CUBIT_EINVAL, "table is NULL");
	}
	if (t->sharded()) {
		return fail(CUBIT_ESTATE, "a sharded table runs on its shards' own streams");
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	t->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->own_stream;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_words_per_bitvector(const cubit_gpu_table *t, uint64_t *n_words) {
	if (!t || !n_words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*n_words = t->n_words;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_row_count(const cubit_gpu_table *t, uint64_t *n_rows) {
	if (!t || !n_rows) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*n_rows = t->n_rows;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_launch_count(const cubit_gpu_table *t, uint64_t *n) {
	if (!t || !n) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	uint64_t total = t->launches.load();
	for (auto *s : t->shards) {
		total += s->launches.load();
	}
	*n = total;
	return CUBIT_OK;
}

// -------------------------------------------------------------------- index
static int index_create(cubit_gpu_table *t, uint32_t cardinality, bool compressed, int32_t *index_id) {
	if (!t || !index_id) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (cardinality == 0 || cardinality > (1u << 20)) {
		return fail(CUBIT_EINVAL, "cardinality %u out of range", cardinality);
	}
	if (t->sharded()) {
		int32_t id = -1;
		for (auto *s : t->shards) {
			int32_t sid = -1;
			int rc = index_create(s, cardinality, compressed, &sid);
			if (rc) {
				return rc;
			}
			if (id >= 0 && sid != id) {
				return fail(CUBIT_ESTATE, "shards disagree on the index id");
			}
			id = sid;
		}
		*index_id = id;
		return CUBIT_OK;
	}
	if (compressed && t->seg_bits > 65536) {
		return fail(CUBIT_EINVAL, "compressed indexes need seg_bits <= 65536 (16-bit positions inside a segment)");
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = new Index();
	ix->card = cardinality;
	ix->compressed = compressed;
	cudaError_t e = cudaSuccess;
	if (compressed) {
		// every container EMPTY (directory entry 0); the pool grows as bitvectors arrive
		const uint64_t n_seg_cap = t->words_per_bv / t->seg_words;
		const size_t bytes = (size_t)cardinality * n_seg_cap * 8;
		e = cudaMalloc(&ix->cs.d_dir, bytes);
		if (e == cudaSuccess) {
			e = cudaMemsetAsync(ix->cs.d_dir, 0, bytes, t->stream);
		}
		ix->cs.n_seg_cap = n_seg_cap;
	} else {
		const size_t bytes = (size_t)cardinality * t->words_per_bv * 8;
		e = cudaMalloc(&ix->d_bits, bytes);
		if (e == cudaSuccess) {
			e = cudaMemsetAsync(ix->d_bits, 0, bytes, t->stream);
		}
	}
	if (e != cudaSuccess) {
		free_compressed(ix->cs);
		if (ix->d_bits) {
			cudaFree(ix->d_bits);
		}
		delete ix;
		return fail(e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "index allocation (cardinality %u): %s",
		            cardinality, cudaGetErrorString(e));
	}
	ix->counts.assign(cardinality, 0);
	ix->counts_valid = true; // all zero
	ix->delta.rows.assign(cardinality, 0);
	t->indexes.push_back(ix);
	*index_id = (int32_t)t->indexes.size() - 1;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_index_create(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id) {
	ABI_BEGIN
	return index_create(t, cardinality, false, index_id);
	ABI_END
}

extern "C" int cubit_gpu_index_create_compressed(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id) {
	ABI_BEGIN
	return index_create(t, cardinality, true, index_id);
	ABI_END
}

// a [lo, hi) row range of the sharded parent → the shard holding `lo` (rows never straddle: callers split)
static size_t shard_of_row(const cubit_gpu_table *t, uint64_t row) {
	size_t s = std::upper_bound(t->shard_row0.begin(), t->shard_row0.end(), row) - t->shard_row0.begin();
	return s == 0 ? 0 : s - 1;
}

// temporary verbatim bitvectors on the device (compressed-index paths): [nv][words_per_bv], zeroed
static int alloc_temp_bits(cubit_gpu_table *t, uint32_t nv, uint64_t **out) {
	const size_t bytes = (size_t)nv * t->words_per_bv * 8;
	CU_TRY(cudaMallocAsync((void **)out, bytes, t->stream));
	CU_TRY(cudaMemsetAsync(*out, 0, bytes, t->stream));
	return CUBIT_OK;
}

extern "C" int cubit_gpu_upload_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id,
                                          const uint64_t *words, uint64_t n_words) {
	ABI_BEGIN
	if (!t || !words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "n_words %llu != ceil(n_rows/64) = %llu", (unsigned long long)n_words,
		            (unsigned long long)t->n_words);
	}
	const unsigned tail = (unsigned)(t->n_rows % 64);
	if (tail && (words[n_words - 1] >> tail) != 0) {
		return fail(CUBIT_EINVAL, "bits at positions >= n_rows must be zero");
	}
	if (t->sharded()) { // shard boundaries are multiples of the segment size, hence of 64
		for (size_t s = 0; s < t->shards.size(); s++) {
			const uint64_t w0 = t->shard_row0[s] / 64;
			int rc = cubit_gpu_upload_bitvector(t->shards[s], index_id, value_id, words + w0, t->shards[s]->n_words);
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
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	if (ix->compressed) {
		uint64_t *tmp = nullptr;
		int rc = alloc_temp_bits(t, 1, &tmp);
		if (rc) {
			return rc;
		}
		cudaError_t e = cudaMemcpyAsync(tmp, words, n_words * 8, cudaMemcpyHostToDevice, t->stream);
		rc = e == cudaSuccess ? compress_value_locked(t, ix, value_id, 1, tmp)
		                      : fail(CUBIT_ECUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
		cudaFreeAsync(tmp, t->stream);
		if (rc) {
			return rc;
		}
	} else {
		CU_TRY(cudaMemcpyAsync(bv_ptr(t, ix, value_id), words, n_words * 8, cudaMemcpyHostToDevice, t->stream));
	}
	CU_TRY(cudaStreamSynchronize(t->stream)); // `words` may be pageable: the caller may reuse it on return
	ix->counts_valid = false;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_upload_bitvector_wah(cubit_gpu_table *t, int32_t index_id, uint32_t value_id,
                                              const cubit_wah_bitvector *bv) {
	ABI_BEGIN
	if (!t || !bv || (!bv->words && bv->n_words)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (bv->active_nbits > 30 || (bv->active_nbits < 32 && (bv->active_val >> bv->active_nbits) != 0)) {
		return fail(CUBIT_EINVAL, "bad active word (nbits %u)", bv->active_nbits);
	}
	if (t->sharded()) {
		return fail(CUBIT_ESTATE, "WAH upload addresses one shard: upload per shard (a WAH stream cannot be cut at a row)");
	}
	// one pass over the compressed words: validate, and the 31-bit-group prefix per block of kWahBlockWords
	const uint64_t n_blocks = std::max<uint64_t>(1, (bv->n_words + kWahBlockWords - 1) / kWahBlockWords);
	std::vector<unsigned long long> block_group0(n_blocks, 0);
	unsigned long long groups = 0;
	for (uint64_t i = 0; i < bv->n_words; i++) {
		if (i % kWahBlockWords == 0) {
			block_group0[i / kWahBlockWords] = groups;
		}
		const uint32_t w = bv->words[i];
		if (w & 0x80000000u) {
			if ((w & 0x3fffffffu) == 0) {
				return fail(CUBIT_EINVAL, "WAH word %llu: zero-length fill", (unsigned long long)i);
			}
			groups += w & 0x3fffffffu;
		} else {
			groups++;
		}
		if (groups * 31ull > t->n_rows) {
			return fail(CUBIT_EINVAL, "WAH bitvector is longer than the table (%llu rows)", (unsigned long long)t->n_rows);
		}
	}
	if (groups * 31ull + bv->active_nbits > t->n_rows) {
		return fail(CUBIT_EINVAL, "WAH bitvector describes %llu bits, table has %llu rows",
		            (unsigned long long)(groups * 31ull + bv->active_nbits), (unsigned long long)t->n_rows);
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad index/value (%d, %u)", index_id, value_id);
	}
	uint32_t *d_wah = nullptr;
	unsigned long long *d_blk = nullptr;
	uint64_t *tmp = nullptr;
	CU_TRY(cudaMalloc(&d_wah, (bv->n_words + 4) * 4));
	cudaError_t e = cudaMalloc(&d_blk, n_blocks * 8);
	int rc = CUBIT_OK;
	unsigned long long *dst = nullptr;
	if (e == cudaSuccess && ix->compressed) { // expand into a scratch bitvector, then store it as containers
		rc = alloc_temp_bits(t, 1, &tmp);
		dst = reinterpret_cast<unsigned long long *>(tmp);
	} else if (e == cudaSuccess) {
		dst = reinterpret_cast<unsigned long long *>(bv_ptr(t, ix, value_id));
		e = cudaMemsetAsync(dst, 0, t->words_per_bv * 8, t->stream);
	}
	if (e == cudaSuccess && rc == CUBIT_OK) {
		e = cudaMemcpyAsync(d_wah, bv->words, bv->n_words * 4, cudaMemcpyHostToDevice, t->stream);
	}
	if (e == cudaSuccess && rc == CUBIT_OK) {
		e = cudaMemcpyAsync(d_blk, block_group0.data(), n_blocks * 8, cudaMemcpyHostToDevice, t->stream);
	}
	if (e == cudaSuccess && rc == CUBIT_OK) {
		e = launch_wah_expand(d_wah, bv->n_words, d_blk, groups, bv->active_val, bv->active_nbits, dst, t->stream);
		t->launches++;
	}
	if (e == cudaSuccess && rc == CUBIT_OK && ix->compressed) {
		rc = compress_value_locked(t, ix, value_id, 1, tmp);
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(t->stream);
	}
	if (tmp) {
		cudaFreeAsync(tmp, t->stream);
	}
	cudaFree(d_wah);
	if (d_blk) {
		cudaFree(d_blk);
	}
	CU_TRY(e);
	if (rc) {
		return rc;
	}
	ix->counts_valid = false;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_download_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *words,
                                            uint64_t n_words) {
	ABI_BEGIN
	if (!t || !words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "n_words mismatch");
	}
	if (t->sharded()) {
		for (size_t s = 0; s < t->shards.size(); s++) {
			int rc = cubit_gpu_download_bitvector(t->shards[s], index_id, value_id, words + t->shard_row0[s] / 64,
			                                      t->shards[s]->n_words);
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
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	if (ix->compressed) {
		uint64_t *tmp = nullptr;
		CU_TRY(cudaMallocAsync((void **)&tmp, t->words_per_bv * 8, t->stream));
		int rc = expand_value_locked(t, ix, value_id, tmp);
		cudaError_t e = cudaSuccess;
		if (rc == CUBIT_OK) {
			e = cudaMemcpyAsync(words, tmp, n_words * 8, cudaMemcpyDeviceToHost, t->stream);
		}
		if (e == cudaSuccess) {
			e = cudaStreamSynchronize(t->stream);
		}
		cudaFreeAsync(tmp, t->stream);
		CU_TRY(e);
		return rc;
	}
	CU_TRY(cudaMemcpyAsync(words, bv_ptr(t, ix, value_id), n_words * 8, cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	return CUBIT_OK;
	ABI_END
}

namespace cubit {

int refresh_counts(cubit_gpu_table *t, Index *ix) {
	if (ix->counts_valid) {
		return CUBIT_OK;
	}
	if (t->scratch_n < ix->card) {
		if (t->d_scratch) {
			CU_TRY(cudaStreamSynchronize(t->stream));
			cudaFree(t->d_scratch);
			t->d_scratch = nullptr;
		}
		CU_TRY(cudaMalloc(&t->d_scratch, sizeof(unsigned long long) * ix->card));
		t->scratch_n = ix->card;
	}
	if (ix->compressed) {
		CU_TRY(launch_compressed_counts(ix->cs.d_dir, ix->cs.n_seg_cap, t->n_seg, ix->card, ix->cs.d_pool, t->seg_words,
		                                t->d_scratch, t->stream));
		t->launches++;
	} else {
		CU_TRY(launch_popcount_many(ix->d_bits, t->words_per_bv, ix->card, t->d_scratch, t->stream));
		t->launches += (ix->card + 32767) / 32768;
	}
	static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
	CU_TRY(cudaMemcpyAsync(ix->counts.data(), t->d_scratch, sizeof(uint64_t) * ix->card, cudaMemcpyDeviceToHost,
	                       t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	ix->counts_valid = true;
	return CUBIT_OK;
}

// ---- compressed containers: expand one value / store nv consecutive values from verbatim words
int expand_value_locked(cubit_gpu_table *t, Index *ix, uint32_t v, uint64_t *dst) {
	CU_TRY(launch_container_expand(ix->cs.d_dir + (uint64_t)v * ix->cs.n_seg_cap, ix->cs.d_pool, t->words_per_bv / t->seg_words,
	                               t->n_seg, t->seg_words, dst, t->stream));
	t->launches++;
	return CUBIT_OK;
}

int compress_value_locked(cubit_gpu_table *t, Index *ix, uint32_t v0, uint32_t nv, const uint64_t *src) {
	CompressedStore &cs = ix->cs;
	// worst case: every segment of every value becomes a BITMAP container
	const uint64_t worst = (uint64_t)nv * t->n_seg * t->seg_words * 8;
	if (cs.pool_bytes + worst + 16 > cs.pool_cap) {
		const uint64_t cap = std::max<uint64_t>(cs.pool_bytes + worst + 16, cs.pool_cap + cs.pool_cap / 2);
		uint8_t *np = nullptr;
		CU_TRY(cudaMalloc(&np, cap));
		if (cs.pool_bytes) {
			CU_TRY(cudaMemcpyAsync(np, cs.d_pool, cs.pool_bytes, cudaMemcpyDeviceToDevice, t->stream));
		}
		CU_TRY(cudaStreamSynchronize(t->stream)); // scans in flight still read the old pool
		if (cs.d_pool) {
			cudaFree(cs.d_pool);
		}
		cs.d_pool = np;
		cs.pool_cap = cap;
	}
	// [0] pool cursor (bytes), [1] bytes of the containers these values had before (garbage from now on)
	unsigned long long *d_cur = nullptr;
	CU_TRY(cudaMallocAsync((void **)&d_cur, 16, t->stream));
	unsigned long long init[2] = {cs.pool_bytes, 0};
	CU_TRY(cudaMemcpyAsync(d_cur, init, 16, cudaMemcpyHostToDevice, t->stream));
	CU_TRY(launch_container_compress(src, t->words_per_bv, nv, t->n_seg, t->seg_words,
	                                 cs.d_dir + (uint64_t)v0 * cs.n_seg_cap, cs.n_seg_cap, cs.d_pool, d_cur, t->stream));
	t->launches++;
	unsigned long long fin[2] = {0, 0};
	CU_TRY(cudaMemcpyAsync(fin, d_cur, 16, cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	cudaFreeAsync(d_cur, t->stream);
	cs.pool_bytes = fin[0];
	cs.garbage_bytes += fin[1];
	ix->counts_valid = false;
	return CUBIT_OK;
}

} // namespace cubit

extern "C" int cubit_gpu_bitvector_count(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *count) {
	ABI_BEGIN
	if (!t || !count) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		uint64_t total = 0;
		for (auto *s : t->shards) {
			uint64_t c = 0;
			int rc = cubit_gpu_bitvector_count(s, index_id, value_id, &c);
			if (rc) {
				return rc;
			}
			total += c;
		}
		*count = total;
		return CUBIT_OK;
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	int rc = refresh_counts(t, ix);
	if (rc) {
		return rc;
	}
	*count = ix->counts[value_id];
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_index_info(cubit_gpu_table *t, int32_t index_id, cubit_index_info *info) {
	ABI_BEGIN
	if (!t || !info) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	memset(info, 0, sizeof(*info));
	if (t->sharded()) {
		for (auto *s : t->shards) {
			cubit_index_info si;
			int rc = cubit_gpu_index_info(s, index_id, &si);
			if (rc) {
				return rc;
			}
			info->cardinality = si.cardinality;
			info->compressed = si.compressed;
			info->resident_bytes += si.resident_bytes;
			info->verbatim_bytes += si.verbatim_bytes;
			info->delta_entries += si.delta_entries;
			info->auto_merges += si.auto_merges;
		}
		return CUBIT_OK;
	}
	TableLock lk(t);
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	int src = delta_settle_locked(t, ix);
	if (src) {
		return src;
	}
	info->cardinality = ix->card;
	info->compressed = ix->compressed ? 1u : 0u;
	info->verbatim_bytes = (uint64_t)ix->card * t->n_seg * t->seg_words * 8;
	info->resident_bytes = ix->compressed ? ix->cs.pool_bytes + (uint64_t)ix->card * ix->cs.n_seg_cap * 8
	                                      : (uint64_t)ix->card * t->words_per_bv * 8;
	info->delta_entries = ix->delta.n_ent;
	info->auto_merges = ix->auto_merges;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_index_build(cubit_gpu_table *t, int32_t index_id, int32_t col_id, int64_t base_value) {
	ABI_BEGIN
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		for (auto *s : t->shards) {
			int rc = cubit_gpu_index_build(s, index_id, col_id, base_value);
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
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	const Column &c = it->second;
	if (!c.d) {
		return fail(CUBIT_ESTATE, "column %d is resident only in packed form; build the index before packing", col_id);
	}
	if (c.n != t->n_rows) {
		return fail(CUBIT_EINVAL, "column %d has %llu rows, table has %llu", col_id, (unsigned long long)c.n,
		            (unsigned long long)t->n_rows);
	}
	if (ix->delta.n_ent) {
		return fail(CUBIT_ESTATE, "index %d has pending deltas; merge them before a rebuild", index_id);
	}
	// NULL keys are not indexed (plan_create_index.cpp:60-78): rows whose validity bit is 0 set no bit
	const unsigned long long *valid = c.d_valid;
	int launches = 0;
	if (!ix->compressed) {
		// the build kernel writes every word of the 4096-row tiles that hold rows (64 words each); only the capacity
		// behind the last tile has to be zeroed here — not the whole index (12.5 GB at cardinality 100 and 10^9 rows)
		const uint64_t built_words = std::min<uint64_t>(t->words_per_bv, (t->n_rows + 4095) / 4096 * 64);
		if (built_words < t->words_per_bv) {
			CU_TRY(cudaMemset2DAsync(ix->d_bits + built_words, t->words_per_bv * 8, 0, (t->words_per_bv - built_words) * 8,
			                         ix->card, t->stream));
		}
		CU_TRY(launch_index_build(c.d, c.elem, valid, 0, t->n_rows, base_value, ix->card, ix->d_bits, t->words_per_bv,
		                          t->sm_count, t->stream, &launches));
		t->launches += launches;
	} else {
		// value batches: verbatim scratch for `nb` values at a time (≤ ~2 GiB), stored as containers, scratch reused
		const uint64_t bv_bytes = t->words_per_bv * 8;
		uint32_t nb = (uint32_t)std::min<uint64_t>(ix->card, std::max<uint64_t>(1, (2ull << 30) / bv_bytes));
		uint64_t *tmp = nullptr;
		CU_TRY(cudaMalloc(&tmp, (size_t)nb * bv_bytes));
		int rc = CUBIT_OK;
		for (uint32_t v0 = 0; v0 < ix->card && rc == CUBIT_OK; v0 += nb) {
			const uint32_t nv = std::min<uint32_t>(nb, ix->card - v0);
			cudaError_t e = cudaMemsetAsync(tmp, 0, (size_t)nv * bv_bytes, t->stream);
			if (e == cudaSuccess) {
				e = launch_index_build(c.d, c.elem, valid, 0, t->n_rows, base_value + v0, nv, tmp, t->words_per_bv,
				                       t->sm_count, t->stream, &launches);
			}
			if (e != cudaSuccess) {
				rc = fail(CUBIT_ECUDA, "index build: %s", cudaGetErrorString(e));
				break;
			}
			t->launches += launches;
			rc = compress_value_locked(t, ix, v0, nv, tmp);
		}
		cudaStreamSynchronize(t->stream);
		cudaFree(tmp);
		if (rc) {
			return rc;
		}
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	ix->counts_valid = false;
	ix->src_col = col_id;
	ix->src_base = base_value;
	return CUBIT_OK;
	ABI_END
}
