// cubit_gpu.cu — host side of the C-ABI declared in include/cubit_gpu.h.
//
// Owns the HBM-resident state of one table shard (value bitvectors padded to
// whole segments, pending-delta CSR lists, decoded column slices), plans a query
// (flattens the AND-of-ORs predicate into an ordered stream list, bounds the
// result size from per-bitvector cardinalities) and launches the kernels in
// scan_kernel.cu / aux_kernels.cu on the table's stream.  There is no CPU
// fallback anywhere in this file: every compute entry point launches CUDA work
// or fails.
#include "../../include/cubit_gpu.h"
#include "kernels.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace cubit;

// ------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_last_error = buf;
	return code;
}

#define CU_TRY(expr)                                                                                                   \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			return fail(_e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "%s: %s (%s:%d)", #expr,         \
			            cudaGetErrorString(_e), __FILE__, __LINE__);                                                   \
		}                                                                                                              \
	} while (0)

// ------------------------------------------------------------------ objects
struct Delta {
	uint32_t *d_off = nullptr; // [n_seg + 1]
	DeltaEnt *d_ent = nullptr;
	uint64_t n_ent = 0;  // delta words
	uint64_t n_rows = 0; // flipped rows (after cancellation)
};

struct Index {
	uint32_t card = 0;
	uint64_t *d_bits = nullptr;   // [card][words_per_bv]
	std::vector<uint64_t> counts; // popcount of every B_v as stored
	bool counts_valid = false;
	int32_t src_col = -1;         // column the index was built from (cubit_gpu_index_build), -1 = uploaded
	int64_t src_base = 0;
	std::vector<Delta> deltas; // [card]
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
	// validity mask (ValidityMask layout, bit = 1: valid), nullptr = the column holds no NULLs
	unsigned long long *d_valid = nullptr;
	uint64_t valid_cap_words = 0;
	bool packed() const {
		return d_words != nullptr;
	}
};

// Which form of the column a probe reads when both are resident (cubit_gpu_pack_column keep_raw = 1).  Packed
// moves fewer bytes but costs more instructions and one more dependent load per value: measured
// (profiles/r1_experiment_packed_payload.log) it wins in the middle band — bit-driven probe, 1/256 .. 1/6 of the
// rows selected, where the raw gather is DRAM-bound on 128-byte fetches — and loses for dense selections (the
// decode becomes issue-bound) and for the sparse gather over row IDs (latency-bound).  prefer_raw = those two.
static ColRef col_ref(const Column *c, bool prefer_raw = false) {
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

static void free_column(Column &c) {
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

struct cubit_gpu_table {
	int device = 0;
	int sm_count = 0;
	uint64_t n_rows = 0;
	int64_t row_base = 0;
	uint32_t seg_bits = 0, seg_words = 0, n_seg = 0;
	uint64_t n_words = 0;      // ceil(n_rows / 64)
	uint64_t words_per_bv = 0; // n_seg * seg_words
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	std::vector<Index *> indexes;
	std::map<int32_t, Column> columns;
	std::mutex mu;
	uint64_t launches = 0;
	unsigned long long *d_scratch = nullptr; // popcount scratch
	uint64_t scratch_n = 0;
	std::vector<ResultHeader *> hdr_pool; // pinned result headers, recycled across queries
	uint8_t *h_stage[2] = {nullptr, nullptr}; // pinned staging chunks of the segment upload (lazy, kept)
	cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};
constexpr uint64_t kStageChunk = 16ull << 20;

struct cubit_gpu_result {
	cubit_gpu_table *t = nullptr;
	cudaStream_t stream = nullptr;
	unsigned char *d_block = nullptr; // hdr | ctrl | partials | probe done ctr
	ResultHeader *d_hdr = nullptr;
	ResultHeader *h_hdr = nullptr; // pinned
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
	bool finished = false;
	cubit_result_info info = {};
};

static int use_device(const cubit_gpu_table *t) {
	CU_TRY(cudaSetDevice(t->device));
	return CUBIT_OK;
}

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
	if (const char *g = getenv("CUBIT_L2_FETCH_BYTES")) { // experiment knob (profiles/): L2→DRAM fetch granularity
		cudaSetDevice(device);
		cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
	}
	if (device < 0 || device >= ndev) {
		return fail(CUBIT_EINVAL, "device %d out of range (have %d)", device, ndev);
	}
	CU_TRY(cudaSetDevice(device));
	cubit_gpu_table *t = new (std::nothrow) cubit_gpu_table();
	if (!t) {
		return fail(CUBIT_ENOMEM, "host allocation failed");
	}
	t->device = device;
	t->n_rows = n_rows;
	t->row_base = row_base;
	t->seg_bits = seg_bits;
	t->seg_words = seg_bits / 64;
	t->n_seg = (uint32_t)((n_rows + seg_bits - 1) / seg_bits);
	t->n_words = (n_rows + 63) / 64;
	t->words_per_bv = (uint64_t)t->n_seg * t->seg_words;
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
	if (e != cudaSuccess) {
		delete t;
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
}

static void free_delta(Delta &d) {
	if (d.d_off) {
		cudaFree(d.d_off);
	}
	if (d.d_ent) {
		cudaFree(d.d_ent);
	}
	d = Delta();
}

extern "C" int cubit_gpu_destroy(cubit_gpu_table *t) {
	if (!t) {
		return CUBIT_OK;
	}
	cudaSetDevice(t->device);
	cudaStreamSynchronize(t->stream);
	for (Index *ix : t->indexes) {
		if (!ix) {
			continue;
		}
		for (Delta &d : ix->deltas) {
			free_delta(d);
		}
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
	std::lock_guard<std::mutex> lk(t->mu);
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

extern "C" int cubit_gpu_launch_count(const cubit_gpu_table *t, uint64_t *n) {
	if (!t || !n) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*n = t->launches;
	return CUBIT_OK;
}

// -------------------------------------------------------------------- index
static Index *get_index(cubit_gpu_table *t, int32_t index_id) {
	if (index_id < 0 || (size_t)index_id >= t->indexes.size()) {
		return nullptr;
	}
	return t->indexes[index_id];
}

extern "C" int cubit_gpu_index_create(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id) {
	if (!t || !index_id) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (cardinality == 0 || cardinality > (1u << 20)) {
		return fail(CUBIT_EINVAL, "cardinality %u out of range", cardinality);
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = new (std::nothrow) Index();
	if (!ix) {
		return fail(CUBIT_ENOMEM, "host allocation failed");
	}
	ix->card = cardinality;
	const size_t bytes = (size_t)cardinality * t->words_per_bv * 8;
	cudaError_t e = cudaMalloc(&ix->d_bits, bytes);
	if (e != cudaSuccess) {
		delete ix;
		return fail(CUBIT_ENOMEM, "cudaMalloc(%zu bytes) for index: %s", bytes, cudaGetErrorString(e));
	}
	e = cudaMemsetAsync(ix->d_bits, 0, bytes, t->stream);
	if (e != cudaSuccess) {
		cudaFree(ix->d_bits);
		delete ix;
		return fail(CUBIT_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
	}
	ix->counts.assign(cardinality, 0);
	ix->counts_valid = true; // all zero
	ix->deltas.resize(cardinality);
	t->indexes.push_back(ix);
	*index_id = (int32_t)t->indexes.size() - 1;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_upload_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id,
                                          const uint64_t *words, uint64_t n_words) {
	if (!t || !words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "n_words %llu != ceil(n_rows/64) = %llu", (unsigned long long)n_words,
		            (unsigned long long)t->n_words);
	}
	const unsigned tail = (unsigned)(t->n_rows % 64);
	if (tail && (words[n_words - 1] >> tail) != 0) {
		return fail(CUBIT_EINVAL, "bits at positions >= n_rows must be zero");
	}
	uint64_t *dst = ix->d_bits + (uint64_t)value_id * t->words_per_bv;
	CU_TRY(cudaMemcpyAsync(dst, words, n_words * 8, cudaMemcpyHostToDevice, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	ix->counts_valid = false;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_upload_bitvector_wah(cubit_gpu_table *t, int32_t index_id, uint32_t value_id,
                                              const cubit_wah_bitvector *bv) {
	if (!t || !bv || (!bv->words && bv->n_words)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (bv->active_nbits > 30 || (bv->active_nbits < 32 && (bv->active_val >> bv->active_nbits) != 0)) {
		return fail(CUBIT_EINVAL, "bad active word (nbits %u)", bv->active_nbits);
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
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad index/value (%d, %u)", index_id, value_id);
	}
	uint32_t *d_wah = nullptr;
	unsigned long long *d_blk = nullptr;
	CU_TRY(cudaMalloc(&d_wah, (bv->n_words + 4) * 4));
	cudaError_t e = cudaMalloc(&d_blk, n_blocks * 8);
	unsigned long long *dst =
	    reinterpret_cast<unsigned long long *>(ix->d_bits + (uint64_t)value_id * t->words_per_bv);
	if (e == cudaSuccess) {
		e = cudaMemcpyAsync(d_wah, bv->words, bv->n_words * 4, cudaMemcpyHostToDevice, t->stream);
	}
	if (e == cudaSuccess) {
		e = cudaMemcpyAsync(d_blk, block_group0.data(), n_blocks * 8, cudaMemcpyHostToDevice, t->stream);
	}
	if (e == cudaSuccess) {
		e = cudaMemsetAsync(dst, 0, t->words_per_bv * 8, t->stream);
	}
	if (e == cudaSuccess) {
		e = launch_wah_expand(d_wah, bv->n_words, d_blk, groups, bv->active_val, bv->active_nbits, dst, t->stream);
	}
	if (e == cudaSuccess) {
		e = cudaStreamSynchronize(t->stream);
	}
	cudaFree(d_wah);
	if (d_blk) {
		cudaFree(d_blk);
	}
	CU_TRY(e);
	t->launches++;
	ix->counts_valid = false;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_download_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *words,
                                            uint64_t n_words) {
	if (!t || !words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	if (n_words != t->n_words) {
		return fail(CUBIT_EINVAL, "n_words mismatch");
	}
	const uint64_t *src = ix->d_bits + (uint64_t)value_id * t->words_per_bv;
	CU_TRY(cudaMemcpyAsync(words, src, n_words * 8, cudaMemcpyDeviceToHost, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	return CUBIT_OK;
}

static int refresh_counts(cubit_gpu_table *t, Index *ix) {
	if (ix->counts_valid) {
		return CUBIT_OK;
	}
	if (t->scratch_n < ix->card) {
		if (t->d_scratch) {
			cudaFree(t->d_scratch);
			t->d_scratch = nullptr;
		}
		CU_TRY(cudaMalloc(&t->d_scratch, sizeof(unsigned long long) * ix->card));
		t->scratch_n = ix->card;
	}
	CU_TRY(launch_popcount_many(ix->d_bits, t->words_per_bv, ix->card, t->d_scratch, t->stream));
	t->launches += (ix->card + 32767) / 32768;
	static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
	CU_TRY(cudaMemcpyAsync(ix->counts.data(), t->d_scratch, sizeof(uint64_t) * ix->card, cudaMemcpyDeviceToHost,
	                       t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	ix->counts_valid = true;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_bitvector_count(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *count) {
	if (!t || !count) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
}

extern "C" int cubit_gpu_index_build(cubit_gpu_table *t, int32_t index_id, int32_t col_id, int64_t base_value) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
	for (Delta &d : ix->deltas) {
		if (d.n_ent) {
			return fail(CUBIT_ESTATE, "index %d has pending deltas; merge them before a rebuild", index_id);
		}
	}
	int launches = 0;
	CU_TRY(launch_index_build(c.d, c.elem, 0, t->n_rows, base_value, ix->card, ix->d_bits, t->words_per_bv,
	                          t->sm_count, t->stream, &launches));
	t->launches += launches;
	CU_TRY(cudaStreamSynchronize(t->stream));
	ix->counts_valid = false;
	ix->src_col = col_id;
	ix->src_base = base_value;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_set_delta(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const int64_t *rows,
                                   uint64_t n) {
	if (!t || (n && !rows)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix || value_id >= ix->card) {
		return fail(CUBIT_EINVAL, "bad (index %d, value %u)", index_id, value_id);
	}
	std::vector<int64_t> sorted(rows, rows + n);
	std::sort(sorted.begin(), sorted.end());
	if (n && (sorted.front() < 0 || (uint64_t)sorted.back() >= t->n_rows)) {
		return fail(CUBIT_EINVAL, "delta row out of range [0, %llu)", (unsigned long long)t->n_rows);
	}
	// CSR per segment of (word-in-segment, mask); a row listed twice cancels
	std::vector<uint32_t> off((size_t)t->n_seg + 1, 0);
	std::vector<DeltaEnt> ent;
	ent.reserve(n);
	uint64_t flipped = 0;
	size_t i = 0;
	while (i < sorted.size()) {
		const uint64_t gw = (uint64_t)sorted[i] / 64; // global word
		uint64_t mask = 0;
		while (i < sorted.size() && (uint64_t)sorted[i] / 64 == gw) {
			mask ^= 1ull << ((uint64_t)sorted[i] % 64);
			i++;
		}
		if (mask) {
			DeltaEnt e;
			e.word = (uint32_t)(gw % t->seg_words);
			e.pad = 0;
			e.mask = mask;
			ent.push_back(e);
			off[gw / t->seg_words + 1]++;
			flipped += (uint64_t)__builtin_popcountll(mask);
		}
	}
	for (size_t s = 0; s < t->n_seg; s++) {
		off[s + 1] += off[s];
	}
	if (ent.size() > 0xfffffff0ull) {
		return fail(CUBIT_EINVAL, "too many delta words");
	}
	CU_TRY(cudaStreamSynchronize(t->stream)); // no query may still read the old lists
	Delta &d = ix->deltas[value_id];
	free_delta(d);
	if (ent.empty()) {
		return CUBIT_OK;
	}
	CU_TRY(cudaMalloc(&d.d_off, off.size() * sizeof(uint32_t)));
	CU_TRY(cudaMalloc(&d.d_ent, ent.size() * sizeof(DeltaEnt)));
	CU_TRY(cudaMemcpyAsync(d.d_off, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, t->stream));
	CU_TRY(cudaMemcpyAsync(d.d_ent, ent.data(), ent.size() * sizeof(DeltaEnt), cudaMemcpyHostToDevice, t->stream));
	CU_TRY(cudaStreamSynchronize(t->stream));
	d.n_ent = ent.size();
	d.n_rows = flipped;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_merge_deltas(cubit_gpu_table *t, int32_t index_id) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	bool any = false;
	for (uint32_t v = 0; v < ix->card; v++) {
		Delta &d = ix->deltas[v];
		if (!d.n_ent) {
			continue;
		}
		CU_TRY(launch_apply_delta(ix->d_bits + (uint64_t)v * t->words_per_bv, d.d_off, d.d_ent, t->n_seg,
		                          t->seg_words, t->stream));
		t->launches++;
		any = true;
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	if (any) {
		for (Delta &d : ix->deltas) {
			free_delta(d);
		}
		ix->counts_valid = false;
	}
	return CUBIT_OK;
}

// ------------------------------------------------------------------ columns
extern "C" int cubit_gpu_upload_column(cubit_gpu_table *t, int32_t col_id, const void *data, uint32_t elem_bytes,
                                       uint64_t n) {
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
	std::lock_guard<std::mutex> lk(t->mu);
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
}

// NULLs of a column: its validity mask in the reference's layout (ValidityMask, validity_mask.hpp:50,163-168 —
// what a validity_uncompressed segment stores, validity_uncompressed.cpp:381).  The probe reports the validity
// of every projected value (cubit_gpu_fetch_validity) and aggregates skip NULL inputs.
extern "C" int cubit_gpu_upload_column_validity(cubit_gpu_table *t, int32_t col_id, const uint64_t *words,
                                                uint64_t n_words) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
}

// Decode the reference's on-disk column segments on the GPU (column_decode.cu).  Everything the kernel will
// dereference is bounds-checked here, on the host copy of the segment, so a malformed segment is an error
// return and never an out-of-bounds device access.
extern "C" int cubit_gpu_upload_column_segments(cubit_gpu_table *t, int32_t col_id, uint32_t elem_bytes,
                                                const cubit_column_segment *segs, uint32_t n_segs,
                                                cubit_decode_info *info) {
	if (!t || (!segs && n_segs)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (elem_bytes != 4 && elem_bytes != 8) {
		return fail(CUBIT_EINVAL, "elem_bytes must be 4 or 8");
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
	std::lock_guard<std::mutex> lk(t->mu);
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
	for (int b = 0; b < 2 && pe == cudaSuccess; b++) {
		if (!t->h_stage[b]) {
			pe = cudaMallocHost(reinterpret_cast<void **>(&t->h_stage[b]), chunk);
			if (pe == cudaSuccess) {
				pe = cudaEventCreateWithFlags(&t->stage_ev[b], cudaEventDisableTiming);
			}
		}
	}
	CU_TRY_CLEAN(pe);
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
}

extern "C" int cubit_gpu_download_column(cubit_gpu_table *t, int32_t col_id, void *data, uint32_t elem_bytes,
                                         uint64_t n) {
	if (!t || !data) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
}

extern "C" int cubit_gpu_synth_column(cubit_gpu_table *t, int32_t col_id, int32_t kind, uint64_t seed,
                                      uint64_t threshold, uint32_t card, uint32_t hot_lo, uint32_t hot_n) {
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
	std::lock_guard<std::mutex> lk(t->mu);
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
}

extern "C" int cubit_gpu_pack_column(cubit_gpu_table *t, int32_t col_id, int keep_raw, uint64_t *packed_bytes) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
	for (uint64_t b = 0; b < n_blk; b++) {
		hdr[b].base = base[b];
		hdr[b].width = width[b];
		if (off > 0xffffffffull) {
			return fail(CUBIT_EINVAL, "packed column exceeds 32 GiB");
		}
		hdr[b].word_off = (uint32_t)off;
		off += 16ull * width[b];
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
	if (!keep_raw) {
		cudaFree(c.d);
		c.d = nullptr;
	}
	if (packed_bytes) {
		*packed_bytes = bytes + n_blk * sizeof(PackHdr);
	}
	return CUBIT_OK;
}

// Append path (INSERT: new rows take the next row ids — DataTable::Append / BoundIndex::Append,
// src/include/duckdb/execution/index/bound_index.hpp:71-75; rowids are dense positions, row_group.cpp:511-514).
// Bitvectors are padded to whole segments, so appending inside the last segment touches no allocation; past it
// every index is re-strided once (capacity grows by half).  Indexes built from a column are extended on the GPU
// by the index-build kernel over the new rows only.
extern "C" int cubit_gpu_append_rows(cubit_gpu_table *t, uint64_t n_new, const cubit_append_column *cols,
                                     uint32_t n_cols) {
	if (!t || (!cols && n_cols)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (n_new == 0) {
		return fail(CUBIT_EINVAL, "n_new must be > 0");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
		if (it->second.packed()) {
			return fail(CUBIT_ESTATE, "append: column %d is bit-packed; appends need the raw form", cols[i].col_id);
		}
	}
	CU_TRY(cudaStreamSynchronize(t->stream)); // nothing in flight may still read the old allocations
	// ---- value bitvectors: re-stride when the new rows leave the padded last segment
	if (new_n_seg * t->seg_words > t->words_per_bv) {
		const uint64_t cap_seg = std::max<uint64_t>(new_n_seg, (uint64_t)(t->words_per_bv / t->seg_words) * 3 / 2 + 1);
		const uint64_t new_stride = cap_seg * t->seg_words;
		for (Index *ix : t->indexes) {
			uint64_t *nb = nullptr;
			const size_t bytes = (size_t)ix->card * new_stride * 8;
			CU_TRY(cudaMalloc(&nb, bytes));
			cudaError_t e = cudaMemsetAsync(nb, 0, bytes, t->stream);
			if (e == cudaSuccess) {
				e = cudaMemcpy2DAsync(nb, new_stride * 8, ix->d_bits, t->words_per_bv * 8, t->words_per_bv * 8, ix->card,
				                      cudaMemcpyDeviceToDevice, t->stream);
			}
			if (e == cudaSuccess) {
				e = cudaStreamSynchronize(t->stream);
			}
			if (e != cudaSuccess) {
				cudaFree(nb);
				CU_TRY(e);
			}
			cudaFree(ix->d_bits);
			ix->d_bits = nb;
		}
		t->words_per_bv = new_stride;
	}
	// ---- pending-delta CSR offsets cover [0, n_seg]: extend them with empty segments
	if (new_n_seg > t->n_seg) {
		for (Index *ix : t->indexes) {
			for (Delta &d : ix->deltas) {
				if (!d.d_off) {
					continue;
				}
				std::vector<uint32_t> off((size_t)new_n_seg + 1);
				CU_TRY(cudaMemcpy(off.data(), d.d_off, ((size_t)t->n_seg + 1) * 4, cudaMemcpyDeviceToHost));
				for (size_t s = (size_t)t->n_seg + 1; s <= new_n_seg; s++) {
					off[s] = off[t->n_seg];
				}
				uint32_t *no = nullptr;
				CU_TRY(cudaMalloc(&no, off.size() * 4));
				CU_TRY(cudaMemcpy(no, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
				cudaFree(d.d_off);
				d.d_off = no;
			}
		}
	}
	// ---- columns: grow, then the new rows host → device behind the old ones
	for (uint32_t i = 0; i < n_cols; i++) {
		Column &c = t->columns[cols[i].col_id];
		if (new_n > c.cap) {
			const uint64_t cap = std::max<uint64_t>(new_n, c.cap + c.cap / 2);
			void *nd = nullptr;
			CU_TRY(cudaMalloc(&nd, (size_t)cap * c.elem + 16));
			CU_TRY(cudaMemcpy(nd, c.d, (size_t)old_n * c.elem, cudaMemcpyDeviceToDevice));
			cudaFree(c.d);
			c.d = nd;
			c.cap = cap;
		}
		CU_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(c.d) + (size_t)old_n * c.elem, cols[i].data, (size_t)n_new * c.elem,
		                       cudaMemcpyHostToDevice, t->stream));
		c.n = new_n;
		if (c.d_valid) { // appended rows are valid until a new mask is uploaded
			const uint64_t old_w = (old_n + 63) / 64, new_w = (new_n + 63) / 64;
			if (new_w > c.valid_cap_words) {
				const uint64_t capw = std::max<uint64_t>(new_w, c.valid_cap_words + c.valid_cap_words / 2);
				unsigned long long *nv = nullptr;
				CU_TRY(cudaMalloc((void **)&nv, (capw + 2) * 8));
				CU_TRY(cudaMemcpy(nv, c.d_valid, old_w * 8, cudaMemcpyDeviceToDevice));
				cudaFree(c.d_valid);
				c.d_valid = nv;
				c.valid_cap_words = capw;
			}
			if (old_n & 63) {
				unsigned long long last = 0;
				CU_TRY(cudaMemcpy(&last, c.d_valid + old_w - 1, 8, cudaMemcpyDeviceToHost));
				last |= ~0ull << (old_n & 63);
				CU_TRY(cudaMemcpy(c.d_valid + old_w - 1, &last, 8, cudaMemcpyHostToDevice));
			}
			if (new_w > old_w) {
				CU_TRY(cudaMemset(c.d_valid + old_w, 0xff, (new_w - old_w) * 8));
			}
		}
	}
	t->n_rows = new_n;
	t->n_seg = (uint32_t)new_n_seg;
	t->n_words = (new_n + 63) / 64;
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
		CU_TRY(launch_index_build(it->second.d, it->second.elem, old_n, new_n, ix->src_base, ix->card, ix->d_bits,
		                          t->words_per_bv, t->sm_count, t->stream, &launches));
		t->launches += launches;
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	return CUBIT_OK;
}

extern "C" int cubit_gpu_drop_column(cubit_gpu_table *t, int32_t col_id) {
	if (!t) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
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
}

// -------------------------------------------------------------- persistence
// Image layout (little endian, every section 8-byte aligned):
//   ImageHeader | per value: ImageEntry, payload words, delta rows | u64 FNV-1a checksum of everything before
namespace {
struct ImageHeader {
	char magic[8]; // "CUBITIX1"
	uint64_t n_rows;
	uint32_t card;
	int32_t src_col; // column the index was built from, -1 = uploaded bitvectors
	int64_t src_base;
};
struct ImageEntry {
	uint32_t encoding; // 0 = verbatim 64-bit words, 1 = WAH 32-bit words
	uint32_t active_val;
	uint32_t active_nbits;
	uint32_t pad;
	uint64_t n_words;      // 64-bit words (verbatim) / 32-bit words (WAH)
	uint64_t n_delta_rows; // pending flipped rows that follow the payload
};
const char kImageMagic[8] = {'C', 'U', 'B', 'I', 'T', 'I', 'X', '1'};

uint64_t fnv1a(const uint8_t *p, uint64_t n) {
	uint64_t h = 1469598103934665603ull;
	for (uint64_t i = 0; i < n; i++) {
		h = (h ^ p[i]) * 1099511628211ull;
	}
	return h;
}

// WAH-compress a verbatim bitvector (row r = bit r%64 of word r/64) — host side of the persistence path.
// 31-bit groups are cut from a 64-bit window; a literal keeps the group's FIRST row in its most significant
// bit, hence the bit reversal.  Stops (returns false) as soon as the output would not be smaller than `limit`
// 32-bit words, so incompressible bitvectors cost one partial pass.
bool wah_compress(const uint64_t *words, uint64_t n_rows, uint64_t limit, std::vector<uint32_t> &out, uint32_t &active_val,
                  uint32_t &active_nbits) {
	out.clear();
	const uint64_t n_groups = n_rows / 31;
	auto bits_at = [&](uint64_t row, uint32_t n) -> uint32_t { // n ≤ 31 rows starting at `row`, LSB = first row
		const uint64_t w = row >> 6, sh = row & 63;
		uint64_t v = words[w] >> sh;
		if (sh + n > 64) {
			v |= words[w + 1] << (64 - sh);
		}
		return (uint32_t)(v & ((1ull << n) - 1ull));
	};
	auto reverse = [](uint32_t v, uint32_t n) -> uint32_t { // first row → most significant of n bits
		v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
		v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
		v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
		v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
		v = (v >> 16) | (v << 16);
		return v >> (32 - n);
	};
	for (uint64_t g = 0; g < n_groups; g++) {
		const uint32_t raw = bits_at(g * 31, 31);
		if (raw == 0 || raw == 0x7fffffffu) {
			const uint32_t fill = 0x80000000u | (raw ? 0x40000000u : 0u);
			if (!out.empty() && (out.back() & 0xc0000000u) == fill && (out.back() & 0x3fffffffu) < 0x3fffffffu) {
				out.back()++;
				continue;
			}
			out.push_back(fill | 1u);
		} else {
			out.push_back(reverse(raw, 31));
		}
		if (out.size() >= limit) {
			return false;
		}
	}
	active_nbits = (uint32_t)(n_rows % 31);
	active_val = active_nbits ? reverse(bits_at(n_groups * 31, active_nbits), active_nbits) : 0u;
	return true;
}
} // namespace

extern "C" void cubit_gpu_free_image(void *image) {
	free(image);
}

extern "C" int cubit_gpu_index_serialize(cubit_gpu_table *t, int32_t index_id, void **image, uint64_t *bytes) {
	if (!t || !image || !bytes) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*image = nullptr;
	*bytes = 0;
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	std::vector<uint8_t> img;
	auto append = [&](const void *p, size_t n) {
		const uint8_t *b = static_cast<const uint8_t *>(p);
		img.insert(img.end(), b, b + n);
		while (img.size() & 7) {
			img.push_back(0);
		}
	};
	ImageHeader h;
	memset(&h, 0, sizeof(h));
	memcpy(h.magic, kImageMagic, 8);
	h.n_rows = t->n_rows;
	h.card = ix->card;
	h.src_col = ix->src_col;
	h.src_base = ix->src_base;
	append(&h, sizeof(h));
	std::vector<uint64_t> words(t->n_words + 1, 0); // + 1: the group extractor may touch one word past the end
	std::vector<uint32_t> wah;
	std::vector<uint32_t> off;
	std::vector<DeltaEnt> ent;
	std::vector<int64_t> rows;
	for (uint32_t v = 0; v < ix->card; v++) {
		CU_TRY(cudaMemcpy(words.data(), ix->d_bits + (uint64_t)v * t->words_per_bv, t->n_words * 8, cudaMemcpyDeviceToHost));
		words[t->n_words] = 0;
		// pending deltas of this value: CSR (segment → (word, mask)) back to a flipped-row list
		rows.clear();
		const Delta &d = ix->deltas[v];
		if (d.n_ent) {
			off.resize((size_t)t->n_seg + 1);
			ent.resize(d.n_ent);
			CU_TRY(cudaMemcpy(off.data(), d.d_off, off.size() * 4, cudaMemcpyDeviceToHost));
			CU_TRY(cudaMemcpy(ent.data(), d.d_ent, d.n_ent * sizeof(DeltaEnt), cudaMemcpyDeviceToHost));
			for (uint32_t sgm = 0; sgm < t->n_seg; sgm++) {
				for (uint32_t e = off[sgm]; e < off[sgm + 1]; e++) {
					const uint64_t row0 = ((uint64_t)sgm * t->seg_words + ent[e].word) * 64;
					for (uint64_t m = ent[e].mask; m; m &= m - 1) {
						rows.push_back((int64_t)(row0 + (uint64_t)__builtin_ctzll(m)));
					}
				}
			}
		}
		ImageEntry e;
		memset(&e, 0, sizeof(e));
		e.n_delta_rows = rows.size();
		if (wah_compress(words.data(), t->n_rows, t->n_words * 2, wah, e.active_val, e.active_nbits)) {
			e.encoding = 1;
			e.n_words = wah.size();
			append(&e, sizeof(e));
			append(wah.data(), wah.size() * 4);
		} else {
			e.encoding = 0;
			e.active_val = e.active_nbits = 0;
			e.n_words = t->n_words;
			append(&e, sizeof(e));
			append(words.data(), t->n_words * 8);
		}
		append(rows.data(), rows.size() * 8);
	}
	const uint64_t sum = fnv1a(img.data(), img.size());
	append(&sum, 8);
	void *outp = malloc(img.size());
	if (!outp) {
		return fail(CUBIT_ENOMEM, "host allocation of %zu bytes failed", img.size());
	}
	memcpy(outp, img.data(), img.size());
	*image = outp;
	*bytes = img.size();
	return CUBIT_OK;
}

extern "C" int cubit_gpu_index_deserialize(cubit_gpu_table *t, const void *image, uint64_t bytes, int32_t *index_id) {
	if (!t || !image || !index_id) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	const uint8_t *p = static_cast<const uint8_t *>(image);
	if (bytes < sizeof(ImageHeader) + 8 || (bytes & 7)) {
		return fail(CUBIT_EINVAL, "index image: bad size %llu", (unsigned long long)bytes);
	}
	ImageHeader h;
	memcpy(&h, p, sizeof(h));
	if (memcmp(h.magic, kImageMagic, 8) != 0) {
		return fail(CUBIT_EINVAL, "index image: bad magic");
	}
	uint64_t sum;
	memcpy(&sum, p + bytes - 8, 8);
	if (sum != fnv1a(p, bytes - 8)) {
		return fail(CUBIT_EINVAL, "index image: checksum mismatch");
	}
	if (h.n_rows != t->n_rows) {
		return fail(CUBIT_EINVAL, "index image describes %llu rows, table has %llu", (unsigned long long)h.n_rows,
		            (unsigned long long)t->n_rows);
	}
	if (h.card == 0) {
		return fail(CUBIT_EINVAL, "index image: cardinality 0");
	}
	// pass 1: structure
	struct Sec {
		ImageEntry e;
		uint64_t payload, rows;
	};
	std::vector<Sec> secs;
	uint64_t at = sizeof(ImageHeader);
	const uint64_t end = bytes - 8;
	for (uint32_t v = 0; v < h.card; v++) {
		if (at + sizeof(ImageEntry) > end) {
			return fail(CUBIT_EINVAL, "index image: truncated at value %u", v);
		}
		Sec s;
		memcpy(&s.e, p + at, sizeof(ImageEntry));
		at += sizeof(ImageEntry);
		if (s.e.encoding > 1 || (s.e.encoding == 0 && s.e.n_words != t->n_words) || s.e.n_words > end ||
		    s.e.n_delta_rows > end) {
			return fail(CUBIT_EINVAL, "index image: bad entry for value %u", v);
		}
		const uint64_t pbytes = ((s.e.encoding ? s.e.n_words * 4 : s.e.n_words * 8) + 7) & ~7ull;
		if (at + pbytes + s.e.n_delta_rows * 8 > end) {
			return fail(CUBIT_EINVAL, "index image: truncated payload of value %u", v);
		}
		s.payload = at;
		s.rows = at + pbytes;
		at = s.rows + s.e.n_delta_rows * 8;
		secs.push_back(s);
	}
	if (at != end) {
		return fail(CUBIT_EINVAL, "index image: %llu trailing bytes", (unsigned long long)(end - at));
	}
	// pass 2: rebuild through the public entry points (each validates its input again)
	int32_t id = -1;
	int rc = cubit_gpu_index_create(t, h.card, &id);
	for (uint32_t v = 0; v < h.card && rc == CUBIT_OK; v++) {
		const Sec &s = secs[v];
		if (s.e.encoding == 1) {
			cubit_wah_bitvector bv;
			bv.words = reinterpret_cast<const uint32_t *>(p + s.payload);
			bv.n_words = s.e.n_words;
			bv.active_val = s.e.active_val;
			bv.active_nbits = s.e.active_nbits;
			rc = cubit_gpu_upload_bitvector_wah(t, id, v, &bv);
		} else {
			rc = cubit_gpu_upload_bitvector(t, id, v, reinterpret_cast<const uint64_t *>(p + s.payload), s.e.n_words);
		}
		if (rc == CUBIT_OK && s.e.n_delta_rows) {
			rc = cubit_gpu_set_delta(t, id, v, reinterpret_cast<const int64_t *>(p + s.rows), s.e.n_delta_rows);
		}
	}
	if (rc != CUBIT_OK) {
		return rc; // (the half-built index stays allocated until the table is destroyed; its id is not returned)
	}
	{
		std::lock_guard<std::mutex> lk(t->mu);
		Index *ix = get_index(t, id);
		if (ix) { // remember where the bitvectors came from, so appends keep extending the index on the GPU
			ix->src_col = h.src_col;
			ix->src_base = h.src_base;
		}
	}
	*index_id = id;
	return CUBIT_OK;
}

// -------------------------------------------------------------------- query
static cudaError_t run_scan(const ScanArgs &sa, uint32_t seg_words, bool has_delta, int sm_count, cudaStream_t st) {
	return launch_scan(sa, seg_words, has_delta, sm_count, st, nullptr);
}

static void release_result(cubit_gpu_result *r) {
	cudaStream_t s = r->stream;
	if (r->d_block) {
		cudaFreeAsync(r->d_block, s);
	}
	if (r->d_ids) {
		cudaFreeAsync(r->d_ids, s);
	}
	if (r->d_q) {
		cudaFreeAsync(r->d_q, s);
	}
	if (r->d_q_tmp) {
		cudaFreeAsync(r->d_q_tmp, s);
	}
	for (auto &p : r->d_vals) {
		if (p) {
			cudaFreeAsync(p, s);
		}
	}
	for (auto &p : r->d_valid) {
		if (p) {
			cudaFreeAsync(p, s);
		}
	}
	if (r->h_hdr) {
		r->t->hdr_pool.push_back(r->h_hdr); // caller holds t->mu
	}
	for (auto &e : r->ev) {
		if (e) {
			cudaEventDestroy(e);
		}
	}
	if (r->ev_done) {
		cudaEventDestroy(r->ev_done);
	}
	delete r;
}

static int finish_result(cubit_gpu_result *r) {
	if (r->finished) {
		return CUBIT_OK;
	}
	CU_TRY(cudaEventSynchronize(r->ev_done));
	r->info.count = r->h_hdr->count;
	r->info.sum_lo = r->h_hdr->sum_lo;
	r->info.sum_hi = r->h_hdr->sum_hi;
	r->info.sum_f64 = r->h_hdr->sum_f64;
	r->info.agg_rows = r->agg_kind == CUBIT_AGG_NONE ? 0 : (r->agg_nulls ? r->h_hdr->agg_rows : r->h_hdr->count);
	r->info.algo_bytes_scan += 8ull * ((r->flags & CUBIT_Q_ROWIDS) ? r->info.count : 0);
	// P of SURVEY §8d: M * Σ width over the distinct columns whose values are needed
	// (+ 8*M when a separate probe kernel re-reads the row IDs)
	r->info.algo_bytes_probe = r->info.count * r->probe_widths + r->probe_fixed_bytes;
	if (r->timing) {
		float ms = 0;
		cudaEventElapsedTime(&ms, r->ev[0], r->ev[1]);
		r->info.ms_scan = ms;
		if (r->probe_timed) {
			cudaEventElapsedTime(&ms, r->ev[1], r->ev[2]);
			r->info.ms_probe = ms;
			cudaEventElapsedTime(&ms, r->ev[0], r->ev[2]);
			r->info.ms_total = ms;
		} else {
			r->info.ms_total = r->info.ms_scan;
		}
	}
	r->finished = true;
	if (r->h_hdr->overflow) {
		return fail(CUBIT_EINVAL, "Overflow in multiplication of INT64 in SUM(a*b)");
	}
	if ((r->flags & (CUBIT_Q_ROWIDS | CUBIT_Q_VALUES)) && r->info.count > r->info.capacity) {
		return fail(CUBIT_ESTATE, "internal: result %llu exceeds capacity bound %llu",
		            (unsigned long long)r->info.count, (unsigned long long)r->info.capacity);
	}
	return CUBIT_OK;
}

extern "C" int cubit_gpu_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out) {
	if (!t || !q || !out) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*out = nullptr;
	if (q->n_groups == 0 || !q->groups) {
		return fail(CUBIT_EINVAL, "query has no predicate groups");
	}
	if (q->n_cols > CUBIT_MAX_PROBE_COLS || (q->n_cols && !q->cols)) {
		return fail(CUBIT_EINVAL, "bad projected column list");
	}
	if (q->agg_kind < CUBIT_AGG_NONE || q->agg_kind > CUBIT_AGG_SUM_F64) {
		return fail(CUBIT_EINVAL, "bad agg_kind %d", q->agg_kind);
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}

	// ---- flatten predicate into the ordered stream list
	ScanArgs sa;
	memset(&sa, 0, sizeof(sa));
	uint32_t k = 0;
	bool has_delta = false;
	uint64_t delta_entries = 0;
	uint64_t cap = t->n_rows;
	for (uint32_t g = 0; g < q->n_groups; g++) {
		const cubit_pred_group &grp = q->groups[g];
		if (grp.n_refs == 0 || !grp.refs) {
			return fail(CUBIT_EINVAL, "predicate group %u is empty", g);
		}
		uint64_t group_bound = 0;
		for (uint32_t i = 0; i < grp.n_refs; i++) {
			if (k >= (uint32_t)kMaxStreams) {
				return fail(CUBIT_EINVAL, "query reads more than %d bitvectors", kMaxStreams);
			}
			Index *ix = get_index(t, grp.refs[i].index_id);
			if (!ix || grp.refs[i].value_id >= ix->card) {
				return fail(CUBIT_EINVAL, "group %u ref %u: bad (index %d, value %u)", g, i, grp.refs[i].index_id,
				            grp.refs[i].value_id);
			}
			int rc = refresh_counts(t, ix);
			if (rc) {
				return rc;
			}
			const uint32_t v = grp.refs[i].value_id;
			sa.bv[k] = ix->d_bits + (uint64_t)v * t->words_per_bv;
			const Delta &d = ix->deltas[v];
			if (d.n_ent) {
				sa.doff[k] = d.d_off;
				sa.dent[k] = d.d_ent;
				has_delta = true;
				delta_entries += d.n_ent;
			}
			group_bound += ix->counts[v] + d.n_rows;
			k++;
		}
		sa.group_end |= 1ull << (k - 1);
		cap = std::min(cap, group_bound);
	}
	sa.k = k;
	if (const char *dbg = getenv("CUBIT_SCAN_DEBUG")) {
		sa.debug = (unsigned)atoi(dbg); // kernel timing experiments: skips parts of the kernel, results invalid
	}
	sa.n_seg = t->n_seg;
	sa.row_base = t->row_base;

	// ---- projected / aggregate columns
	const bool want_ids = (q->flags & CUBIT_Q_ROWIDS) != 0;
	const bool want_vals = (q->flags & CUBIT_Q_VALUES) != 0 && q->n_cols > 0;
	const bool want_q = (q->flags & CUBIT_Q_BITVECTOR) != 0;
	const bool unfused = (q->flags & CUBIT_Q_UNFUSED) != 0;
	const Column *vcols[CUBIT_MAX_PROBE_COLS] = {};
	bool fusable = !unfused;
	if (want_vals) {
		for (uint32_t c = 0; c < q->n_cols; c++) {
			auto it = t->columns.find(q->cols[c]);
			if (it == t->columns.end()) {
				return fail(CUBIT_EINVAL, "no column %d", q->cols[c]);
			}
			vcols[c] = &it->second;
			if (it->second.elem != 8) {
				fusable = false;
			}
		}
		if (q->n_cols > (uint32_t)kMaxFusedCols) {
			fusable = false;
		}
	}
	const Column *agg_a = nullptr, *agg_b = nullptr;
	if (q->agg_kind != CUBIT_AGG_NONE) {
		auto it = t->columns.find(q->agg_col_a);
		if (it == t->columns.end() || it->second.elem != 8) {
			return fail(CUBIT_EINVAL, "aggregate column %d missing or not 8 bytes wide", q->agg_col_a);
		}
		agg_a = &it->second;
		if (q->agg_kind == CUBIT_AGG_SUM_PROD) {
			it = t->columns.find(q->agg_col_b);
			if (it == t->columns.end() || it->second.elem != 8) {
				return fail(CUBIT_EINVAL, "aggregate column %d missing or not 8 bytes wide", q->agg_col_b);
			}
			agg_b = &it->second;
		}
	}
	// NULL-bearing columns are probed by the gather kernel over the row-ID list (validity gathered per
	// projected column, NULL inputs skipped by the aggregate); the bit-driven / fused paths assume no NULLs
	bool any_nulls = (agg_a && agg_a->d_valid) || (agg_b && agg_b->d_valid);
	for (uint32_t c = 0; want_vals && c < q->n_cols; c++) {
		any_nulls |= vcols[c]->d_valid != nullptr;
	}
	if (any_nulls) {
		fusable = false;
	}
	const bool need_probe = want_vals || q->agg_kind != CUBIT_AGG_NONE;
	// The scan-side probe paths gather at most kMaxFusedCols DISTINCT int64 columns per row.
	const Column *dist_cols[kMaxFusedCols] = {};
	int dist_out[kMaxFusedCols] = {-1, -1}; // which projected column each distinct column feeds
	int n_dist = 0, agg_ia = 0, agg_ib = 0;
	if (fusable && need_probe) {
		auto slot_of = [&](const Column *c) -> int {
			for (int d = 0; d < n_dist; d++) {
				if (dist_cols[d] == c) {
					return d;
				}
			}
			if (n_dist == kMaxFusedCols) {
				return -1;
			}
			dist_cols[n_dist] = c;
			return n_dist++;
		};
		if (want_vals) {
			for (uint32_t c = 0; c < q->n_cols && fusable; c++) {
				const int d = slot_of(vcols[c]);
				if (d < 0 || dist_out[d] >= 0) {
					fusable = false; // too many columns, or one column projected twice
				} else {
					dist_out[d] = (int)c;
				}
			}
		}
		if (fusable && agg_a) {
			agg_ia = slot_of(agg_a);
			fusable = agg_ia >= 0;
		}
		if (fusable && agg_b) {
			agg_ib = slot_of(agg_b);
			fusable = agg_ib >= 0;
		}
	}
	// How the probe runs:
	//   PROBE_BITS    (default) bit-driven probe kernel right after the scan kernel: re-decodes the
	//                 merged bitvector (1 bit/row instead of 8 bytes/selected row) as a plain fully
	//                 occupied grid — the gathers need far more loads in flight than the scan
	//                 kernel's 8 consumer warps per CTA can hold, and inside the scan kernel their
	//                 latency lands on the consumers' critical path (measured: profiles/)
	//   PROBE_FUSED   inside the scan kernel (CUBIT_Q_FUSE_PROBE): one launch
	//   PROBE_GATHER  gather kernel over the row-ID list — sparse selections whose row IDs are
	//                 materialised anyway (< 1/256 of the rows), 4-byte columns, > 2 columns, UNFUSED
	const bool dense_sel = cap > t->n_rows / 6; // upper bound of the selection (exact for disjoint ORs)
	enum { PROBE_NONE, PROBE_FUSED, PROBE_BITS, PROBE_GATHER } probe_mode = PROBE_NONE;
	if (need_probe) {
		if (!fusable) {
			probe_mode = PROBE_GATHER;
		} else {
			if (q->flags & CUBIT_Q_FUSE_PROBE) {
				probe_mode = PROBE_FUSED;
			} else if (cap <= t->n_rows / 256 && (want_ids || want_vals || want_q || k > 1 || has_delta)) {
				// sparse: gathering over the short row-ID list (materialised internally when the caller did
				// not ask for it: 8 bytes per selected row) beats writing + re-reading the N/8-byte bitvector
				// (measured: profiles/); a single clean bitvector is probed in place instead (probe_on_bv)
				probe_mode = PROBE_GATHER;
			} else {
				probe_mode = PROBE_BITS;
			}
		}
	}
	// Single value bitvector, no pending deltas, aggregate only (the equality-predicate + SUM query of config 1):
	// the merge is the identity, so the bit-driven probe reads B_v itself — no scan launch, no copy of Q —
	// and counts the set bits on the way.
	const bool probe_on_bv = probe_mode == PROBE_BITS && k == 1 && !has_delta && !want_ids && !want_vals && !want_q &&
	                         !unfused && sa.debug == 0;
	const bool separate_probe = probe_mode == PROBE_GATHER;
	const bool need_ids_buf = want_ids || separate_probe || (probe_mode == PROBE_BITS && want_vals);
	if (!need_ids_buf && !want_vals) {
		cap = 0;
	}
	cap = (cap + 1) & ~1ull; // even: the probe kernel moves row IDs in pairs

	// ---- result object
	cubit_gpu_result *r = new (std::nothrow) cubit_gpu_result();
	if (!r) {
		return fail(CUBIT_ENOMEM, "host allocation failed");
	}
	r->t = t;
	r->stream = t->stream;
	r->flags = q->flags;
	r->agg_kind = q->agg_kind;
	r->n_cols = want_vals ? q->n_cols : 0;
	r->timing = (q->flags & CUBIT_Q_TIMING) != 0;
	cudaStream_t st = t->stream;
	int rc = CUBIT_OK;
#define Q_TRY(expr)                                                                                                    \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			rc = fail(_e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "%s: %s (%s:%d)", #expr,           \
			          cudaGetErrorString(_e), __FILE__, __LINE__);                                                     \
			release_result(r);                                                                                         \
			return rc;                                                                                                 \
		}                                                                                                              \
	} while (0)

	const int max_grid = std::max(scan_max_grid(t->seg_words, t->sm_count), probe_grid(t->sm_count));
	const size_t hdr_bytes = 64;
	const size_t ctrl_bytes = ((size_t)t->n_seg + 1) * 8;
	const size_t ctrl_pad = (ctrl_bytes + 63) & ~(size_t)63;
	const size_t part_bytes = (size_t)max_grid * sizeof(BlockPartial);
	// layout: hdr | ctrl A | ctrl B (decode pass of the unfused path) | partials | probe done | segment prefixes
	const size_t excl_bytes = probe_mode == PROBE_BITS ? ctrl_pad : 0;
	const size_t block_bytes = hdr_bytes + 2 * ctrl_pad + part_bytes + 64 + excl_bytes;
	Q_TRY(cudaMallocAsync((void **)&r->d_block, block_bytes, st));
	r->d_hdr = reinterpret_cast<ResultHeader *>(r->d_block);
	unsigned long long *ctrl_a = reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes);
	unsigned long long *ctrl_b = reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes + ctrl_pad);
	BlockPartial *partials = reinterpret_cast<BlockPartial *>(r->d_block + hdr_bytes + 2 * ctrl_pad);
	unsigned int *probe_done = reinterpret_cast<unsigned int *>(r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes);
	unsigned long long *tile_excl =
	    reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes + 64);
	if (!t->hdr_pool.empty()) {
		r->h_hdr = t->hdr_pool.back();
		t->hdr_pool.pop_back();
	} else {
		Q_TRY(cudaHostAlloc((void **)&r->h_hdr, sizeof(ResultHeader), cudaHostAllocDefault));
	}
	memset(r->h_hdr, 0, sizeof(ResultHeader));
	if (need_ids_buf && cap) {
		Q_TRY(cudaMallocAsync((void **)&r->d_ids, cap * 8, st));
	}
	if (want_vals && cap) {
		for (uint32_t c = 0; c < q->n_cols; c++) {
			r->val_elem[c] = vcols[c]->elem;
			Q_TRY(cudaMallocAsync(&r->d_vals[c], cap * vcols[c]->elem, st));
		}
	}
	if (want_q) {
		Q_TRY(cudaMallocAsync((void **)&r->d_q, t->words_per_bv * 8, st));
	}
	if ((unfused || probe_mode == PROBE_BITS) && !want_q && !probe_on_bv) {
		Q_TRY(cudaMallocAsync((void **)&r->d_q_tmp, t->words_per_bv * 8, st));
	}
	Q_TRY(cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming));
	if (r->timing) {
		for (auto &e : r->ev) {
			Q_TRY(cudaEventCreate(&e));
		}
	}

	// zero hdr + both control blocks (ticket counters and per-segment status words)
	Q_TRY(cudaMemsetAsync(r->d_block, 0, hdr_bytes + 2 * ctrl_pad, st));

	sa.partials = partials;
	sa.hdr = r->d_hdr;
	sa.ids_cap = cap;
	uint32_t n_launch = 0;
	if (r->timing) {
		Q_TRY(cudaEventRecord(r->ev[0], st));
	}
	if (probe_on_bv) {
		sa.q_out = const_cast<uint64_t *>(sa.bv[0]);
		r->info.fused = 1;
	} else if (!unfused) {
		// one pass: merge (+delta XOR) + decode (+ fused probe / aggregate when eligible)
		sa.ctrl = ctrl_a;
		sa.q_out = probe_mode == PROBE_BITS && !want_q ? r->d_q_tmp : r->d_q;
		sa.ids_out = need_ids_buf ? r->d_ids : nullptr;
		sa.tile_excl = probe_mode == PROBE_BITS && want_vals ? tile_excl : nullptr;
		if (probe_mode == PROBE_FUSED) {
			sa.n_load = n_dist;
			for (int d = 0; d < n_dist; d++) {
				sa.lcol[d] = col_ref(dist_cols[d], dense_sel);
				sa.lout[d] = dist_out[d] >= 0 && cap ? static_cast<long long *>(r->d_vals[dist_out[d]]) : nullptr;
			}
			sa.agg_kind = q->agg_kind;
			sa.agg_ia = agg_ia;
			sa.agg_ib = agg_ib;
		}
		Q_TRY(run_scan(sa, t->seg_words, has_delta, t->sm_count, st));
		n_launch++;
		r->info.fused = 1;
	} else {
		// three separate kernels: K1 merge → Q, K2 decode Q → row IDs, K3 probe
		uint64_t *qbuf = want_q ? r->d_q : r->d_q_tmp;
		sa.ctrl = ctrl_a;
		sa.q_out = qbuf;
		sa.ids_out = nullptr;
		Q_TRY(run_scan(sa, t->seg_words, has_delta, t->sm_count, st));
		n_launch++;
		if (need_ids_buf) {
			ScanArgs sd;
			memset(&sd, 0, sizeof(sd));
			sd.bv[0] = qbuf;
			sd.group_end = 1;
			sd.k = 1;
			sd.n_seg = t->n_seg;
			sd.row_base = t->row_base;
			sd.ctrl = ctrl_b;
			sd.ids_out = r->d_ids;
			sd.ids_cap = cap;
			sd.partials = partials;
			sd.hdr = r->d_hdr;
			sd.skip_count = 1; // K1 already counted the selection
			Q_TRY(run_scan(sd, t->seg_words, false, t->sm_count, st));
			n_launch++;
		}
		r->info.fused = 0;
	}
	if (r->timing) {
		Q_TRY(cudaEventRecord(r->ev[1], st));
	}
	if (probe_mode == PROBE_BITS) {
		ScanArgs pb;
		memset(&pb, 0, sizeof(pb));
		pb.q_out = sa.q_out; // input of the bit-driven probe
		pb.tile_excl = sa.tile_excl;
		pb.n_seg = t->n_seg;
		pb.row_base = t->row_base;
		pb.ids_cap = cap;
		pb.n_load = n_dist;
		for (int d = 0; d < n_dist; d++) {
			pb.lcol[d] = col_ref(dist_cols[d], dense_sel);
			pb.lout[d] = dist_out[d] >= 0 && cap ? static_cast<long long *>(r->d_vals[dist_out[d]]) : nullptr;
		}
		pb.agg_kind = q->agg_kind;
		pb.agg_ia = agg_ia;
		pb.agg_ib = agg_ib;
		pb.hdr = r->d_hdr;
		pb.count_rows = probe_on_bv ? 1 : 0;
		Q_TRY(launch_probe_bits(pb, t->seg_words, want_vals && cap, t->sm_count, st));
		n_launch++;
		if (r->timing) {
			Q_TRY(cudaEventRecord(r->ev[2], st));
			r->probe_timed = true;
		}
	}
	if (separate_probe) {
		ProbeArgs pa;
		memset(&pa, 0, sizeof(pa));
		pa.ids = r->d_ids;
		pa.count_ptr = &r->d_hdr->count;
		pa.row_base = t->row_base;
		pa.n_cols = want_vals && cap ? (int)q->n_cols : 0;
		for (int c = 0; c < pa.n_cols; c++) {
			pa.col[c] = vcols[c]->d; // the gather over row IDs prefers the raw form when it is resident: one
			pa.packed[c] = col_ref(vcols[c], true); // dependent load per value instead of header + payload
			pa.out[c] = r->d_vals[c];
			pa.elem_bytes[c] = vcols[c]->elem;
		}
		pa.agg_kind = q->agg_kind;
		pa.agg_a = col_ref(agg_a, true);
		pa.agg_b = col_ref(agg_b, true);
		pa.agg_valid_a = agg_a ? agg_a->d_valid : nullptr;
		pa.agg_valid_b = agg_b ? agg_b->d_valid : nullptr;
		r->agg_nulls = pa.agg_valid_a || pa.agg_valid_b;
		pa.partials = partials;
		pa.done = probe_done;
		pa.hdr = r->d_hdr;
		Q_TRY(launch_probe(pa, t->sm_count, st));
		n_launch++;
		for (int c = 0; c < pa.n_cols; c++) {
			if (!vcols[c]->d_valid) {
				continue;
			}
			const size_t vbytes = ((size_t)(cap + 31) / 32 + 2) * 4;
			Q_TRY(cudaMallocAsync((void **)&r->d_valid[c], vbytes, st));
			Q_TRY(cudaMemsetAsync(r->d_valid[c], 0, vbytes, st));
			Q_TRY(launch_validity_gather(r->d_ids, &r->d_hdr->count, t->row_base, vcols[c]->d_valid, r->d_valid[c],
			                             t->sm_count, st));
			n_launch++;
		}
		if (r->timing) {
			Q_TRY(cudaEventRecord(r->ev[2], st));
			r->probe_timed = true;
		}
	}
	Q_TRY(cudaMemcpyAsync(r->h_hdr, r->d_hdr, sizeof(ResultHeader), cudaMemcpyDeviceToHost, st));
	Q_TRY(cudaEventRecord(r->ev_done, st));
#undef Q_TRY
	t->launches += n_launch;

	{
		std::vector<int32_t> seen;
		auto add_col = [&](int32_t id, uint32_t w) {
			if (std::find(seen.begin(), seen.end(), id) == seen.end()) {
				seen.push_back(id);
				r->probe_widths += w;
			}
		};
		for (uint32_t c = 0; c < r->n_cols; c++) {
			add_col(q->cols[c], vcols[c]->elem);
		}
		if (q->agg_kind != CUBIT_AGG_NONE) {
			add_col(q->agg_col_a, 8);
		}
		if (q->agg_kind == CUBIT_AGG_SUM_PROD) {
			add_col(q->agg_col_b, 8);
		}
		if (separate_probe) {
			r->probe_widths += 8; // the gather kernel re-reads the 8-byte row IDs
		}
		r->probe_fixed_bytes = probe_mode == PROBE_BITS ? t->n_words * 8 : 0; // ... the bit-driven one re-reads Q
	}
	r->info.capacity = cap;
	r->info.n_streams = k;
	r->info.n_launches = n_launch;
	r->info.delta_entries = delta_entries;
	// (probe_on_bv: the one bitvector is read once, by the probe — accounted in probe_fixed_bytes)
	r->info.algo_bytes_scan = probe_on_bv ? 0 : (uint64_t)k * t->n_words * 8 + delta_entries * sizeof(DeltaEnt);
	r->info.d_rowids = want_ids ? reinterpret_cast<const int64_t *>(r->d_ids) : nullptr;
	r->info.d_bitvector = r->d_q;
	for (uint32_t c = 0; c < r->n_cols; c++) {
		r->info.d_values[c] = r->d_vals[c];
		r->info.d_validity[c] = r->d_valid[c];
	}
	if (!(q->flags & CUBIT_Q_ASYNC)) {
		rc = finish_result(r);
		if (rc) {
			release_result(r);
			return rc;
		}
	}
	*out = r;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_result_wait(cubit_gpu_result *r) {
	if (!r) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	cudaSetDevice(r->t->device);
	return finish_result(r);
}

extern "C" int cubit_gpu_result_get(cubit_gpu_result *r, cubit_result_info *info) {
	if (!r || !info) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	cudaSetDevice(r->t->device);
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	*info = r->info;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids,
                               uint32_t n_cols, void *const *host_cols) {
	if (!r) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	cudaSetDevice(r->t->device);
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (offset > r->info.count || n > r->info.count - offset) {
		return fail(CUBIT_EINVAL, "fetch range [%llu, +%llu) outside result of %llu rows", (unsigned long long)offset,
		            (unsigned long long)n, (unsigned long long)r->info.count);
	}
	if (n_cols > r->n_cols) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, %u requested", r->n_cols, n_cols);
	}
	if (host_rowids && !(r->flags & CUBIT_Q_ROWIDS)) {
		return fail(CUBIT_ESTATE, "query did not materialise row IDs (CUBIT_Q_ROWIDS)");
	}
	if (n == 0) {
		return CUBIT_OK;
	}
	std::lock_guard<std::mutex> lk(r->t->mu);
	cudaStream_t st = r->stream;
	if (host_rowids) {
		CU_TRY(cudaMemcpyAsync(host_rowids, r->d_ids + offset, n * 8, cudaMemcpyDeviceToHost, st));
	}
	for (uint32_t c = 0; c < n_cols; c++) {
		if (host_cols && host_cols[c]) {
			const size_t w = r->val_elem[c];
			CU_TRY(cudaMemcpyAsync(host_cols[c], static_cast<const char *>(r->d_vals[c]) + offset * w, n * w,
			                       cudaMemcpyDeviceToHost, st));
		}
	}
	CU_TRY(cudaStreamSynchronize(st));
	return CUBIT_OK;
}

// Validity of projected column `col` for result rows [offset, offset + n): bit j of host_words = row offset + j
// (1 = valid), ceil(n / 64) words, bits past n zero — the mask a DataChunk vector carries (vector.hpp:242-256).
extern "C" int cubit_gpu_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n,
                                        uint64_t *host_words, int *all_valid) {
	if (!r || (!host_words && !all_valid)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	cudaSetDevice(r->t->device);
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (col >= r->n_cols) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, column %u requested", r->n_cols, col);
	}
	if (offset > r->info.count || n > r->info.count - offset) {
		return fail(CUBIT_EINVAL, "fetch range [%llu, +%llu) outside result of %llu rows", (unsigned long long)offset,
		            (unsigned long long)n, (unsigned long long)r->info.count);
	}
	const uint64_t out_words = (n + 63) / 64;
	if (!r->d_valid[col]) { // the column holds no NULLs
		if (all_valid) {
			*all_valid = 1;
		}
		for (uint64_t w = 0; host_words && w < out_words; w++) {
			const uint64_t left = n - w * 64;
			host_words[w] = left >= 64 ? ~0ull : ((1ull << left) - 1);
		}
		return CUBIT_OK;
	}
	if (n == 0) {
		if (all_valid) {
			*all_valid = 1;
		}
		return CUBIT_OK;
	}
	// device mask is 32-bit words over result positions (zero past count, 2 spare words): copy the covering
	// 64-bit words and shift so that bit 0 = row `offset`
	const uint64_t w0 = offset / 64, sh = offset % 64;
	const uint64_t src_words = (sh + n + 63) / 64;
	std::vector<uint64_t> tmp(src_words + 1, 0);
	{
		std::lock_guard<std::mutex> lk(r->t->mu);
		const uint64_t avail32 = (r->info.capacity + 31) / 32 + 2; // words allocated
		uint64_t copy32 = src_words * 2;
		if (w0 * 2 + copy32 > avail32) {
			copy32 = avail32 - w0 * 2;
		}
		CU_TRY(cudaMemcpyAsync(tmp.data(), r->d_valid[col] + w0 * 2, copy32 * 4, cudaMemcpyDeviceToHost, r->stream));
		CU_TRY(cudaStreamSynchronize(r->stream));
	}
	bool all = true;
	for (uint64_t w = 0; w < out_words; w++) {
		uint64_t v = tmp[w] >> sh;
		if (sh) {
			v |= tmp[w + 1] << (64 - sh);
		}
		const uint64_t left = n - w * 64;
		const uint64_t mask = left >= 64 ? ~0ull : ((1ull << left) - 1);
		v &= mask;
		all &= v == mask;
		if (host_words) {
			host_words[w] = v;
		}
	}
	if (all_valid) {
		*all_valid = all ? 1 : 0;
	}
	return CUBIT_OK;
}

extern "C" int cubit_gpu_alloc_host(uint64_t bytes, void **ptr) {
	if (!ptr) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*ptr = nullptr;
	cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
		return fail(CUBIT_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
	}
	if (e != cudaSuccess) {
		return fail(e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "cudaHostAlloc(%llu): %s",
		            (unsigned long long)bytes, cudaGetErrorString(e));
	}
	return CUBIT_OK;
}

extern "C" int cubit_gpu_free_host(void *ptr) {
	if (ptr) {
		cudaFreeHost(ptr);
	}
	return CUBIT_OK;
}

extern "C" int cubit_gpu_fetch_bitvector(cubit_gpu_result *r, uint64_t *host_words, uint64_t n_words) {
	if (!r || !host_words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	cudaSetDevice(r->t->device);
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (!r->d_q) {
		return fail(CUBIT_ESTATE, "query did not materialise the bitvector (CUBIT_Q_BITVECTOR)");
	}
	if (n_words != r->t->n_words) {
		return fail(CUBIT_EINVAL, "n_words mismatch");
	}
	std::lock_guard<std::mutex> lk(r->t->mu);
	CU_TRY(cudaMemcpyAsync(host_words, r->d_q, n_words * 8, cudaMemcpyDeviceToHost, r->stream));
	CU_TRY(cudaStreamSynchronize(r->stream));
	return CUBIT_OK;
}

extern "C" int cubit_gpu_free_result(cubit_gpu_result *r) {
	if (!r) {
		return CUBIT_OK;
	}
	cudaSetDevice(r->t->device);
	std::lock_guard<std::mutex> lk(r->t->mu);
	release_result(r);
	return CUBIT_OK;
}

// -------------------------------------------------------------------- probe
extern "C" int cubit_gpu_probe(cubit_gpu_table *t, int32_t col_id, const int64_t *host_rowids, uint64_t n,
                               void *host_out, uint64_t *sum_lo, int64_t *sum_hi) {
	if (!t || (n && !host_rowids)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	std::lock_guard<std::mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	auto it = t->columns.find(col_id);
	if (it == t->columns.end()) {
		return fail(CUBIT_EINVAL, "no column %d", col_id);
	}
	const Column &c = it->second;
	const bool want_sum = sum_lo || sum_hi;
	if (want_sum && c.elem != 8) {
		return fail(CUBIT_EINVAL, "SUM needs an 8-byte column");
	}
	for (uint64_t i = 0; i < n; i++) {
		const int64_t l = host_rowids[i] - t->row_base;
		if (l < 0 || (uint64_t)l >= t->n_rows) {
			return fail(CUBIT_EINVAL, "row id %lld outside this shard", (long long)host_rowids[i]);
		}
	}
	if (sum_lo) {
		*sum_lo = 0;
	}
	if (sum_hi) {
		*sum_hi = 0;
	}
	if (n == 0) {
		return CUBIT_OK;
	}
	cudaStream_t st = t->stream;
	const uint64_t cap = (n + 1) & ~1ull;
	long long *d_ids = nullptr;
	void *d_out = nullptr;
	unsigned char *d_blk = nullptr;
	const int grid = probe_grid(t->sm_count);
	const size_t blk_bytes = 64 + 64 + (size_t)grid * sizeof(BlockPartial);
	CU_TRY(cudaMallocAsync((void **)&d_ids, cap * 8, st));
	CU_TRY(cudaMallocAsync(&d_out, cap * c.elem, st));
	CU_TRY(cudaMallocAsync((void **)&d_blk, blk_bytes, st));
	CU_TRY(cudaMemsetAsync(d_blk, 0, 128, st));
	CU_TRY(cudaMemcpyAsync(d_ids, host_rowids, n * 8, cudaMemcpyHostToDevice, st));
	ProbeArgs pa;
	memset(&pa, 0, sizeof(pa));
	pa.ids = d_ids;
	pa.n = n;
	pa.row_base = t->row_base;
	pa.n_cols = host_out ? 1 : 0;
	pa.col[0] = c.d;
	pa.packed[0] = col_ref(&c, true);
	pa.out[0] = d_out;
	pa.elem_bytes[0] = c.elem;
	pa.agg_kind = want_sum ? CUBIT_AGG_SUM : CUBIT_AGG_NONE;
	pa.agg_a = col_ref(&c, true);
	pa.hdr = reinterpret_cast<ResultHeader *>(d_blk);
	pa.done = reinterpret_cast<unsigned int *>(d_blk + 64);
	pa.partials = reinterpret_cast<BlockPartial *>(d_blk + 128);
	CU_TRY(launch_probe(pa, t->sm_count, st));
	t->launches++;
	ResultHeader h;
	if (host_out) {
		CU_TRY(cudaMemcpyAsync(host_out, d_out, n * c.elem, cudaMemcpyDeviceToHost, st));
	}
	CU_TRY(cudaMemcpyAsync(&h, d_blk, sizeof(h), cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	cudaFreeAsync(d_ids, st);
	cudaFreeAsync(d_out, st);
	cudaFreeAsync(d_blk, st);
	if (sum_lo) {
		*sum_lo = h.sum_lo;
	}
	if (sum_hi) {
		*sum_hi = h.sum_hi;
	}
	return CUBIT_OK;
}
