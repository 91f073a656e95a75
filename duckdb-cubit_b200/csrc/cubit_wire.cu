// cubit_wire.cu — the narrow-wire DataChunk hand-off (format: include/cubit_gpu_wire.h).
//
// What it replaces: the ≤ 2048-row vectors a scan hands to PhysicalTableScan::GetData
// (src/function/table/table_scan.cpp:251-273), produced by several workers in batch-index order
// (table_scan.cpp:179-189, table_function.hpp:45-67).  The wide hand-off (cubit_gpu_fetch*) moves 8 bytes per
// row ID and value over PCIe — 55 GB/s on this box, the whole cost of a row-returning query.  Here the GPU
// re-encodes every (stream, DataChunk) as base + narrow deltas and stores them straight into the caller's
// page-locked window (zero-copy: no staging buffer, no second pass, no host round trip to learn a size), and the
// consumer widens one chunk at a time into its DataChunk — in cache, right before the next operator reads it.
//
//   cubit_wire_pack_kernel      one CTA per (chunk, stream): coalesced load of 2048 values, block min / max, width
//                               choice, deltas staged in shared memory, contiguous 16-byte stores to host memory
//   cubit_gpu_fetch_wire_async  enqueue it on a copy stream behind the query
//   cubit_gpu_drain             n worker threads: claim window → two wires in flight → unpack chunk → callback
#include "table.h"
#include "../../include/cubit_gpu_wire.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

using namespace cubit;

namespace {

constexpr int kPackThreads = 256;
constexpr int kPerThread = CUBIT_WIRE_CHUNK / kPackThreads; // 8

struct WireArgs {
	const void *src[CUBIT_MAX_PROBE_COLS + 1]; // first value of the window, per stream
	uint32_t elem[CUBIT_MAX_PROBE_COLS + 1];   // 4 or 8
	uint64_t n_rows;
	uint64_t n_chunks;
	cubit_wire_dir *dir; // device-visible address of the wire's directory
	unsigned char *slots;
};

__device__ __forceinline__ long long shfl_xor_ll(long long v, int m) {
	return __shfl_xor_sync(0xffffffffu, v, m);
}

__global__ void __launch_bounds__(kPackThreads) cubit_wire_pack_kernel(const WireArgs a) {
	__shared__ __align__(16) unsigned char stage[CUBIT_WIRE_SLOT_BYTES];
	__shared__ long long s_min[kPackThreads / 32], s_max[kPackThreads / 32];
	const uint32_t chunk = blockIdx.x, stream = blockIdx.y, t = threadIdx.x;
	const uint64_t row0 = (uint64_t)chunk * CUBIT_WIRE_CHUNK;
	const uint32_t n = (uint32_t)min((uint64_t)CUBIT_WIRE_CHUNK, a.n_rows - row0);
	long long v[kPerThread];
	long long lo = LLONG_MAX, hi = LLONG_MIN;
	if (a.elem[stream] == 8) {
		const long long *s = static_cast<const long long *>(a.src[stream]) + row0;
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			const uint32_t i = t + j * kPackThreads;
			v[j] = i < n ? __ldcs(s + i) : 0;
		}
	} else {
		const int *s = static_cast<const int *>(a.src[stream]) + row0;
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			const uint32_t i = t + j * kPackThreads;
			v[j] = i < n ? (long long)__ldcs(s + i) : 0;
		}
	}
#pragma unroll
	for (int j = 0; j < kPerThread; j++) {
		if (t + j * kPackThreads < n) {
			lo = min(lo, v[j]);
			hi = max(hi, v[j]);
		}
	}
#pragma unroll
	for (int m = 16; m; m >>= 1) {
		lo = min(lo, shfl_xor_ll(lo, m));
		hi = max(hi, shfl_xor_ll(hi, m));
	}
	if ((t & 31) == 0) {
		s_min[t >> 5] = lo;
		s_max[t >> 5] = hi;
	}
	__syncthreads();
#pragma unroll
	for (int w = 0; w < kPackThreads / 32; w++) {
		lo = min(lo, s_min[w]);
		hi = max(hi, s_max[w]);
	}
	const unsigned long long range = (unsigned long long)hi - (unsigned long long)lo; // true difference, < 2^64
	const uint32_t width = range == 0 ? 0 : range < 256ull ? 1 : range < 65536ull ? 2 : range < (1ull << 32) ? 4 : 8;
#pragma unroll
	for (int j = 0; j < kPerThread; j++) {
		const uint32_t i = t + j * kPackThreads;
		const unsigned long long d = (unsigned long long)v[j] - (unsigned long long)lo;
		if (i < n) {
			switch (width) {
			case 1:
				stage[i] = (unsigned char)d;
				break;
			case 2:
				reinterpret_cast<unsigned short *>(stage)[i] = (unsigned short)d;
				break;
			case 4:
				reinterpret_cast<unsigned int *>(stage)[i] = (unsigned int)d;
				break;
			case 8:
				reinterpret_cast<unsigned long long *>(stage)[i] = d;
				break;
			default:
				break;
			}
		}
	}
	__syncthreads();
	const uint64_t slot = (uint64_t)stream * a.n_chunks + chunk;
	// contiguous 16-byte stores: a warp writes 512 consecutive bytes of host memory per instruction (the tail of the
	// last 16-byte unit is shared-memory garbage the unpacker never reads)
	const uint32_t units = (n * width + 15) / 16;
	uint4 *dst = reinterpret_cast<uint4 *>(a.slots + slot * CUBIT_WIRE_SLOT_BYTES);
	const uint4 *src16 = reinterpret_cast<const uint4 *>(stage);
	for (uint32_t u = t; u < units; u += kPackThreads) {
		dst[u] = src16[u];
	}
	if (t == 0) {
		uint4 e;
		e.x = (unsigned int)(unsigned long long)lo;
		e.y = (unsigned int)((unsigned long long)lo >> 32);
		e.z = width;
		e.w = n;
		reinterpret_cast<uint4 *>(a.dir)[slot] = e;
	}
}

// page-locked wire windows recycled across drains (cudaHostAlloc costs milliseconds)
void *wire_pool_acquire(cubit_gpu_table *t, uint64_t bytes) {
	{
		std::lock_guard<std::mutex> ml(t->meta_mu);
		for (size_t i = 0; i < t->wire_pool.size(); i++) {
			if (t->wire_pool[i].second >= bytes) {
				void *p = t->wire_pool[i].first;
				t->wire_pool[i] = t->wire_pool.back();
				t->wire_pool.pop_back();
				return p;
			}
		}
	}
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

struct PooledWire {
	void *p = nullptr;
	uint64_t bytes = 0;
};

} // namespace

namespace cubit {

void wire_pool_free(cubit_gpu_table *t) {
	std::lock_guard<std::mutex> ml(t->meta_mu);
	for (auto &w : t->wire_pool) {
		cudaFreeHost(w.first);
	}
	t->wire_pool.clear();
}

} // namespace cubit

extern "C" uint64_t cubit_gpu_wire_bytes(uint64_t n_rows, uint32_t n_streams) {
	return cubit_wire_bytes(n_rows, n_streams);
}

extern "C" uint64_t cubit_gpu_wire_payload_bytes(const void *host_wire) {
	return host_wire ? cubit_wire_payload_bytes(host_wire) : 0;
}

extern "C" int cubit_gpu_wire_unpack(const void *host_wire, uint32_t stream, uint64_t chunk, void *out,
                                     uint32_t out_elem) {
	if (!host_wire || !out) {
		fail(CUBIT_EINVAL, "NULL argument");
		return -1;
	}
	const int n = cubit_wire_unpack_chunk(host_wire, stream, chunk, out, out_elem);
	if (n < 0) {
		fail(CUBIT_EINVAL, "malformed wire, or stream / element size out of range");
	}
	return n;
}

extern "C" int cubit_gpu_fetch_wire_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int with_rowids,
                                          uint32_t n_cols, void *host_wire, uint64_t host_wire_bytes,
                                          cubit_gpu_fetch_ticket **ticket) {
	ABI_BEGIN
	if (!r || !ticket || !host_wire) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*ticket = nullptr;
	cubit_result_info info;
	int rc = cubit_gpu_result_get(r, &info); // waits for the query (outside the table lock)
	if (rc) {
		return rc;
	}
	if (!r->parts.empty()) {
		return fail(CUBIT_ESTATE, "a sharded result has no single device to write a wire from: use cubit_gpu_drain "
		                          "or cubit_gpu_fetch_async");
	}
	if (offset > info.count || n > info.count - offset) {
		return fail(CUBIT_EINVAL, "fetch range [%llu, +%llu) outside result of %llu rows", (unsigned long long)offset,
		            (unsigned long long)n, (unsigned long long)info.count);
	}
	if (n_cols > r->n_cols) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, %u requested", r->n_cols, n_cols);
	}
	if (with_rowids && !(r->flags & CUBIT_Q_ROWIDS)) {
		return fail(CUBIT_ESTATE, "query did not materialise row IDs (CUBIT_Q_ROWIDS)");
	}
	const uint32_t n_streams = (with_rowids ? 1u : 0u) + n_cols;
	if (n_streams == 0) {
		return fail(CUBIT_EINVAL, "a wire needs at least one stream");
	}
	const uint64_t need = cubit_wire_bytes(n, n_streams);
	if (host_wire_bytes < need) {
		return fail(CUBIT_EINVAL, "wire buffer of %llu bytes, %llu needed (cubit_gpu_wire_bytes)",
		            (unsigned long long)host_wire_bytes, (unsigned long long)need);
	}
	if (reinterpret_cast<uintptr_t>(host_wire) % 16 != 0) {
		return fail(CUBIT_EINVAL, "wire buffer must be 16-byte aligned");
	}
	cubit_gpu_table *t = r->t;
	CU_TRY(cudaSetDevice(t->device));
	// the device stores into the buffer directly: it has to be page-locked, mapped memory
	cudaPointerAttributes attr;
	cudaError_t pe = cudaPointerGetAttributes(&attr, host_wire);
	if (pe != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
		cudaGetLastError();
		return fail(CUBIT_EINVAL, "wire buffer must be page-locked host memory (cubit_gpu_alloc_host)");
	}
	unsigned char *dev_wire = static_cast<unsigned char *>(attr.devicePointer);
	cubit_wire_header *h = static_cast<cubit_wire_header *>(host_wire);
	memset(h, 0, sizeof(*h));
	h->magic = CUBIT_WIRE_MAGIC;
	h->n_streams = n_streams;
	h->n_rows = n;
	h->n_chunks = (n + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK;
	h->data_offset = (sizeof(cubit_wire_header) + (uint64_t)n_streams * h->n_chunks * sizeof(cubit_wire_dir) + 255) &
	                 ~255ull;
	WireArgs a {};
	uint32_t s = 0;
	if (with_rowids) {
		a.src[s] = r->d_ids + offset;
		a.elem[s] = 8;
		h->elem[s++] = 8;
	}
	for (uint32_t c = 0; c < n_cols; c++) {
		const uint32_t w = r->val_elem[c];
		a.src[s] = static_cast<const char *>(r->d_vals[c]) + offset * w;
		a.elem[s] = w;
		h->elem[s++] = (uint8_t)w;
	}
	a.n_rows = n;
	a.n_chunks = h->n_chunks;
	a.dir = reinterpret_cast<cubit_wire_dir *>(dev_wire + sizeof(cubit_wire_header));
	a.slots = dev_wire + h->data_offset;
	cubit_gpu_fetch_ticket *tk = new cubit_gpu_fetch_ticket();
	if (n) {
		r->copies_in_flight.store(1); // free_result drains the copy streams before the buffers go back to the pool
		cudaStream_t cs = t->copy_stream[t->next_copy.fetch_add(1) % kCopyStreams];
		cudaError_t e = cudaStreamWaitEvent(cs, r->ev_done, 0);
		if (e == cudaSuccess) {
			cubit_wire_pack_kernel<<<dim3((unsigned)h->n_chunks, n_streams), kPackThreads, 0, cs>>>(a);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) {
			e = cudaEventCreateWithFlags(&tk->ev, cudaEventDisableTiming);
		}
		if (e == cudaSuccess) {
			e = cudaEventRecord(tk->ev, cs);
		}
		if (e != cudaSuccess) {
			if (tk->ev) {
				cudaEventDestroy(tk->ev);
			}
			delete tk;
			CU_TRY(e);
		}
		t->launches++;
	}
	*ticket = tk;
	return CUBIT_OK;
	ABI_END
}

// ------------------------------------------------------------------------------------------------ drain
namespace {

struct DrainWindow {
	cubit_gpu_result *part; // the (unsharded) result the window belongs to
	uint64_t begin, n;      // rows inside `part`
	uint64_t global_row;    // first row of the window in the whole result
};

struct DrainShared {
	std::vector<DrainWindow> windows;
	std::atomic<uint64_t> next {0};
	std::atomic<int> rc {CUBIT_OK};
	std::string err;
	std::mutex err_mu;
	int with_rowids;
	uint32_t n_cols, n_streams;
	uint32_t elem[CUBIT_MAX_PROBE_COLS];
	bool has_nulls[CUBIT_MAX_PROBE_COLS];
	uint64_t wire_bytes_cap;
	cubit_chunk_fn fn;
	void *ctx;
	cubit_gpu_table *pool_owner;
};

struct DrainLocal {
	uint64_t rows = 0, chunks = 0, windows = 0, wire_bytes = 0, sum_rowids = 0;
	uint64_t sum_cols[CUBIT_MAX_PROBE_COLS] = {};
};

void drain_fail(DrainShared &sh, int rc, const char *why) {
	std::lock_guard<std::mutex> lk(sh.err_mu);
	if (sh.rc.load() == CUBIT_OK) {
		sh.err = why;
		sh.rc.store(rc);
	}
}

void drain_worker(DrainShared &sh, uint32_t worker, DrainLocal &out) {
	struct Slot {
		void *wire = nullptr;
		cubit_gpu_fetch_ticket *ticket = nullptr;
		uint64_t index = UINT64_MAX;
		std::vector<std::vector<uint64_t>> validity;
	} slot[2];
	for (auto &s : slot) {
		s.wire = wire_pool_acquire(sh.pool_owner, sh.wire_bytes_cap);
		if (!s.wire) {
			drain_fail(sh, CUBIT_ENOMEM, "page-locked wire window");
		}
		s.validity.resize(sh.n_cols);
	}
	// worker-local DataChunk: the vectors a GetData call would fill
	std::vector<int64_t> rowids(sh.with_rowids ? CUBIT_WIRE_CHUNK : 0);
	std::vector<std::vector<uint64_t>> colbuf(sh.n_cols, std::vector<uint64_t>(CUBIT_WIRE_CHUNK));
	const void *col_ptr[CUBIT_MAX_PROBE_COLS] = {};
	const uint64_t *val_ptr[CUBIT_MAX_PROBE_COLS] = {};
	for (uint32_t c = 0; c < sh.n_cols; c++) {
		col_ptr[c] = colbuf[c].data();
	}
	auto claim = [&](Slot &s) {
		s.ticket = nullptr;
		s.index = sh.rc.load() == CUBIT_OK ? sh.next.fetch_add(1) : UINT64_MAX;
		if (s.index >= sh.windows.size()) {
			s.index = UINT64_MAX;
			return;
		}
		const DrainWindow &w = sh.windows[s.index];
		int rc = cubit_gpu_fetch_wire_async(w.part, w.begin, w.n, sh.with_rowids, sh.n_cols, s.wire, sh.wire_bytes_cap,
		                                    &s.ticket);
		for (uint32_t c = 0; rc == CUBIT_OK && c < sh.n_cols; c++) {
			s.validity[c].clear();
			if (!sh.has_nulls[c]) {
				continue;
			}
			int all = 1;
			s.validity[c].assign((w.n + 63) / 64, 0);
			rc = cubit_gpu_fetch_validity(w.part, c, w.begin, w.n, s.validity[c].data(), &all);
			if (all) {
				s.validity[c].clear();
			}
		}
		if (rc) {
			drain_fail(sh, rc, last_error_cstr());
			cubit_gpu_fetch_wait(s.ticket);
			s.ticket = nullptr;
			s.index = UINT64_MAX;
		}
	};
	if (slot[0].wire && slot[1].wire) {
		claim(slot[0]); // two windows in flight per worker from the start
		claim(slot[1]);
		int cur = 0;
		while (slot[cur].index != UINT64_MAX) {
			Slot &s = slot[cur];
			const int wrc = cubit_gpu_fetch_wait(s.ticket);
			s.ticket = nullptr;
			if (wrc) {
				drain_fail(sh, wrc, last_error_cstr());
				break;
			}
			const DrainWindow &w = sh.windows[s.index];
			const uint64_t n_chunks = (w.n + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK;
			out.wire_bytes += cubit_wire_payload_bytes(s.wire) + sizeof(cubit_wire_header);
			for (uint64_t ch = 0; ch < n_chunks && sh.rc.load(std::memory_order_relaxed) == CUBIT_OK; ch++) {
				uint32_t stream = 0;
				int n = 0;
				if (sh.with_rowids) {
					n = cubit_wire_unpack_chunk(s.wire, stream++, ch, rowids.data(), 8);
				}
				for (uint32_t c = 0; c < sh.n_cols && n >= 0; c++) {
					n = cubit_wire_unpack_chunk(s.wire, stream++, ch, colbuf[c].data(), sh.elem[c]);
					val_ptr[c] = s.validity[c].empty() ? nullptr : s.validity[c].data() + ch * (CUBIT_WIRE_CHUNK / 64);
				}
				if (n < 0) {
					drain_fail(sh, CUBIT_ECUDA, "malformed wire from the device");
					break;
				}
				if (sh.fn) {
					if (sh.fn(sh.ctx, worker, s.index, w.global_row + ch * CUBIT_WIRE_CHUNK, (uint32_t)n,
					          sh.with_rowids ? rowids.data() : nullptr, col_ptr, val_ptr)) {
						drain_fail(sh, CUBIT_ESTATE, "the consumer stopped the drain");
						break;
					}
				} else { // checksum consumer: reads every delivered value
					if (sh.with_rowids) {
						uint64_t acc = 0;
						for (int i = 0; i < n; i++) {
							acc += (uint64_t)rowids[i];
						}
						out.sum_rowids += acc;
					}
					for (uint32_t c = 0; c < sh.n_cols; c++) {
						uint64_t acc = 0;
						if (sh.elem[c] == 8) {
							const uint64_t *p = colbuf[c].data();
							for (int i = 0; i < n; i++) {
								acc += p[i];
							}
						} else {
							const uint32_t *p = reinterpret_cast<const uint32_t *>(colbuf[c].data());
							for (int i = 0; i < n; i++) {
								acc += p[i];
							}
						}
						out.sum_cols[c] += acc;
					}
				}
				out.rows += (uint64_t)n;
				out.chunks++;
			}
			out.windows++;
			claim(s); // re-arm this window's buffer; the other one has been in flight all along
			cur ^= 1;
		}
	}
	for (auto &s : slot) {
		cubit_gpu_fetch_wait(s.ticket);
		if (s.wire) {
			std::lock_guard<std::mutex> ml(sh.pool_owner->meta_mu);
			sh.pool_owner->wire_pool.emplace_back(s.wire, sh.wire_bytes_cap);
		}
	}
}

} // namespace

extern "C" int cubit_gpu_drain(cubit_gpu_result *r, int with_rowids, uint32_t n_cols, uint32_t n_threads,
                               uint64_t window_rows, cubit_chunk_fn fn, void *ctx, cubit_drain_stats *stats) {
	ABI_BEGIN
	if (!r) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (stats) {
		memset(stats, 0, sizeof(*stats));
	}
	cubit_result_info info;
	int rc = cubit_gpu_result_get(r, &info);
	if (rc) {
		return rc;
	}
	if (n_cols > r->n_cols) {
		return fail(CUBIT_EINVAL, "result has %u projected columns, %u requested", r->n_cols, n_cols);
	}
	if (with_rowids && !(r->flags & CUBIT_Q_ROWIDS)) {
		return fail(CUBIT_ESTATE, "query did not materialise row IDs (CUBIT_Q_ROWIDS)");
	}
	if (!with_rowids && n_cols == 0) {
		return fail(CUBIT_EINVAL, "nothing to drain: no row IDs and no columns requested");
	}
	if (window_rows == 0) {
		window_rows = 128 * CUBIT_WIRE_CHUNK;
	}
	if (window_rows % CUBIT_WIRE_CHUNK) {
		return fail(CUBIT_EINVAL, "window_rows must be a multiple of %u", CUBIT_WIRE_CHUNK);
	}
	DrainShared sh;
	sh.with_rowids = with_rowids ? 1 : 0;
	sh.n_cols = n_cols;
	sh.n_streams = sh.with_rowids + n_cols;
	sh.fn = fn;
	sh.ctx = ctx;
	sh.pool_owner = r->t;
	sh.wire_bytes_cap = cubit_wire_bytes(window_rows, sh.n_streams);
	std::vector<cubit_gpu_result *> parts = r->parts.empty() ? std::vector<cubit_gpu_result *> {r} : r->parts;
	for (uint32_t c = 0; c < n_cols; c++) {
		sh.elem[c] = parts[0]->val_elem[c];
		sh.has_nulls[c] = false;
		for (auto *p : parts) {
			sh.has_nulls[c] = sh.has_nulls[c] || p->d_valid[c] != nullptr;
		}
	}
	uint64_t global_row = 0;
	for (auto *p : parts) { // shard order is row order: windows never straddle a shard
		cubit_result_info pi;
		rc = cubit_gpu_result_get(p, &pi);
		if (rc) {
			return rc;
		}
		for (uint64_t b = 0; b < pi.count; b += window_rows) {
			const uint64_t n = std::min<uint64_t>(window_rows, pi.count - b);
			sh.windows.push_back(DrainWindow {p, b, n, global_row + b});
		}
		global_row += pi.count;
	}
	if (n_threads == 0) {
		n_threads = 1;
	}
	n_threads = (uint32_t)std::min<uint64_t>(n_threads, std::max<uint64_t>(1, (sh.windows.size() + 1) / 2));
	std::vector<DrainLocal> locals(n_threads);
	if (n_threads == 1) {
		drain_worker(sh, 0, locals[0]);
	} else {
		std::vector<std::thread> th;
		for (uint32_t i = 0; i < n_threads; i++) {
			th.emplace_back([&sh, &locals, i]() { drain_worker(sh, i, locals[i]); });
		}
		for (auto &x : th) {
			x.join();
		}
	}
	if (stats) {
		for (auto &l : locals) {
			stats->rows += l.rows;
			stats->chunks += l.chunks;
			stats->windows += l.windows;
			stats->wire_bytes += l.wire_bytes;
			stats->sum_rowids += l.sum_rowids;
			for (uint32_t c = 0; c < n_cols; c++) {
				stats->sum_cols[c] += l.sum_cols[c];
			}
		}
		stats->workers = n_threads;
		for (uint32_t c = 0; c < n_cols; c++) {
			stats->wide_bytes += stats->rows * sh.elem[c];
		}
		stats->wide_bytes += with_rowids ? stats->rows * 8 : 0;
	}
	if (sh.rc.load() != CUBIT_OK) {
		return fail(sh.rc.load(), "drain: %s", sh.err.c_str());
	}
	return CUBIT_OK;
	ABI_END
}
