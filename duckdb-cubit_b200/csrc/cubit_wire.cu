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
//   cubit_wire_stats_kernel     one CTA per frame (chunk, stream): coalesced load of 2048 values, block min / max → form
//   cubit_wire_pack_kernel      one CTA per frame: its offset from the forms before it, deltas / bitmap staged in
//                               shared memory, contiguous 16-byte stores to host memory — frames end up back to back
//                               in consumption order, so the host reads one sequential stream
//   cubit_gpu_fetch_wire_async  enqueue it on a copy stream behind the query
//   cubit_gpu_drain             n worker threads: claim window → two wires in flight → unpack chunk → callback
#include "table.h"
#include "../../include/cubit_gpu_wire.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace cubit;

namespace cubit {

constexpr int kPackThreads = 256;
constexpr uint64_t kWireStatsCap = 1ull << 18; // frames in a result's ring of forms (4 MiB): ≫ the windows in flight
constexpr uint64_t kWireMaxFrames = kWireStatsCap / 8; // frames of ONE wire (32 Ki: e.g. 16 Mi rows of two streams)
constexpr int kPerThread = CUBIT_WIRE_CHUNK / kPackThreads; // 8

struct WireArgs {
	const void *src[CUBIT_MAX_PROBE_COLS + 1]; // first value of the window, per stream
	uint32_t elem[CUBIT_MAX_PROBE_COLS + 1];   // 4 or 8
	uint32_t ascending;                        // bit s: stream s is strictly ascending (the row IDs): bitmap frames allowed
	uint32_t n_streams;
	uint64_t n_rows;
	uint64_t n_chunks;
	uint4 *stats;        // device scratch: per frame (base lo, base hi, width field, n), frame = chunk * n_streams + stream
	cubit_wire_dir *dir; // device-visible address of the wire's directory (host memory)
	unsigned char *frames; // device-visible address of header.data_offset
};

__device__ __forceinline__ long long shfl_xor_ll(long long v, int m) {
	return __shfl_xor_sync(0xffffffffu, v, m);
}

__device__ __forceinline__ void load_chunk(const WireArgs &a, uint32_t chunk, uint32_t stream, uint32_t t, uint32_t n,
                                           long long (&v)[kPerThread]) {
	const uint64_t row0 = (uint64_t)chunk * CUBIT_WIRE_CHUNK;
	if (a.elem[stream] == 8) {
		const long long *s = static_cast<const long long *>(a.src[stream]) + row0;
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			const uint32_t i = t + j * kPackThreads;
			v[j] = i < n ? __ldg(s + i) : 0;
		}
	} else {
		const int *s = static_cast<const int *>(a.src[stream]) + row0;
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			const uint32_t i = t + j * kPackThreads;
			v[j] = i < n ? (long long)__ldg(s + i) : 0;
		}
	}
}

__device__ __forceinline__ uint32_t frame_bytes(uint32_t width, uint32_t n) {
	const uint32_t b = (width & 255u) == CUBIT_WIRE_BITMAP ? (width >> 8) * 8u : n * width;
	return (b + 15u) & ~15u;
}

// pass 1: the form of every frame (one CTA per frame): block min / max → width, or a bitmap for a dense ascending stream
__global__ void __launch_bounds__(kPackThreads) cubit_wire_stats_kernel(const WireArgs a) {
	__shared__ long long s_min[kPackThreads / 32], s_max[kPackThreads / 32];
	const uint32_t chunk = blockIdx.x, stream = blockIdx.y, t = threadIdx.x;
	const uint32_t n = (uint32_t)min((uint64_t)CUBIT_WIRE_CHUNK, a.n_rows - (uint64_t)chunk * CUBIT_WIRE_CHUNK);
	long long v[kPerThread];
	load_chunk(a, chunk, stream, t, n, v);
	long long lo = LLONG_MAX, hi = LLONG_MIN;
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
	if (t == 0) {
#pragma unroll
		for (int w = 0; w < kPackThreads / 32; w++) {
			lo = min(lo, s_min[w]);
			hi = max(hi, s_max[w]);
		}
		const unsigned long long range = (unsigned long long)hi - (unsigned long long)lo; // true difference, < 2^64
		uint32_t width = range == 0 ? 0 : range < 256ull ? 1 : range < 65536ull ? 2 : range < (1ull << 32) ? 4 : 8;
		// a strictly ascending stream of a dense selection is cheaper as a bitmap over [lo, hi] (≥ 2x fewer bytes, else
		// the slower host decode is not worth it)
		const uint32_t words = (uint32_t)min(range / 64ull + 1ull, 0xffffffull);
		if (((a.ascending >> stream) & 1u) && (unsigned long long)words * 16ull <= (unsigned long long)n * width) {
			width = CUBIT_WIRE_BITMAP | (words << 8);
		}
		a.stats[(uint64_t)chunk * a.n_streams + stream] =
		    make_uint4((unsigned int)(unsigned long long)lo, (unsigned int)((unsigned long long)lo >> 32), width, n);
	}
}

// pass 2: one CTA per frame: its offset = the padded sizes of all frames before it (≤ 1152 per window: summed by the
// CTA itself, no scan kernel), deltas / bitmap staged in shared memory, contiguous 16-byte stores into host memory
__global__ void __launch_bounds__(kPackThreads) cubit_wire_pack_kernel(const WireArgs a) {
	__shared__ __align__(16) unsigned char stage[CUBIT_WIRE_SLOT_BYTES];
	__shared__ unsigned long long s_off[kPackThreads / 32];
	const uint32_t chunk = blockIdx.x, stream = blockIdx.y, t = threadIdx.x;
	const uint64_t frame = (uint64_t)chunk * a.n_streams + stream;
	unsigned long long off = 0;
	for (uint64_t f = t; f < frame; f += kPackThreads) {
		const uint4 e = __ldg(a.stats + f);
		off += frame_bytes(e.z, e.w);
	}
#pragma unroll
	for (int m = 16; m; m >>= 1) {
		off += __shfl_xor_sync(0xffffffffu, off, m);
	}
	if ((t & 31) == 0) {
		s_off[t >> 5] = off;
	}
	const uint4 me = __ldg(a.stats + frame);
	const uint32_t width = me.z, n = me.w;
	const unsigned long long lo = (unsigned long long)me.x | ((unsigned long long)me.y << 32);
	long long v[kPerThread];
	load_chunk(a, chunk, stream, t, n, v);
	if ((width & 255u) == CUBIT_WIRE_BITMAP) {
		unsigned int *bits = reinterpret_cast<unsigned int *>(stage);
		for (uint32_t w = t; w < (width >> 8) * 2; w += kPackThreads) {
			bits[w] = 0;
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			if (t + j * kPackThreads < n) {
				const unsigned int d = (unsigned int)((unsigned long long)v[j] - lo);
				atomicOr(bits + (d >> 5), 1u << (d & 31u));
			}
		}
	} else {
#pragma unroll
		for (int j = 0; j < kPerThread; j++) {
			const uint32_t i = t + j * kPackThreads;
			const unsigned long long d = (unsigned long long)v[j] - lo;
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
	}
	__syncthreads();
	off = 0;
#pragma unroll
	for (int w = 0; w < kPackThreads / 32; w++) {
		off += s_off[w];
	}
	// contiguous 16-byte stores: a warp writes 512 consecutive bytes of host memory per instruction (the padding of
	// the last 16-byte unit is shared-memory garbage the unpacker never reads)
	const uint32_t units = frame_bytes(width, n) / 16;
	uint4 *dst = reinterpret_cast<uint4 *>(a.frames + off);
	const uint4 *src16 = reinterpret_cast<const uint4 *>(stage);
	for (uint32_t u = t; u < units; u += kPackThreads) {
		dst[u] = src16[u];
	}
	if (t == 0) {
		unsigned long long *d = reinterpret_cast<unsigned long long *>(a.dir + frame);
		d[0] = lo;
		d[1] = off;
		d[2] = (unsigned long long)width | ((unsigned long long)n << 32);
	}
}

// page-locked wire windows recycled across drains (cudaHostAlloc costs milliseconds)
struct PooledWire {
	void *p = nullptr;
	uint64_t bytes = 0; // what the buffer really holds (it goes back to the pool with this, whatever was asked for)
};

PooledWire wire_pool_acquire(cubit_gpu_table *t, uint64_t bytes) {
	{
		std::lock_guard<std::mutex> ml(t->meta_mu);
		size_t best = SIZE_MAX; // smallest buffer that fits
		for (size_t i = 0; i < t->wire_pool.size(); i++) {
			if (t->wire_pool[i].second >= bytes && (best == SIZE_MAX || t->wire_pool[i].second < t->wire_pool[best].second)) {
				best = i;
			}
		}
		if (best != SIZE_MAX) {
			PooledWire w {t->wire_pool[best].first, t->wire_pool[best].second};
			t->wire_pool[best] = t->wire_pool.back();
			t->wire_pool.pop_back();
			return w;
		}
	}
	PooledWire w;
	if (cudaHostAlloc(&w.p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		w.p = nullptr;
		return w;
	}
	w.bytes = bytes;
	return w;
}

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
	if (!r->parts.empty()) {
		// a sharded result: the shard that holds the whole window writes the wire; a window that straddles two shards
		// has no single device to write it from
		for (size_t i = 0; i < r->parts.size(); i++) {
			if (offset >= r->count_prefix[i] && offset + n <= r->count_prefix[i + 1] && (n > 0 || i + 1 == r->parts.size() || offset < r->count_prefix[i + 1])) {
				return cubit_gpu_fetch_wire_async(r->parts[i], offset - r->count_prefix[i], n, with_rowids, n_cols, host_wire,
				                                  host_wire_bytes, ticket);
			}
		}
		return fail(CUBIT_ESTATE, "rows [%llu, +%llu) straddle two shards of a sharded result: no single device can write "
		                          "the wire — use cubit_gpu_drain, a window inside one shard, or cubit_gpu_fetch_async",
		            (unsigned long long)offset, (unsigned long long)n);
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
	if (((n + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK) * n_streams > kWireMaxFrames) {
		// (every frame's CTA sums the forms of the frames before it: fine for windows, quadratic for whole results)
		return fail(CUBIT_EINVAL, "a wire holds at most %llu frames (chunks x streams): fetch the result in windows",
		            (unsigned long long)kWireMaxFrames);
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
	h->data_offset = cubit_wire_data_offset(h->n_chunks, n_streams);
	WireArgs a {};
	uint32_t s = 0;
	if (with_rowids) {
		// sorted unique row IDs (art.cpp:974-985) may travel as bitmaps.  Opt-in: on a 16-vCPU host the bit-by-bit
		// decode costs the workers more time than the saved bus bytes buy (profiles/r2_narrow_wire.md)
		a.ascending = getenv("CUBIT_WIRE_BITMAP") ? 1u : 0u;
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
	a.n_streams = n_streams;
	a.dir = reinterpret_cast<cubit_wire_dir *>(dev_wire + sizeof(cubit_wire_header));
	a.frames = dev_wire + h->data_offset;
	cubit_gpu_fetch_ticket *tk = new cubit_gpu_fetch_ticket();
	if (n) {
		r->copies_in_flight.store(1); // free_result drains the copy streams before the buffers go back to the pool
		cudaStream_t cs = t->copy_stream[t->next_copy.fetch_add(1) % kCopyStreams];
		cudaError_t e = cudaStreamWaitEvent(cs, r->ev_done, 0);
		// the frames' forms between the two passes live in a per-result ring (no allocation per window: sixteen
		// workers calling cudaMallocAsync / cudaFreeAsync per window serialise in the driver)
		const uint64_t frames = h->n_chunks * n_streams;
		uint64_t slot = 256;
		while (slot < frames) {
			slot <<= 1;
		}
		void *scratch = nullptr;
		if (e == cudaSuccess) {
			if (!r->d_wire_stats.load(std::memory_order_acquire)) {
				std::lock_guard<std::mutex> lk(r->fin_mu);
				if (!r->d_wire_stats.load(std::memory_order_relaxed)) {
					void *p = nullptr;
					e = cudaMallocAsync(&p, kWireStatsCap * sizeof(uint4), cs);
					if (e == cudaSuccess) {
						e = cudaStreamSynchronize(cs); // the other copy streams use it too: once per result
					}
					if (e == cudaSuccess) {
						r->d_wire_stats.store(static_cast<uint4 *>(p), std::memory_order_release);
					}
				}
			}
			if (e == cudaSuccess) {
				uint64_t pos;
				do { // a slot never straddles the end of the ring; the ring is far larger than what can be in flight
					pos = r->wire_stats_cursor.fetch_add(slot) % kWireStatsCap;
				} while (pos + slot > kWireStatsCap);
				scratch = r->d_wire_stats.load(std::memory_order_acquire) + pos;
			}
		}
		if (e == cudaSuccess) {
			a.stats = static_cast<uint4 *>(scratch);
			const dim3 grid((unsigned)h->n_chunks, n_streams);
			cubit_wire_stats_kernel<<<grid, kPackThreads, 0, cs>>>(a);
			cubit_wire_pack_kernel<<<grid, kPackThreads, 0, cs>>>(a);
			e = cudaGetLastError();
			t->launches += 2;
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

// wrapping sum with four independent chains (a single chain is one add per cycle: 0.46 ns per value)
inline uint64_t sum_u64(const uint64_t *p, int n) {
	uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
	int i = 0;
	for (; i + 4 <= n; i += 4) {
		a0 += p[i];
		a1 += p[i + 1];
		a2 += p[i + 2];
		a3 += p[i + 3];
	}
	for (; i < n; i++) {
		a0 += p[i];
	}
	return a0 + a1 + a2 + a3;
}

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
		uint64_t wire_cap = 0;
		cubit_gpu_fetch_ticket *ticket = nullptr;
		uint64_t index = UINT64_MAX;
		std::vector<std::vector<uint64_t>> validity;
	} slot[2];
	for (auto &s : slot) {
		const PooledWire pw = wire_pool_acquire(sh.pool_owner, sh.wire_bytes_cap);
		s.wire = pw.p;
		s.wire_cap = pw.bytes;
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
		int rc = cubit_gpu_fetch_wire_async(w.part, w.begin, w.n, sh.with_rowids, sh.n_cols, s.wire, s.wire_cap, &s.ticket);
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
						out.sum_rowids += sum_u64(reinterpret_cast<const uint64_t *>(rowids.data()), n);
					}
					for (uint32_t c = 0; c < sh.n_cols; c++) {
						uint64_t acc = 0;
						if (sh.elem[c] == 8) {
							acc = sum_u64(colbuf[c].data(), n);
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
			sh.pool_owner->wire_pool.emplace_back(s.wire, s.wire_cap);
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
	if (n_threads == 0) {
		n_threads = 1;
	}
	if (window_rows == 0) { // ≈ 4 windows per worker, between 16 and 256 DataChunks each
		const uint64_t per = info.count / (4ull * n_threads) / CUBIT_WIRE_CHUNK;
		window_rows = std::min<uint64_t>(256, std::max<uint64_t>(16, per)) * CUBIT_WIRE_CHUNK;
	}
	if (window_rows % CUBIT_WIRE_CHUNK) {
		return fail(CUBIT_EINVAL, "window_rows must be a multiple of %u", CUBIT_WIRE_CHUNK);
	}
	{ // one wire holds at most kWireMaxFrames frames: larger windows are cut down
		const uint64_t max_chunks = kWireMaxFrames / ((with_rowids ? 1u : 0u) + n_cols);
		window_rows = std::min<uint64_t>(window_rows, max_chunks * CUBIT_WIRE_CHUNK);
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
	n_threads = (uint32_t)std::min<uint64_t>(n_threads, std::max<uint64_t>(1, (sh.windows.size() + 1) / 2));
	std::vector<DrainLocal> locals(n_threads);
	if (n_threads == 1) {
		drain_worker(sh, 0, locals[0]);
	} else {
		std::vector<std::thread> th;
		th.reserve(n_threads);
		try {
			for (uint32_t i = 0; i < n_threads; i++) {
				th.emplace_back([&sh, &locals, i]() { drain_worker(sh, i, locals[i]); });
			}
		} catch (...) { // the system is out of threads: the workers that did start drain everything
			if (th.empty()) {
				drain_worker(sh, 0, locals[0]);
			}
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
