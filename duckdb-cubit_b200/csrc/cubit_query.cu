#include "table.h"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>

using namespace cubit;

// cubit_query.cu — query planning, kernel launches and the result hand-off of the C-ABI.
//
// cubit_gpu_query flattens the AND-of-ORs predicate into the ordered stream list, bounds the result size from
// the per-bitvector cardinalities and enqueues the kernels of scan_kernel.cu / aux_kernels.cu on the shard's
// kernel stream.  The table mutex is held only while the work is planned and ENQUEUED; waiting for the GPU
// (finish_result) and moving rows to the host (cubit_gpu_fetch*, on the copy streams) happen outside it, so
// several host threads keep queries and hand-offs in flight on one table (table.h).
// -------------------------------------------------------------------- query

static cudaError_t run_scan(const ScanArgs &sa, uint32_t tile_words, bool has_delta, bool compressed, int sm_count,
                            cudaStream_t st) {
	return launch_scan(sa, tile_words, has_delta, compressed, sm_count, st, nullptr);
}

// (the pinned header and the completion event go back to the table's pools)
static void release_result_locked(cubit_gpu_result *r) {
	cudaStream_t s = r->stream;
	if (r->copies_in_flight.load()) { // rows still crossing PCIe: the buffers must outlive the copies
		for (auto cs : r->t->copy_stream) {
			cudaStreamSynchronize(cs);
		}
	}
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
	if (uint4 *ws = r->d_wire_stats.load()) {
		cudaFreeAsync(ws, s);
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
	std::lock_guard<std::mutex> ml(r->t->meta_mu);
	if (r->h_hdr) {
		r->t->hdr_pool.push_back(r->h_hdr);
	}
	for (auto &e : r->ev) {
		if (e) {
			cudaEventDestroy(e);
		}
	}
	if (r->ev_done) {
		r->t->ev_pool.push_back(r->ev_done); // (completed or never recorded: safe to record again)
	}
	delete r;
}

// Waits for the query (outside the table mutex) and finalises info.  Several consumer threads of one result may
// race here (parallel DataChunk hand-off): the first one does the work, the others get its outcome.
static int finish_result(cubit_gpu_result *r) {
	if (!r->parts.empty()) {
		return sharded_result_finish(r);
	}
	if (r->finished.load(std::memory_order_acquire)) {
		return r->fin_rc ? fail(r->fin_rc, "%s", r->fin_err.c_str()) : CUBIT_OK;
	}
	std::lock_guard<std::mutex> lk(r->fin_mu);
	if (r->finished.load(std::memory_order_acquire)) {
		return r->fin_rc ? fail(r->fin_rc, "%s", r->fin_err.c_str()) : CUBIT_OK;
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
	int rc = CUBIT_OK;
	if (r->h_hdr->overflow) {
		rc = fail(CUBIT_EINVAL, "Overflow in multiplication of INT64 in SUM(a*b)");
	} else if ((r->flags & (CUBIT_Q_ROWIDS | CUBIT_Q_VALUES)) && r->info.count > r->info.capacity) {
		rc = fail(CUBIT_ESTATE, "internal: result %llu exceeds capacity bound %llu", (unsigned long long)r->info.count,
		          (unsigned long long)r->info.capacity);
	}
	r->fin_rc = rc;
	if (rc) {
		r->fin_err = last_error_cstr();
	}
	r->finished.store(true, std::memory_order_release);
	return rc;
}

// Raw or FOR-bit-packed form of a column for THIS query, when both are resident (cubit_gpu_pack_column keep_raw).
// The gather is bound either by the DRAM lines it touches or by the per-value decode, so the choice follows the
// selection density s and the packed width w (constants measured on B200: profiles/r1_experiment_packed_payload.log,
// profiles/r2_probe_forms.md):
//   raw    : 128-byte lines of 16 values → touches 8·N·(1 − (1−s)^16) bytes; issue floor ≈ 360 G values/s
//   packed : lines of 1024/w values      → touches (w/8)·N·(1 − (1−s)^(1024/w)) bytes; decode ≈ 170 G values/s
static bool prefer_raw_form(const Column *c, uint64_t n_rows, uint64_t cap_rows, bool gather_over_ids) {
	if (!c->packed() || !c->d) {
		return false; // only one form is resident
	}
	if (gather_over_ids) {
		return true; // sparse gather over the row-ID list is latency-bound: one dependent load beats header + payload
	}
	const double n = (double)n_rows, s = n_rows ? std::min(1.0, (double)cap_rows / n) : 0.0;
	const double w = std::max(1.0, c->pack_avg_width);
	const double kBw = 6.2e12, kRawRate = 360e9, kPackRate = 170e9;
	const double raw_bytes = 8.0 * n * (1.0 - pow(1.0 - s, 16.0));
	const double pk_bytes = (w / 8.0) * n * (1.0 - pow(1.0 - s, 1024.0 / w));
	const double t_raw = std::max(raw_bytes / kBw, s * n / kRawRate);
	const double t_pk = std::max(pk_bytes / kBw, s * n / kPackRate);
	return t_raw <= t_pk;
}

// plans the query and enqueues its kernels; caller holds t->mu
static int plan_and_launch(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out) {
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}

	// ---- flatten predicate into the ordered stream list
	ScanArgs sa;
	memset(&sa, 0, sizeof(sa));
	uint32_t k = 0;
	bool has_delta = false, has_compressed = false;
	uint64_t delta_entries = 0, stream_bytes = 0;
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
			int rc = CUBIT_OK;
			if (ix->delta.voff_pending || !ix->counts_valid) { // (rare: right after maintenance)
				std::lock_guard<std::mutex> ml(t->meta_mu);
				rc = delta_settle_locked(t, ix);
				if (rc == CUBIT_OK) {
					rc = refresh_counts(t, ix);
				}
			}
			if (rc) {
				return rc;
			}
			const uint32_t v = grp.refs[i].value_id;
			if (ix->compressed) {
				// containers: the producer reads the directory entry of (value, segment) and copies the container
				// (or nothing, for EMPTY / FULL) — bytes actually read are reported by the index, not k·N/8
				sa.bv[k] = reinterpret_cast<const uint64_t *>(ix->cs.d_pool);
				sa.cdir[k] = ix->cs.d_dir + (uint64_t)v * ix->cs.n_seg_cap;
				has_compressed = true;
				stream_bytes += (uint64_t)t->n_seg * 8; // directory entries; container payload is data dependent
			} else {
				sa.bv[k] = bv_ptr(t, ix, v);
				stream_bytes += t->n_words * 8;
			}
			DeltaSet &d = ix->delta;
			if (d.d_off && d.rows[v]) {
				if (d.n_seg != t->n_seg) {
					rc = delta_restride_locked(t, ix, t->n_seg);
					if (rc) {
						return rc;
					}
				}
				sa.doff[k] = d.d_off + (uint64_t)v * t->n_seg;
				sa.dent[k] = d.d_ent;
				has_delta = true;
				delta_entries += d.rows[v];
			}
			group_bound += ix->counts[v] + d.rows[v];
			k++;
		}
		sa.group_end |= 1ull << (k - 1);
		cap = std::min(cap, group_bound);
	}
	sa.k = k;
	if (getenv("CUBIT_FORCE_DELTA_KERNEL")) { // timing experiment: the delta-aware kernel instance on clean bitvectors
		has_delta = true;
	}
	if (const char *dbg = getenv("CUBIT_SCAN_DEBUG")) {
		sa.debug = (unsigned)atoi(dbg); // kernel timing experiments: skips parts of the kernel, results invalid
	}
	// Tile = the unit the scan kernel hands out.  It is the table's segment, except for SHORT queries: every tile costs
	// a CTA's consumers ≈ 1 µs of dependent latency (ring wait → fold → count → publish → look-back hand-off → emit),
	// so a tile must carry ≳ 24 KiB for that to hide behind the bulk copies (profiles/r1_ncu_scan_s01.md,
	// profiles/r2_small_k.md).  With few bitvectors (or 32768-row segments) the kernel therefore treats 2 or 4
	// CONSECUTIVE segments as one tile — bitvectors are contiguous, so this is the same kernel instantiated for the
	// larger segment size.  Pending deltas and containers are keyed by the table's own segments and keep them.
	uint32_t tile_mult = 1;
	if (!has_delta && !has_compressed && !(q->flags & CUBIT_Q_UNFUSED)) {
		while (t->seg_words * tile_mult * 2 <= 2048 && (uint64_t)k * t->seg_words * tile_mult * 8 < 24 * 1024 &&
		       (uint64_t)((t->n_seg + tile_mult * 2 - 1) / (tile_mult * 2)) * (t->seg_words * tile_mult * 2) <= t->words_per_bv) {
			tile_mult *= 2;
		}
		if (t->seg_words == 512 && tile_mult == 1 &&
		    (uint64_t)((t->n_seg + 1) / 2) * 1024 <= t->words_per_bv) {
			tile_mult = 2; // 32768-row segments: never below 65536-row tiles (consumer-bound otherwise)
		}
		if (const char *e = getenv("CUBIT_TILE_MULT")) { // experiments: 1 disables
			const uint32_t m = (uint32_t)atoi(e);
			if ((m == 1 || m == 2 || m == 4) && t->seg_words * m <= 2048 &&
			    (uint64_t)((t->n_seg + m - 1) / m) * (t->seg_words * m) <= t->words_per_bv) {
				tile_mult = m;
			}
		}
	}
	const uint32_t tile_words = t->seg_words * tile_mult;
	const uint32_t n_tile = (t->n_seg + tile_mult - 1) / tile_mult;
	sa.n_seg = n_tile;
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
	const uint64_t sel_bound = cap; // upper bound of the selection (exact for disjoint ORs)
	enum { PROBE_NONE, PROBE_FUSED, PROBE_BITS, PROBE_GATHER } probe_mode = PROBE_NONE;
	if (need_probe) {
		if (!fusable) {
			probe_mode = PROBE_GATHER;
		} else {
			if ((q->flags & CUBIT_Q_FUSE_PROBE) && !has_compressed) {
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
	const bool probe_on_bv = probe_mode == PROBE_BITS && k == 1 && !has_delta && !has_compressed && !want_ids && !want_vals && !want_q &&
	                         !unfused && sa.debug == 0;
	// Dense selections over bit-packed columns (widths <= 32) are STREAMED by the dense probe (probe_dense_kernel.cu):
	// from a few percent of the rows upward every DRAM line of the packed form is touched anyway, and decoding from
	// shared-memory stages costs a third of the per-value gather (profiles/r2_probe_dense.md).  Below the density
	// threshold the bit-driven gather probe keeps the job (and picks raw or packed per column, prefer_raw_form).
	bool dense_probe = false;
	DenseProbeArgs dp;
	memset(&dp, 0, sizeof(dp));
	if (probe_mode == PROBE_BITS && !probe_on_bv) {
		uint64_t inv = 28; // selected rows >= n_rows / inv: below ~3.5 % the per-block bookkeeping of the dense probe
		                   // (≈ 300 instructions per touched block, a floor of 0.35 ms per 10^9 rows) loses to the gather
		if (const char *e = getenv("CUBIT_DENSE_MIN_INV")) {
			inv = strtoull(e, nullptr, 10); // 0 disables the dense probe (experiments)
		}
		if (inv && sel_bound >= t->n_rows / inv) {
			uint32_t mw[kMaxFusedCols] = {};
			dp.n_load = n_dist;
			for (int d = 0; d < n_dist; d++) {
				const bool pk_ok = dist_cols[d]->packed() && dist_cols[d]->pack_max_width <= 32;
				dp.lcol[d] = col_ref(dist_cols[d], !pk_ok);
				mw[d] = dist_cols[d]->pack_max_width;
			}
			dense_probe = dense_probe_plan(dp, mw);
		}
	}
	const bool separate_probe = probe_mode == PROBE_GATHER;
	const bool need_ids_buf = want_ids || separate_probe || (probe_mode == PROBE_BITS && want_vals);
	if (!need_ids_buf && !want_vals) {
		cap = 0;
	}
	cap = (cap + 1) & ~1ull; // even: the probe kernel moves row IDs in pairs

	// SHORT queries on large tables run merge + decode as two streaming passes (small_scan_kernels.cu): the ring
	// kernel's per-tile count → publish → look-back → emit chain costs a CTA 1–2 µs per tile whatever the tile holds,
	// which with one to four bitvectors per tile IS the run time, while re-reading the ONE merged bitvector is cheap.
	// Units of 8192 rows; the probes then see 65536-row tiles (8 units) whatever the table's segment size.
	uint64_t two_pass_min_rows = 16ull << 20;
	if (const char *e = getenv("CUBIT_TWO_PASS_MIN_ROWS")) { // tests: 0 = every eligible query; huge = never
		two_pass_min_rows = strtoull(e, nullptr, 10);
	}
	uint32_t lb_max_k = 3; // with row positions (measured crossover against the ring kernel, profiles/r2_small_k.md)
	if (const char *e = getenv("CUBIT_LB_MAX_K")) { // experiment knob
		lb_max_k = (uint32_t)std::min<unsigned long>(8ul, strtoul(e, nullptr, 10));
	}
	const uint32_t n_units = (uint32_t)(((uint64_t)t->n_seg * t->seg_words + 1023) / 1024) * 8u;
	// measured crossover against the ring kernel at 10^9 rows (profiles/r2_small_k.md): k <= 3, with row positions
	// (the one-pass look-back kernel) as well as count / bitvector / aggregate only
	const bool two_pass = !has_delta && !has_compressed && !unfused && probe_mode != PROBE_FUSED && !probe_on_bv &&
	                      k <= (need_ids_buf ? lb_max_k : 3u) &&
	                      t->n_rows >= two_pass_min_rows && (uint64_t)n_units * 128 <= t->words_per_bv && sa.debug == 0;
	const uint32_t probe_tile_words = two_pass ? 1024u : tile_words;
	const uint32_t probe_n_tile = two_pass ? n_units / 8u : n_tile;

	// ---- result object
	cubit_gpu_result *r = new (std::nothrow) cubit_gpu_result();
	if (!r) {
		return fail(CUBIT_ENOMEM, "host allocation failed");
	}
	r->t = t;
	// Which stream: queries that produce row positions use the look-back between CTAs (a co-resident, persistent
	// grid) and stay on the in-order kernel stream; aggregate-only / bitvector-only queries have no inter-CTA
	// dependency and go to the stream pool, where the small grids of concurrent callers overlap on the GPU.
	int agg_slot = -1;
	cudaStream_t st = t->stream;
	if (t->stream == t->own_stream && !need_ids_buf && !want_vals && !unfused && probe_mode != PROBE_GATHER) {
		agg_slot = (int)(t->next_agg++ % kAggStreams);
		st = t->agg_stream[agg_slot];
	}
	r->stream = st;
	r->flags = q->flags;
	r->agg_kind = q->agg_kind;
	r->n_cols = want_vals ? q->n_cols : 0;
	r->timing = (q->flags & CUBIT_Q_TIMING) != 0;
	int rc = CUBIT_OK;
#define Q_TRY(expr)                                                                                                    \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			rc = fail(_e == cudaErrorMemoryAllocation ? CUBIT_ENOMEM : CUBIT_ECUDA, "%s: %s (%s:%d)", #expr,           \
			          cudaGetErrorString(_e), __FILE__, __LINE__);                                                     \
			release_result_locked(r);                                                                                  \
			return rc;                                                                                                 \
		}                                                                                                              \
	} while (0)

	const int max_grid = std::max(scan_max_grid(t->seg_words, t->sm_count), probe_grid(t->sm_count));
	const size_t hdr_bytes = 64;
	const size_t ctrl_bytes = ((size_t)std::max<uint32_t>(t->n_seg, probe_n_tile) + 1) * 8;
	const size_t ctrl_pad = (ctrl_bytes + 63) & ~(size_t)63;
	const size_t part_bytes = (size_t)max_grid * sizeof(BlockPartial);
	// layout: hdr | ctrl A | ctrl B (decode pass of the unfused path) | partials | probe done | segment prefixes
	const size_t excl_bytes = probe_mode == PROBE_BITS ? ctrl_pad : 0;
	const size_t span_bytes = dense_probe ? (size_t)std::max<uint32_t>(t->n_seg * kConsumerWarps, n_units) * 8 : 0; // per-span prefixes (dense probe)
	SmallScanArgs ss;
	memset(&ss, 0, sizeof(ss));
	if (two_pass) {
		small_scan_plan(ss, n_units, t->sm_count);
	}
	const size_t chunk_bytes = two_pass ? ((size_t)ss.n_chunks + 1) * 8 : 0;
	const size_t block_bytes = hdr_bytes + 2 * ctrl_pad + part_bytes + 64 + excl_bytes + span_bytes + chunk_bytes;
	Q_TRY(cudaMallocAsync((void **)&r->d_block, block_bytes, st));
	r->d_hdr = reinterpret_cast<ResultHeader *>(r->d_block);
	unsigned long long *ctrl_a = reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes);
	unsigned long long *ctrl_b = reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes + ctrl_pad);
	BlockPartial *partials = reinterpret_cast<BlockPartial *>(r->d_block + hdr_bytes + 2 * ctrl_pad);
	unsigned int *probe_done = reinterpret_cast<unsigned int *>(r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes);
	unsigned long long *tile_excl =
	    reinterpret_cast<unsigned long long *>(r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes + 64);
	unsigned long long *span_excl = reinterpret_cast<unsigned long long *>(
	    r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes + 64 + excl_bytes);
	unsigned long long *chunk_tot = reinterpret_cast<unsigned long long *>(
	    r->d_block + hdr_bytes + 2 * ctrl_pad + part_bytes + 64 + excl_bytes + span_bytes);
	{
		std::lock_guard<std::mutex> ml(t->meta_mu);
		if (!t->hdr_pool.empty()) {
			r->h_hdr = t->hdr_pool.back();
			t->hdr_pool.pop_back();
		}
		if (!t->ev_pool.empty()) {
			r->ev_done = t->ev_pool.back();
			t->ev_pool.pop_back();
		}
	}
	if (!r->h_hdr) {
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
	// (two passes over ONE bitvector: Q is that bitvector itself, nothing is written)
	if ((((unfused || probe_mode == PROBE_BITS) && !probe_on_bv) || (two_pass && need_ids_buf)) && !want_q &&
	    !(two_pass && k == 1)) {
		Q_TRY(cudaMallocAsync((void **)&r->d_q_tmp, t->words_per_bv * 8, st));
	}
	if (!r->ev_done) {
		Q_TRY(cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming));
	}
	if (agg_slot >= 0 && t->mut_recorded) {
		Q_TRY(cudaStreamWaitEvent(st, t->mut_event, 0)); // behind the last maintenance section of the kernel stream
	}
	if (r->timing) {
		for (auto &e : r->ev) {
			Q_TRY(cudaEventCreate(&e));
		}
	}

	// zero hdr + both control blocks (ticket counters and per-segment status words)
	Q_TRY(cudaMemsetAsync(r->d_block, 0, hdr_bytes + 2 * ctrl_pad, st));

	sa.partials = partials;
	sa.hdr = r->d_hdr;
	uint32_t n_launch = 0;
	bool lookback = false;
	if (r->timing) {
		Q_TRY(cudaEventRecord(r->ev[0], st));
	}
	if (probe_on_bv) {
		sa.q_out = const_cast<uint64_t *>(sa.bv[0]);
		r->info.fused = 1;
	} else if (two_pass) {
		for (uint32_t i = 0; i < k; i++) {
			ss.bv[i] = sa.bv[i];
		}
		ss.group_end = sa.group_end;
		ss.k = k;
		ss.chunk_tot = chunk_tot;
		ss.hdr = r->d_hdr;
		uint64_t *qbuf = want_q ? r->d_q : r->d_q_tmp; // (nullptr when nobody reads a merged bitvector)
		ss.q_out = (k > 1 || want_q) ? qbuf : nullptr;
		ss.q_in = ss.q_out ? ss.q_out : sa.bv[0];
		ss.count_here = need_ids_buf ? 0 : 1;
		// with row positions: ONE pass with a decoupled look-back (lookback_scan_kernel.cu) instead of merge + count,
		// chunk prefix and decode (CUBIT_NO_LOOKBACK=1 keeps the three launches: A/B runs in profiles/r2_small_k.md)
		lookback = need_ids_buf && getenv("CUBIT_NO_LOOKBACK") == nullptr;
		if (lookback) {
			ss.span_excl = dense_probe && want_vals && cap ? span_excl : nullptr;
			ss.tile_excl = probe_mode == PROBE_BITS && want_vals && !dense_probe ? tile_excl : nullptr;
			ScanArgs ea;
			memset(&ea, 0, sizeof(ea));
			ea.ids_out = (want_ids || separate_probe) && cap ? r->d_ids : nullptr; // positions alone need no row IDs
			ea.row_base = t->row_base;
			ea.ctrl = ctrl_a;      // ticket counter + the total of every 32-unit tile (zeroed above)
			ss.chunk_tot = ctrl_b; // the tiles' inclusive prefixes (zeroed above)
			Q_TRY(launch_lookback_scan(ss, ea, t->sm_count, st));
			n_launch++;
		} else {
			Q_TRY(launch_small_merge_count(ss, st));
			n_launch++;
			if (need_ids_buf) {
				ss.span_excl = dense_probe && want_vals && cap ? span_excl : nullptr;
				ss.tile_excl = probe_mode == PROBE_BITS && want_vals && !dense_probe ? tile_excl : nullptr;
				Q_TRY(launch_small_prefix(ss, st));
				ScanArgs ea;
				memset(&ea, 0, sizeof(ea));
				ea.ids_out = (want_ids || separate_probe) && cap ? r->d_ids : nullptr; // positions alone need no row IDs
				ea.row_base = t->row_base;
				Q_TRY(launch_small_decode(ss, ea, st));
				n_launch += 2;
			}
		}
		sa.q_out = const_cast<uint64_t *>(ss.q_in); // what the probe kernels re-decode
		sa.tile_excl = ss.tile_excl;
		sa.span_excl = ss.span_excl;
		r->info.fused = 0;
	} else if (!unfused) {
		// one pass: merge (+delta XOR) + decode (+ fused probe / aggregate when eligible)
		sa.ctrl = ctrl_a;
		sa.q_out = probe_mode == PROBE_BITS && !want_q ? r->d_q_tmp : r->d_q;
		sa.ids_out = need_ids_buf ? r->d_ids : nullptr;
		sa.tile_excl = probe_mode == PROBE_BITS && want_vals && !dense_probe ? tile_excl : nullptr;
		sa.span_excl = dense_probe && want_vals && cap ? span_excl : nullptr;
		if (probe_mode == PROBE_FUSED) {
			sa.n_load = n_dist;
			for (int d = 0; d < n_dist; d++) {
				sa.lcol[d] = col_ref(dist_cols[d], prefer_raw_form(dist_cols[d], t->n_rows, sel_bound, false));
				sa.lout[d] = dist_out[d] >= 0 && cap ? static_cast<long long *>(r->d_vals[dist_out[d]]) : nullptr;
			}
			sa.agg_kind = q->agg_kind;
			sa.agg_ia = agg_ia;
			sa.agg_ib = agg_ib;
		}
		Q_TRY(run_scan(sa, tile_words, has_delta, has_compressed, t->sm_count, st));
		n_launch++;
		r->info.fused = 1;
	} else {
		// three separate kernels: K1 merge → Q, K2 decode Q → row IDs, K3 probe
		uint64_t *qbuf = want_q ? r->d_q : r->d_q_tmp;
		sa.ctrl = ctrl_a;
		sa.q_out = qbuf;
		sa.ids_out = nullptr;
		Q_TRY(run_scan(sa, t->seg_words, has_delta, has_compressed, t->sm_count, st));
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
			sd.partials = partials;
			sd.hdr = r->d_hdr;
			sd.skip_count = 1; // K1 already counted the selection
			Q_TRY(run_scan(sd, t->seg_words, false, false, t->sm_count, st));
			n_launch++;
		}
		r->info.fused = 0;
	}
	if (r->timing) {
		Q_TRY(cudaEventRecord(r->ev[1], st));
	}
	if (dense_probe) {
		dp.q = sa.q_out;
		dp.span_excl = sa.span_excl;
		dp.n_span = probe_n_tile * (uint32_t)kConsumerWarps;
		dp.n_blk = (t->n_rows + kPackBlock - 1) / kPackBlock;
		for (int d = 0; d < n_dist; d++) {
			dp.lout[d] = dist_out[d] >= 0 && cap ? static_cast<long long *>(r->d_vals[dist_out[d]]) : nullptr;
		}
		dp.agg_kind = q->agg_kind;
		dp.agg_ia = agg_ia;
		dp.agg_ib = agg_ib;
		dp.hdr = r->d_hdr;
		Q_TRY(launch_probe_dense(dp, probe_tile_words, want_vals && cap, t->sm_count, st));
		n_launch++;
		if (r->timing) {
			Q_TRY(cudaEventRecord(r->ev[2], st));
			r->probe_timed = true;
		}
	} else if (probe_mode == PROBE_BITS) {
		ScanArgs pb;
		memset(&pb, 0, sizeof(pb));
		pb.q_out = sa.q_out; // input of the bit-driven probe
		pb.tile_excl = sa.tile_excl;
		pb.n_seg = probe_on_bv ? t->n_seg : probe_n_tile;
		pb.row_base = t->row_base;
		pb.n_load = n_dist;
		for (int d = 0; d < n_dist; d++) {
			pb.lcol[d] = col_ref(dist_cols[d], prefer_raw_form(dist_cols[d], t->n_rows, sel_bound, false));
			pb.lout[d] = dist_out[d] >= 0 && cap ? static_cast<long long *>(r->d_vals[dist_out[d]]) : nullptr;
		}
		pb.agg_kind = q->agg_kind;
		pb.agg_ia = agg_ia;
		pb.agg_ib = agg_ib;
		pb.hdr = r->d_hdr;
		pb.count_rows = probe_on_bv ? 1 : 0;
		Q_TRY(launch_probe_bits(pb, probe_on_bv ? t->seg_words : probe_tile_words, want_vals && cap, t->sm_count, st));
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
	if (agg_slot >= 0) {
		std::lock_guard<std::mutex> ml(t->meta_mu); // (two callers may share a pool stream: record + flag together)
		Q_TRY(cudaEventRecord(t->agg_last[agg_slot], st));
		t->agg_used[agg_slot] = true;
	}
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
		// The merged bitvector the bit-driven / dense probes re-read is intermediate traffic of an unfused design and is
		// NOT counted; when the probe runs straight on a value bitvector (probe_on_bv) that bitvector is the
		// query's one input stream and is counted here, since no scan kernel reads it.
		r->probe_fixed_bytes = probe_on_bv ? t->n_words * 8 : 0;
	}
	r->info.capacity = cap;
	r->info.probe_path = dense_probe                    ? CUBIT_PROBE_DENSE
	                     : probe_mode == PROBE_BITS     ? CUBIT_PROBE_BITS
	                     : probe_mode == PROBE_GATHER   ? CUBIT_PROBE_GATHER
	                     : probe_mode == PROBE_FUSED    ? CUBIT_PROBE_FUSED
	                                                    : CUBIT_PROBE_NONE;
	r->info.scan_path = probe_on_bv ? CUBIT_SCAN_NONE
	                                : (lookback ? CUBIT_SCAN_LOOKBACK : (two_pass ? CUBIT_SCAN_TWO_PASS : CUBIT_SCAN_RING));
	r->info.n_streams = k;
	r->info.n_launches = n_launch;
	r->info.delta_entries = delta_entries;
	// (probe_on_bv: the one bitvector is read once, by the probe — accounted in probe_fixed_bytes)
	r->info.algo_bytes_scan = probe_on_bv ? 0 : stream_bytes + delta_entries * sizeof(DeltaEnt);
	r->info.d_rowids = want_ids ? reinterpret_cast<const int64_t *>(r->d_ids) : nullptr;
	r->info.d_bitvector = r->d_q;
	for (uint32_t c = 0; c < r->n_cols; c++) {
		r->info.d_values[c] = r->d_vals[c];
		r->info.d_validity[c] = r->d_valid[c];
	}
	*out = r;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out) {
	ABI_BEGIN
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
	if (t->sharded()) {
		return sharded_query(t, q, out);
	}
	cubit_gpu_result *r = nullptr;
	int rc;
	{
		std::shared_lock<std::shared_mutex> lk(t->mu);
		rc = plan_and_launch(t, q, &r);
	}
	if (rc) {
		return rc;
	}
	if (!(q->flags & CUBIT_Q_ASYNC)) { // the wait happens OUTSIDE the table mutex: other threads plan and enqueue meanwhile
		rc = finish_result(r);
		if (rc) {
			const std::string why = last_error_cstr();
			cubit_gpu_free_result(r);
			return fail(rc, "%s", why.c_str());
		}
	}
	*out = r;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_result_wait(cubit_gpu_result *r) {
	ABI_BEGIN
	if (!r) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	return finish_result(r);
	ABI_END
}

extern "C" int cubit_gpu_result_get(cubit_gpu_result *r, cubit_result_info *info) {
	ABI_BEGIN
	if (!r || !info) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	*info = r->info;
	return CUBIT_OK;
	ABI_END
}

// Split (count, 128-bit sum) of a finished-or-not result into five non-negative 32-bit limbs stored as int64 and
// ADD them to dst[0..5) on the device, in stream order behind the query — so a multi-process driver can hand
// `dst` straight to one ncclAllReduce(sum) without the aggregates ever visiting the host ("aggregates
// → ncclAllReduce(sum) on int64 limbs").  Limbs: count (< 2^62 as one limb), sum bits [0,32), [32,64), [64,96),
// [96,128) — the top limb is the signed one.
extern "C" int cubit_gpu_result_add_limbs(cubit_gpu_result *r, int64_t *device_dst) {
	ABI_BEGIN
	if (!r || !device_dst) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (!r->parts.empty()) {
		return fail(CUBIT_ESTATE, "a sharded result is already reduced by the library");
	}
	cubit_gpu_table *t = r->t;
	std::shared_lock<std::shared_mutex> lk(t->mu);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	CU_TRY(launch_add_limbs(r->d_hdr, reinterpret_cast<long long *>(device_dst), r->stream));
	t->launches++;
	return CUBIT_OK;
	ABI_END
}

// ---- DataChunk hand-off: result rows device → host on the shard's copy streams
static int fetch_check(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols) {
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
	return CUBIT_OK;
}

static int fetch_enqueue(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                         void *const *host_cols, cudaEvent_t *ev_out) {
	cubit_gpu_table *t = r->t;
	CU_TRY(cudaSetDevice(t->device));
	cudaStream_t cs = t->copy_stream[t->next_copy.fetch_add(1) % kCopyStreams];
	CU_TRY(cudaStreamWaitEvent(cs, r->ev_done, 0));
	if (host_rowids) {
		CU_TRY(cudaMemcpyAsync(host_rowids, r->d_ids + offset, n * 8, cudaMemcpyDeviceToHost, cs));
	}
	for (uint32_t c = 0; c < n_cols; c++) {
		if (host_cols && host_cols[c]) {
			const size_t w = r->val_elem[c];
			CU_TRY(cudaMemcpyAsync(host_cols[c], static_cast<const char *>(r->d_vals[c]) + offset * w, n * w,
			                       cudaMemcpyDeviceToHost, cs));
		}
	}
	cudaEvent_t ev = nullptr;
	CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	cudaError_t e = cudaEventRecord(ev, cs);
	if (e != cudaSuccess) {
		cudaEventDestroy(ev);
		CU_TRY(e);
	}
	*ev_out = ev;
	return CUBIT_OK;
}

extern "C" int cubit_gpu_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids,
                               uint32_t n_cols, void *const *host_cols) {
	ABI_BEGIN
	if (!r) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (!r->parts.empty()) {
		return sharded_fetch(r, offset, n, host_rowids, n_cols, host_cols, nullptr);
	}
	rc = fetch_check(r, offset, n, host_rowids, n_cols);
	if (rc || n == 0) {
		return rc;
	}
	cudaEvent_t ev = nullptr;
	rc = fetch_enqueue(r, offset, n, host_rowids, n_cols, host_cols, &ev);
	if (rc) {
		return rc;
	}
	cudaError_t e = cudaEventSynchronize(ev);
	cudaEventDestroy(ev);
	CU_TRY(e);
	return CUBIT_OK;
	ABI_END
}

// The same copy without the wait: window i+1 of a result crosses PCIe while the caller consumes window i (and
// while the kernel stream runs the next query).  Host buffers should be page-locked (cubit_gpu_alloc_host) — a
// pageable destination makes the copy synchronous.  cubit_gpu_fetch_wait completes and frees the ticket.
extern "C" int cubit_gpu_fetch_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids,
                                     uint32_t n_cols, void *const *host_cols, cubit_gpu_fetch_ticket **ticket) {
	ABI_BEGIN
	if (!r || !ticket) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*ticket = nullptr;
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (!r->parts.empty()) {
		return sharded_fetch(r, offset, n, host_rowids, n_cols, host_cols, ticket);
	}
	rc = fetch_check(r, offset, n, host_rowids, n_cols);
	if (rc) {
		return rc;
	}
	cubit_gpu_fetch_ticket *tk = new cubit_gpu_fetch_ticket();
	if (n) {
		r->copies_in_flight.store(1); // free_result drains the copy streams before the buffers go back to the pool
		rc = fetch_enqueue(r, offset, n, host_rowids, n_cols, host_cols, &tk->ev);
		if (rc) {
			delete tk;
			return rc;
		}
	}
	*ticket = tk;
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_fetch_wait(cubit_gpu_fetch_ticket *ticket) {
	ABI_BEGIN
	if (!ticket) {
		return CUBIT_OK;
	}
	int rc = CUBIT_OK;
	for (auto *p : ticket->parts) {
		const int prc = cubit_gpu_fetch_wait(p);
		rc = rc ? rc : prc;
	}
	if (ticket->ev) {
		cudaError_t e = cudaEventSynchronize(ticket->ev);
		cudaEventDestroy(ticket->ev);
		if (e != cudaSuccess && rc == CUBIT_OK) {
			rc = fail(CUBIT_ECUDA, "fetch: %s", cudaGetErrorString(e));
		}
	}
	delete ticket;
	return rc;
	ABI_END
}

// Validity of projected column `col` for result rows [offset, offset + n): bit j of host_words = row offset + j
// (1 = valid), ceil(n / 64) words, bits past n zero — the mask a DataChunk vector carries (vector.hpp:242-256).
extern "C" int cubit_gpu_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n,
                                        uint64_t *host_words, int *all_valid) {
	ABI_BEGIN
	if (!r || (!host_words && !all_valid)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (!r->parts.empty()) {
		return sharded_fetch_validity(r, col, offset, n, host_words, all_valid);
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
		CU_TRY(cudaSetDevice(r->t->device));
		const uint64_t avail32 = (r->info.capacity + 31) / 32 + 2; // words allocated
		uint64_t copy32 = src_words * 2;
		if (w0 * 2 + copy32 > avail32) {
			copy32 = avail32 - w0 * 2;
		}
		cudaStream_t cs = r->t->copy_stream[r->t->next_copy.fetch_add(1) % kCopyStreams];
		CU_TRY(cudaMemcpyAsync(tmp.data(), r->d_valid[col] + w0 * 2, copy32 * 4, cudaMemcpyDeviceToHost, cs));
		CU_TRY(cudaStreamSynchronize(cs));
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
	ABI_END
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
	ABI_BEGIN
	if (!r || !host_words) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	int rc = finish_result(r);
	if (rc) {
		return rc;
	}
	if (!r->parts.empty()) {
		return sharded_fetch_bitvector(r, host_words, n_words);
	}
	if (!r->d_q) {
		return fail(CUBIT_ESTATE, "query did not materialise the bitvector (CUBIT_Q_BITVECTOR)");
	}
	if (n_words != r->t->n_words) {
		return fail(CUBIT_EINVAL, "n_words mismatch");
	}
	CU_TRY(cudaSetDevice(r->t->device));
	cudaStream_t cs = r->t->copy_stream[r->t->next_copy.fetch_add(1) % kCopyStreams];
	CU_TRY(cudaMemcpyAsync(host_words, r->d_q, n_words * 8, cudaMemcpyDeviceToHost, cs));
	CU_TRY(cudaStreamSynchronize(cs));
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_free_result(cubit_gpu_result *r) {
	if (!r) {
		return CUBIT_OK;
	}
	if (!r->parts.empty()) {
		return sharded_free_result(r);
	}
	cudaSetDevice(r->t->device);
	std::shared_lock<std::shared_mutex> lk(r->t->mu);
	release_result_locked(r);
	return CUBIT_OK;
}

// -------------------------------------------------------------------- probe
extern "C" int cubit_gpu_probe(cubit_gpu_table *t, int32_t col_id, const int64_t *host_rowids, uint64_t n,
                               void *host_out, uint64_t *sum_lo, int64_t *sum_hi) {
	ABI_BEGIN
	if (!t || (n && !host_rowids)) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		return sharded_probe(t, col_id, host_rowids, n, host_out, sum_lo, sum_hi);
	}
	std::shared_lock<std::shared_mutex> lk(t->mu);
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
	pa.agg_valid_a = want_sum ? c.d_valid : nullptr; // SUM skips NULL inputs (sum.cpp: only valid rows reach the state)
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
	ABI_END
}
