// host_scan_test.cpp — drives the C++ host layer the way PhysicalTableScan drives a table
// function (bind → init_global → GetData until FINISHED) and checks every chunk against a
// brute-force evaluation of the same predicate over the columns.  Mirrors the shape of the
// reference's own scan tests (test/sql/index/art/scan/test_art_many_matches.test: exact
// counts for =, <, <=, >, >= over duplicates; test/api/capi/capi_table_functions.cpp).
// Exit code 0 = all checks passed.  Needs a B200 (no CPU fallback).
#include "cubit_scan.hpp"

#include "cubit_gpu.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

using namespace cubit_host;

#define REQUIRE(cond)                                                                                                  \
	do {                                                                                                               \
		if (!(cond)) {                                                                                                 \
			fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);                                 \
			exit(1);                                                                                                   \
		}                                                                                                              \
	} while (0)

static uint64_t rng_state = 0x1234567;
static uint64_t Rng() {
	rng_state += 0x9E3779B97F4A7C15ull;
	uint64_t z = rng_state;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

static void ArtManyMatches(idx_t reps) {
	// [0,1,0,1,...]: i<1, i<=1, i=0, i=1, i>0, i>=0 (test_art_many_matches.test)
	std::vector<int32_t> col(2 * reps);
	for (idx_t r = 0; r < col.size(); r++) {
		col[r] = (int32_t)(r & 1);
	}
	CubitTable table(col.size());
	table.AddColumn(0, col.data());
	CubitIndex index(table, 0, 0, 2);
	index.Build();
	std::vector<row_t> ids;
	REQUIRE(index.Scan(0, 0, 1 << 30, ids) && ids.size() == reps);
	REQUIRE(index.Scan(0, 1, 1 << 30, ids) && ids.size() == 2 * reps);
	REQUIRE(index.Scan(1, 1, 1 << 30, ids) && ids.size() == reps);
	for (idx_t i = 0; i < ids.size(); i++) {
		REQUIRE(ids[i] == (row_t)(2 * i + 1)); // sorted, unique (art.cpp:974-985)
	}
	REQUIRE(index.Scan(1, 100, 1 << 30, ids) && ids.size() == reps);  // i > 0
	REQUIRE(index.Scan(-5, 100, 1 << 30, ids) && ids.size() == 2 * reps); // i >= 0
	REQUIRE(!index.Scan(0, 1, reps, ids));                                // more than max_count → false
	REQUIRE(index.Scan(7, 9, 10, ids) && ids.empty());
}

// Several host threads on ONE table: the table's lock covers planning and enqueueing only, so the
// config-1 style query (equality predicate + SUM: ~13 us of kernel inside ~40 us of call) overlaps across threads.
// Straight through the C-ABI, as a DuckDB worker thread would call it.
static void ConcurrentQueriesOverlap() {
	const idx_t n = 6001215; // TPC-H SF1 lineitem rows
	std::vector<int32_t> key(n);
	std::vector<int64_t> pay(n);
	std::vector<int64_t> want_sum(50, 0), want_cnt(50, 0);
	for (idx_t r = 0; r < n; r++) {
		key[r] = (int32_t)(Rng() % 50);
		pay[r] = (int64_t)(Rng() % 10000000) - 5000;
		want_sum[key[r]] += pay[r];
		want_cnt[key[r]]++;
	}
	CubitTable table(n);
	table.AddColumn(0, key.data());
	table.AddColumn(1, pay.data());
	CubitIndex index(table, 0, 0, 50);
	index.Build();
	std::atomic<int> bad {0};
	auto run = [&](int n_threads, int per_thread) {
		std::vector<std::thread> th;
		const auto t0 = std::chrono::steady_clock::now();
		for (int k = 0; k < n_threads; k++) {
			th.emplace_back([&, k]() {
				for (int i = 0; i < per_thread; i++) {
					const uint32_t v = (uint32_t)((k * 7 + i) % 50);
					cubit_bv_ref ref {index.Id(), v};
					cubit_pred_group grp {1, &ref};
					cubit_query q {};
					q.n_groups = 1;
					q.groups = &grp;
					q.agg_kind = CUBIT_AGG_SUM;
					q.agg_col_a = 1;
					cubit_gpu_result *res = nullptr;
					cubit_result_info info;
					if (cubit_gpu_query(table.Handle(), &q, &res) != CUBIT_OK || cubit_gpu_result_get(res, &info) != CUBIT_OK ||
					    (int64_t)info.count != want_cnt[v] || (int64_t)info.sum_lo != want_sum[v]) {
						bad++;
					}
					cubit_gpu_free_result(res);
				}
			});
		}
		for (auto &t : th) {
			t.join();
		}
		const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		return n_threads * per_thread / s;
	};
	run(2, 100);
	const double one = run(1, 2000), eight = run(8, 2000);
	printf("concurrent config-1 queries on one table: 1 thread %.0f/s, 8 threads %.0f/s (%.2fx)\n", one, eight, eight / one);
	REQUIRE(bad.load() == 0);
	REQUIRE(eight > 2.5 * one); // measured 3.3-3.6x on a B200 box with 16 host cores (profiles/r2_concurrency.md)
}

// INSERT after CREATE INDEX: appended rows take the next row ids and show up in index scans
// (test/sql/index/art/insert_update_delete: scans after inserts see the new keys)
static void AppendAfterBuild() {
	const idx_t n0 = 70000, n1 = 130001; // crosses a 64-row word, a 4096-row build tile and a 65536-row segment
	std::vector<int32_t> key(n0 + n1);
	std::vector<int64_t> pay(n0 + n1);
	for (idx_t r = 0; r < key.size(); r++) {
		key[r] = (int32_t)(Rng() % 7);
		pay[r] = (int64_t)(Rng() % 1000003) - 500000;
	}
	CubitTable table(n0, 4096);
	table.AddColumn(0, key.data());
	table.AddColumn(1, pay.data());
	CubitIndex index(table, 0, 0, 7);
	index.Build();
	index.Delete(4096 + 5, key[5]);
	index.CommitDeltas();
	table.Append(n1, {{0, key.data() + n0}, {1, pay.data() + n0}});
	REQUIRE(table.RowCount() == n0 + n1);
	std::vector<row_t> ids, want;
	for (idx_t r = 0; r < key.size(); r++) {
		if (key[r] >= 2 && key[r] <= 3 && r != 5) {
			want.push_back((row_t)(4096 + r));
		} else if (r == 5 && !(key[r] >= 2 && key[r] <= 3)) {
			// deleted row of another key: not in range anyway
		}
	}
	REQUIRE(index.Scan(2, 3, 1 << 30, ids));
	REQUIRE(ids == want);
}

// cubit_gpu_drain with a C++ consumer on four library worker threads at once: every DataChunk exactly once, the batch
// index orders them (PhysicalTableScan's ordered parallel source, table_scan.cpp:179-189), values = brute force
struct DrainSink {
	std::mutex mu;
	std::map<std::pair<uint64_t, uint64_t>, std::pair<std::vector<int64_t>, std::vector<int64_t>>> chunks;
	std::atomic<int> in_flight {0}, max_in_flight {0};
};
static int DrainCallback(void *ctx, uint32_t, uint64_t batch, uint64_t row_offset, uint32_t n, const int64_t *rowids,
                         const void *const *cols, const uint64_t *const *validity) {
	auto &sink = *static_cast<DrainSink *>(ctx);
	const int now = ++sink.in_flight;
	int seen = sink.max_in_flight.load();
	while (now > seen && !sink.max_in_flight.compare_exchange_weak(seen, now)) {
	}
	std::vector<int64_t> ids(rowids, rowids + n), vals(static_cast<const int64_t *>(cols[0]), static_cast<const int64_t *>(cols[0]) + n);
	REQUIRE(validity[0] == nullptr);
	{
		std::lock_guard<std::mutex> lk(sink.mu);
		REQUIRE(sink.chunks.emplace(std::make_pair(batch, row_offset), std::make_pair(std::move(ids), std::move(vals))).second);
	}
	--sink.in_flight;
	return 0;
}
static void ParallelDrain() {
	const idx_t n = 2000003;
	const row_t base = 65536 * 2;
	std::vector<int32_t> key(n);
	std::vector<int64_t> pay(n);
	for (idx_t r = 0; r < n; r++) {
		key[r] = (int32_t)(Rng() % 5);
		pay[r] = (int64_t)(Rng() % 100000) * 1000 - 7;
	}
	CubitTable table(n, base);
	table.AddColumn(0, key.data());
	table.AddColumn(1, pay.data());
	CubitIndex index(table, 0, 0, 5);
	index.Build();
	cubit_bv_ref refs[2] = {{index.Id(), 1}, {index.Id(), 3}};
	cubit_pred_group grp {2, refs};
	int32_t col = 1;
	cubit_query q {};
	q.n_groups = 1;
	q.groups = &grp;
	q.flags = CUBIT_Q_ROWIDS | CUBIT_Q_VALUES;
	q.n_cols = 1;
	q.cols = &col;
	cubit_gpu_result *r = nullptr;
	REQUIRE(cubit_gpu_query(table.Handle(), &q, &r) == CUBIT_OK);
	DrainSink sink;
	cubit_drain_stats st;
	REQUIRE(cubit_gpu_drain(r, 1, 1, 4, 16 * 2048, DrainCallback, &sink, &st) == CUBIT_OK);
	std::vector<int64_t> want_ids, want_vals;
	for (idx_t i = 0; i < n; i++) {
		if (key[i] == 1 || key[i] == 3) {
			want_ids.push_back(base + (row_t)i);
			want_vals.push_back(pay[i]);
		}
	}
	REQUIRE(st.rows == want_ids.size() && st.workers == 4 && st.wire_bytes < st.wide_bytes);
	std::vector<int64_t> got_ids, got_vals;
	uint64_t next_row = 0;
	for (auto &kv : sink.chunks) { // (batch, row offset) order = row order
		REQUIRE(kv.first.second == next_row);
		next_row += kv.second.first.size();
		got_ids.insert(got_ids.end(), kv.second.first.begin(), kv.second.first.end());
		got_vals.insert(got_vals.end(), kv.second.second.begin(), kv.second.second.end());
	}
	REQUIRE(got_ids == want_ids && got_vals == want_vals);
	printf("parallel drain: %llu rows in %llu chunks, %llu wire bytes (%.1f per row, wide %.0f), up to %d consumers at once\n",
	       (unsigned long long)st.rows, (unsigned long long)st.chunks, (unsigned long long)st.wire_bytes,
	       (double)st.wire_bytes / (double)st.rows, (double)st.wide_bytes / (double)st.rows, sink.max_in_flight.load());
	cubit_gpu_free_result(r);
}

// NULLs in a projected column: the chunk's vector carries the validity mask of the probed rows
static void NullsInProjectedColumn() {
	const idx_t n = 300007;
	std::vector<int32_t> key(n);
	std::vector<int64_t> val(n);
	std::vector<uint64_t> valid((n + 63) / 64, 0);
	for (idx_t r = 0; r < n; r++) {
		key[r] = (int32_t)(r % 4);
		val[r] = (int64_t)r * 3;
		if (r % 5 != 2) {
			valid[r / 64] |= 1ull << (r % 64);
		}
	}
	CubitTable table(n, 0);
	table.AddColumn(0, key.data());
	table.AddColumn(1, val.data());
	table.SetValidity(1, valid.data());
	CubitIndex ix(table, 0, 0, 4);
	ix.Build();
	auto bind = CubitScanBind(table, {{&ix, 1, 2}});
	std::vector<column_t> column_ids = {COLUMN_IDENTIFIER_ROW_ID, 1};
	auto gstate = CubitScanInitGlobal(*bind, column_ids);
	DataChunk chunk;
	chunk.Initialize(CubitScanReturnTypes(*bind, column_ids));
	idx_t seen = 0, nulls = 0;
	while (CubitScanGetData(*bind, *gstate, chunk) == SourceResultType::HAVE_MORE_OUTPUT) {
		for (idx_t i = 0; i < chunk.size(); i++, seen++) {
			const row_t r = chunk.data[0].GetData<row_t>()[i];
			REQUIRE(r % 4 == 1 || r % 4 == 2);
			REQUIRE(chunk.data[0].RowIsValid(i));
			REQUIRE(chunk.data[1].RowIsValid(i) == (r % 5 != 2));
			if (chunk.data[1].RowIsValid(i)) {
				REQUIRE(chunk.data[1].GetData<int64_t>()[i] == r * 3);
			} else {
				nulls++;
			}
		}
	}
	REQUIRE(seen == gstate->row_count && seen == 150004 && nulls > 0);
}

int main() {
	ParallelDrain();
	NullsInProjectedColumn();
	AppendAfterBuild();
	ArtManyMatches(1024);
	ArtManyMatches(2048);

	const idx_t n = 1000003;
	const row_t row_base = 65536 * 4;
	std::vector<int32_t> qty(n), disc(n);
	std::vector<int64_t> price(n);
	for (idx_t r = 0; r < n; r++) {
		qty[r] = 1 + (int32_t)(Rng() % 50);
		disc[r] = (int32_t)(Rng() % 11);
		price[r] = 90000 + (int64_t)(Rng() % 10000000);
	}
	std::vector<int64_t> disc64(disc.begin(), disc.end());
	CubitTable table(n, row_base);
	table.AddColumn(0, qty.data());
	table.AddColumn(1, disc.data());
	table.AddColumn(2, price.data());
	table.AddColumn(3, disc64.data());
	CubitIndex iq(table, 0, 1, 50), id(table, 1, 0, 11);
	iq.Build();
	id.Build();

	// pending deltas: updates and deletes, XOR-ed at query time
	std::vector<bool> deleted(n, false);
	for (int i = 0; i < 5000; i++) {
		const idx_t r = Rng() % n;
		if (deleted[r]) {
			continue;
		}
		if (i & 1) {
			const int32_t nv = 1 + (int32_t)(Rng() % 50);
			iq.Update(row_base + (row_t)r, qty[r], nv);
			qty[r] = nv;
		} else {
			iq.Delete(row_base + (row_t)r, qty[r]);
			id.Delete(row_base + (row_t)r, disc[r]);
			deleted[r] = true;
		}
	}
	iq.CommitDeltas();
	id.CommitDeltas();
	REQUIRE(iq.PendingDeltaRows() > 0);

	for (int round = 0; round < 2; round++) {
		// WHERE qty BETWEEN 10 AND 19 AND disc BETWEEN 5 AND 7 → rowid, price, disc
		auto bind = CubitScanBind(table, {{&iq, 10, 19}, {&id, 5, 7}});
		std::vector<column_t> column_ids = {COLUMN_IDENTIFIER_ROW_ID, 2, 1};
		auto gstate = CubitScanInitGlobal(*bind, column_ids);
		REQUIRE(gstate->MaxThreads() == 1);
		DataChunk chunk;
		chunk.Initialize(CubitScanReturnTypes(*bind, column_ids));
		idx_t r = 0, seen = 0, calls = 0;
		__int128 expect_sum = 0;
		idx_t expect_count = 0;
		while (CubitScanGetData(*bind, *gstate, chunk) == SourceResultType::HAVE_MORE_OUTPUT) {
			calls++;
			REQUIRE(chunk.size() <= STANDARD_VECTOR_SIZE);
			for (idx_t i = 0; i < chunk.size(); i++, seen++) {
				while (r < n && (deleted[r] || qty[r] < 10 || qty[r] > 19 || disc[r] < 5 || disc[r] > 7)) {
					r++;
				}
				REQUIRE(r < n);
				REQUIRE(chunk.data[0].GetData<row_t>()[i] == row_base + (row_t)r);
				REQUIRE(chunk.data[1].GetData<int64_t>()[i] == price[r]);
				REQUIRE(chunk.data[2].GetData<int32_t>()[i] == disc[r]);
				expect_sum += (__int128)price[r] * disc[r];
				expect_count++;
				r++;
			}
		}
		while (r < n && (deleted[r] || qty[r] < 10 || qty[r] > 19 || disc[r] < 5 || disc[r] > 7)) {
			r++;
		}
		REQUIRE(r == n && seen == gstate->row_count && seen > 0);
		REQUIRE(calls >= (seen + STANDARD_VECTOR_SIZE - 1) / STANDARD_VECTOR_SIZE);

		// aggregate push-down: SELECT count(*), sum(price * disc) → one row
		auto abind = CubitScanBind(table, {{&iq, 10, 19}, {&id, 5, 7}}, CubitAggregate::SUM_PRODUCT, 2, 3);
		auto astate = CubitScanInitGlobal(*abind, {});
		DataChunk arow;
		arow.Initialize(CubitScanReturnTypes(*abind, {}));
		REQUIRE(CubitScanGetData(*abind, *astate, arow) == SourceResultType::HAVE_MORE_OUTPUT && arow.size() == 1);
		REQUIRE((idx_t)arow.data[0].GetData<int64_t>()[0] == expect_count);
		const __int128 got = ((__int128)arow.data[2].GetData<int64_t>()[0] << 64) |
		                     (unsigned __int128)(uint64_t)arow.data[1].GetData<int64_t>()[0];
		REQUIRE(got == expect_sum);
		REQUIRE(CubitScanGetData(*abind, *astate, arow) == SourceResultType::FINISHED);

		if (round == 0) { // merge-back must not change any answer
			iq.MergeDeltas();
			id.MergeDeltas();
			REQUIRE(iq.PendingDeltaRows() == 0);
		}
	}

	// errors are exceptions, as in the reference's table functions
	bool threw = false;
	try {
		CubitScanBind(table, {});
	} catch (const InvalidInputException &) {
		threw = true;
	}
	REQUIRE(threw);
	threw = false;
	try {
		iq.Delete(row_base, 77);
	} catch (const InvalidInputException &) {
		threw = true;
	}
	REQUIRE(threw);
	ConcurrentQueriesOverlap();
	printf("host_scan_test ok\n");
	return 0;
}
