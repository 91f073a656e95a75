// duckdb_cfg2_baseline.cpp — the REFERENCE's own CPU path on the bench workload (config 2), for bench.py's
// cpu_baseline / reference arm: the unmodified reference DuckDB (libduckdb.so built from /root/reference) scans a
// bounded sample of the synthetic table with the same six range predicates
//     SELECT count(*), sum(payload) FROM t WHERE v_i BETWEEN 10 AND 19
// (seq_scan + pushed-down filter, src/function/table/table_scan.cpp:119-146 → TemplatedFilterSelection,
// src/storage/table/column_segment.cpp:261-276; the tree holds no CUBIT source, so this IS its path for the query).
// The table is filled from the oracle's generator (the same SplitMix64 columns the GPU synthesises), so the
// answers must equal the oracle's — bench.py checks that.  Usage: duckdb_cfg2_baseline <rows> <reps> <thr0> .. <thr5>
#include "duckdb.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

extern "C" void oracle_synth_column(void *col, int kind, uint64_t n_rows, int64_t row_base, uint64_t seed, uint64_t threshold,
                                    uint32_t card, uint32_t hot_lo, uint32_t hot_n);
using namespace duckdb;

static unique_ptr<MaterializedQueryResult> Run(Connection &con, const std::string &sql) {
	auto r = con.Query(sql);
	if (r->HasError()) {
		fprintf(stderr, "SQL failed: %s\n%s\n", sql.c_str(), r->GetError().c_str());
		exit(1);
	}
	return r;
}

int main(int argc, char **argv) {
	if (argc < 4) {
		fprintf(stderr, "usage: %s rows reps threshold...\n", argv[0]);
		return 2;
	}
	const uint64_t rows = strtoull(argv[1], nullptr, 10);
	const int reps = atoi(argv[2]);
	std::vector<uint64_t> thr;
	for (int i = 3; i < argc; i++) {
		thr.push_back(strtoull(argv[i], nullptr, 10));
	}
	const size_t nq = thr.size();
	DuckDB db(nullptr);
	Connection con(db);
	std::string ddl = "CREATE TABLE t(";
	for (size_t i = 0; i < nq; i++) {
		ddl += "v" + std::to_string(i) + " INTEGER, ";
	}
	Run(con, ddl + "payload BIGINT)");
	{
		std::vector<std::vector<int32_t>> v(nq, std::vector<int32_t>(rows));
		std::vector<int64_t> payload(rows);
		for (size_t i = 0; i < nq; i++) {
			oracle_synth_column(v[i].data(), 1, rows, 0, 0xC0B17, thr[i], 100, 10, 10);
		}
		oracle_synth_column(payload.data(), 0, rows, 0, 0, 0, 0, 0, 0);
		vector<LogicalType> types(nq, LogicalType::INTEGER);
		types.push_back(LogicalType::BIGINT);
		DataChunk chunk;
		chunk.Initialize(Allocator::DefaultAllocator(), types);
		Appender app(con, "t");
		for (uint64_t r0 = 0; r0 < rows; r0 += STANDARD_VECTOR_SIZE) {
			const idx_t n = (idx_t)std::min<uint64_t>(STANDARD_VECTOR_SIZE, rows - r0);
			chunk.Reset();
			for (size_t i = 0; i < nq; i++) {
				memcpy(FlatVector::GetData<int32_t>(chunk.data[i]), v[i].data() + r0, n * 4);
			}
			memcpy(FlatVector::GetData<int64_t>(chunk.data[nq]), payload.data() + r0, n * 8);
			chunk.SetCardinality(n);
			app.AppendDataChunk(chunk);
		}
		app.Close();
	}
	const unsigned cores = std::thread::hardware_concurrency();
	Run(con, "SET threads=" + std::to_string(cores));
	std::vector<double> ms;
	std::string answers;
	for (int it = 0; it < reps + 1; it++) {
		auto t0 = std::chrono::steady_clock::now();
		answers = "";
		for (size_t i = 0; i < nq; i++) {
			auto r = Run(con, "SELECT count(*), sum(payload) FROM t WHERE v" + std::to_string(i) + " BETWEEN 10 AND 19");
			answers += std::string(i ? ", " : "") + "[" + r->GetValue(0, 0).ToString() + ", " +
			           (r->GetValue(1, 0).IsNull() ? "0" : r->GetValue(1, 0).ToString()) + "]";
		}
		auto t1 = std::chrono::steady_clock::now();
		if (it > 0) {
			ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
		}
	}
	std::sort(ms.begin(), ms.end());
	const double med = ms[ms.size() / 2];
	printf("{\"rows\": %llu, \"threads\": %u, \"sweep_ms\": %.3f, \"rows_per_s\": %.1f, \"answers\": [%s], \"version\": \"%s\"}\n",
	       (unsigned long long)rows, cores, med, (double)nq * (double)rows / (med * 1e-3), answers.c_str(), DuckDB::LibraryVersion());
	return 0;
}
