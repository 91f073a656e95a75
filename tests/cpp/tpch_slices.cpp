// tpch_slices.cpp — TEST / BENCH INFRASTRUCTURE: TPC-H lineitem slices from the REFERENCE's own dbgen
// (extension/tpch/tpch_extension.cpp:46-54 → dbgen(sf, children, step), dbgen.cpp:626-634: one slice generated
// single-threaded, deterministic, in canonical row order), for BASELINE config 3 (SF100 does not fit a box as one
// DuckDB table next to the GPU copy; slice by slice it needs one slice of host memory).
//
//   tpch_slices <sf> <children> <first_step> <step_stride> <out_dir>
// For every step s = first_step, first_step + stride, ... < children:
//   * a fresh in-memory database of the bundled libduckdb.so runs CALL dbgen(sf, children, step = s)
//   * the four columns the Q6-style scan needs are written RAW to <out_dir>/slice_<s>.bin
//       header  u64 n_rows
//       int64   l_quantity      [n]   DECIMAL(15,2) cents (dbgen.cpp:48-50)
//       int64   l_extendedprice [n]
//       int64   l_discount      [n]
//       int32   l_shipdate      [n]   days since 1970-01-01
//     (written to a temporary name and renamed, so a reader never sees a partial file)
//   * the REFERENCE's own answers on the slice go to <out_dir>/slice_<s>.json: row count, TPC-H Q6 as written
//     (count and revenue), and the config-1 query (l_quantity = 24: count, sum(l_extendedprice)).  Summed over all
//     slices they are what the unmodified reference answers on the whole table — the oracle for the GPU run.
// Links only libduckdb.so (no GPU library, no glue).
#include "duckdb.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

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
	if (argc < 6) {
		fprintf(stderr, "usage: tpch_slices <sf> <children> <first_step> <step_stride> <out_dir>\n");
		return 2;
	}
	const std::string sf = argv[1];
	const int children = atoi(argv[2]), first = atoi(argv[3]), stride = atoi(argv[4]);
	const std::string dir = argv[5];
	for (int step = first; step < children; step += stride) {
		DuckDB db(nullptr);
		Connection con(db);
		Run(con, "SET threads = 1");
		if (children > 1) {
			Run(con, "CALL dbgen(sf = " + sf + ", children = " + std::to_string(children) + ", step = " + std::to_string(step) + ")");
		} else {
			Run(con, "CALL dbgen(sf = " + sf + ")");
		}
		auto res = Run(con, "SELECT l_quantity, l_extendedprice, l_discount, l_shipdate FROM lineitem");
		const idx_t n = res->RowCount();
		std::vector<int64_t> q(n), p(n), d(n);
		std::vector<int32_t> s(n);
		idx_t at = 0;
		for (auto &chunk : res->Collection().Chunks()) {
			const idx_t c = chunk.size();
			for (int k = 0; k < 4; k++) {
				chunk.data[k].Flatten(c);
			}
			// DECIMAL(15,2) is physically int64 (unscaled), DATE is int32 days
			memcpy(q.data() + at, FlatVector::GetData<int64_t>(chunk.data[0]), c * 8);
			memcpy(p.data() + at, FlatVector::GetData<int64_t>(chunk.data[1]), c * 8);
			memcpy(d.data() + at, FlatVector::GetData<int64_t>(chunk.data[2]), c * 8);
			memcpy(s.data() + at, FlatVector::GetData<int32_t>(chunk.data[3]), c * 4);
			at += c;
		}
		auto q6 = Run(con, "SELECT count(*), sum(l_extendedprice * l_discount) FROM lineitem WHERE l_shipdate >= CAST('1994-01-01' AS date) "
		                   "AND l_shipdate < CAST('1995-01-01' AS date) AND l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24");
		auto c1 = Run(con, "SELECT count(*), sum(l_extendedprice) FROM lineitem WHERE l_quantity = 24");
		const std::string base = dir + "/slice_" + std::to_string(step);
		{
			FILE *f = fopen((base + ".bin.tmp").c_str(), "wb");
			if (!f) {
				perror("fopen");
				return 1;
			}
			const uint64_t n64 = n;
			fwrite(&n64, 8, 1, f);
			fwrite(q.data(), 8, n, f);
			fwrite(p.data(), 8, n, f);
			fwrite(d.data(), 8, n, f);
			fwrite(s.data(), 4, n, f);
			fclose(f);
		}
		{
			FILE *f = fopen((base + ".json.tmp").c_str(), "w");
			auto val = [](MaterializedQueryResult &r, idx_t c) {
				return r.GetValue(c, 0).IsNull() ? std::string("0") : r.GetValue(c, 0).ToString();
			};
			fprintf(f, "{\"step\": %d, \"rows\": %llu, \"q6_count\": %s, \"q6_revenue\": \"%s\", \"q24_count\": %s, \"q24_sum_price\": \"%s\"}\n",
			        step, (unsigned long long)n, val(*q6, 0).c_str(), val(*q6, 1).c_str(), val(*c1, 0).c_str(), val(*c1, 1).c_str());
			fclose(f);
		}
		rename((base + ".json.tmp").c_str(), (base + ".json").c_str());
		rename((base + ".bin.tmp").c_str(), (base + ".bin").c_str()); // the .bin appears last: both files are complete
	}
	return 0;
}
