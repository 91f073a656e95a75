// duckdb_sql_test.cpp — the integration glue (integration/duckdb_cubit_extension.cpp) loaded into the REAL
// reference DuckDB: the same SQL through cubit_scan / cubit_agg and through the vanilla scan must agree.
// Runs only where the reference build exists (the build container); see tests/test_duckdb_integration.py.
#include "duckdb.hpp"
#include "cubit_gpu.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace duckdb {
void RegisterCubitGpuFunctions(DatabaseInstance &db);
idx_t CubitRewriteCount();
idx_t CubitSegmentRouteCount();
idx_t CubitAggPushdownCount();
idx_t CubitDmlAppendedRows();
idx_t CubitDmlDeltaPairs();
idx_t CubitImageLoads();
}
using namespace duckdb;

#define REQUIRE(cond)                                                                                                  \
	do {                                                                                                               \
		if (!(cond)) {                                                                                                 \
			fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);                                 \
			exit(1);                                                                                                   \
		}                                                                                                              \
	} while (0)

static unique_ptr<MaterializedQueryResult> Run(Connection &con, const string &sql) {
	auto r = con.Query(sql);
	if (r->HasError()) {
		fprintf(stderr, "SQL failed: %s\n%s\n", sql.c_str(), r->GetError().c_str());
		exit(1);
	}
	return r;
}

// the same statement through the vanilla scan of the SAME table (the optimizer rewrite switched off for one query)
static unique_ptr<MaterializedQueryResult> RunVanilla(Connection &con, const string &sql) {
	setenv("CUBIT_DISABLE_REWRITE", "1", 1);
	auto r = Run(con, sql);
	unsetenv("CUBIT_DISABLE_REWRITE");
	return r;
}

// every predicate, through the GPU (rewritten: the rewrite counter moves) and through the vanilla scan of the same
// table: aggregates, and rows with their row ids
static void CompareWithVanilla(Connection &con, const string &table, const string &cols, const vector<string> &wheres) {
	for (auto &where : wheres) {
		const string agg = "SELECT count(*), sum(price), sum(price * disc), min(rowid), max(rowid) FROM " + table + " WHERE " + where;
		const idx_t before = CubitRewriteCount();
		auto a = Run(con, agg);
		auto b = RunVanilla(con, agg);
		if (b->GetValue(0, 0).GetValue<int64_t>() > 0) {
			REQUIRE(CubitRewriteCount() == before + 1);
		}
		for (idx_t c = 0; c < 5; c++) {
			if (a->GetValue(c, 0).ToString() != b->GetValue(c, 0).ToString()) {
				fprintf(stderr, "mismatch on %s WHERE %s col %llu: gpu %s vanilla %s\n", table.c_str(), where.c_str(),
				        (unsigned long long)c, a->GetValue(c, 0).ToString().c_str(), b->GetValue(c, 0).ToString().c_str());
			}
			REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
		}
		const string rows = "SELECT rowid, " + cols + " FROM " + table + " WHERE " + where + " ORDER BY rowid";
		auto x = Run(con, rows);
		auto y = RunVanilla(con, rows);
		REQUIRE(x->RowCount() == y->RowCount());
		for (idx_t r = 0; r < x->RowCount(); r += 41) {
			for (idx_t c = 0; c < x->ColumnCount(); c++) {
				REQUIRE(x->GetValue(c, r) == y->GetValue(c, r));
			}
		}
	}
}

int main(int argc, char **argv) {
	const bool expect_no_device = argc > 1 && string(argv[1]) == "--expect-no-device";
	DuckDB db(nullptr);
	Connection con(db);
	RegisterCubitGpuFunctions(*db.instance);
	Run(con, "CREATE TABLE t AS SELECT (i * 7919 % 50 + 1)::BIGINT AS q, (i * 104729 % 1000003 - 500000)::BIGINT AS price, "
	         "(i % 11)::BIGINT AS disc FROM range(300000) r(i)");
	if (expect_no_device) {
		auto r = con.Query("CALL cubit_load('t', 'q', 1, 50)");
		REQUIRE(r->HasError());
		REQUIRE(r->GetError().find("no CUDA device") != string::npos); // the C-ABI error became a DuckDB exception
		printf("duckdb_sql_test ok (no device: error propagated)\n");
		return 0;
	}
	auto loaded = Run(con, "CALL cubit_load('t', 'q', 1, 50)");
	REQUIRE(loaded->GetValue(0, 0).GetValue<int64_t>() == 300000);
	const char *preds[][3] = {{"10", "19", "q BETWEEN 10 AND 19"}, {"24", "24", "q = 24"}, {"-5", "3", "q <= 3"},
	                          {"48", "900", "q >= 48"}, {"60", "70", "q BETWEEN 60 AND 70"}};
	for (auto &p : preds) {
		const string lo = p[0], hi = p[1], where = p[2];
		auto a = Run(con, "SELECT count(*), sum(price), sum(price * disc) FROM cubit_scan('t', " + lo + ", " + hi + ")");
		auto b = Run(con, "SELECT count(*), sum(price), sum(price * disc) FROM t WHERE " + where);
		for (idx_t c = 0; c < 3; c++) {
			REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
		}
		// row-for-row, in row-id order
		auto x = Run(con, "SELECT q, price FROM cubit_scan('t', " + lo + ", " + hi + ")");
		auto y = Run(con, "SELECT q, price FROM t WHERE " + where + " ORDER BY rowid");
		REQUIRE(x->RowCount() == y->RowCount());
		for (idx_t r = 0; r < x->RowCount(); r++) {
			REQUIRE(x->GetValue(0, r) == y->GetValue(0, r) && x->GetValue(1, r) == y->GetValue(1, r));
		}
		// aggregate push-down: one row
		auto g = Run(con, "SELECT * FROM cubit_agg('t', " + lo + ", " + hi + ", 'price')");
		REQUIRE(g->RowCount() == 1);
		REQUIRE(g->GetValue(0, 0).ToString() == b->GetValue(0, 0).ToString());
		if (b->GetValue(0, 0).GetValue<int64_t>() > 0) {
			REQUIRE(g->GetValue(1, 0).ToString() == b->GetValue(1, 0).ToString());
		}
	}
	// transparent rewrite: plain SQL on the indexed table is re-pointed at the GPU scan by the optimizer
	// extension; t_plain holds the same rows but has no GPU index, so it takes the vanilla seq_scan
	Run(con, "CREATE TABLE t_plain AS SELECT * FROM t");
	const char *wheres[] = {"q BETWEEN 10 AND 19", "q = 24", "q < 4", "q >= 48 AND q <= 49", "q > 45", "q >= 7"};
	for (auto w : wheres) {
		const string where = w;
		const idx_t before = CubitRewriteCount();
		auto a = Run(con, "SELECT count(*), sum(price), sum(price * disc), min(rowid), max(rowid) FROM t WHERE " + where);
		REQUIRE(CubitRewriteCount() == before + 1); // the scan really went through cubit_scan
		auto b = Run(con, "SELECT count(*), sum(price), sum(price * disc), min(rowid), max(rowid) FROM t_plain WHERE " + where);
		REQUIRE(CubitRewriteCount() == before + 1);
		for (idx_t c = 0; c < 5; c++) {
			REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
		}
		auto x = Run(con, "SELECT rowid, price FROM t WHERE " + where + " ORDER BY rowid");
		auto y = Run(con, "SELECT rowid, price FROM t_plain WHERE " + where + " ORDER BY rowid");
		REQUIRE(x->RowCount() == y->RowCount());
		for (idx_t r = 0; r < x->RowCount(); r += 97) {
			REQUIRE(x->GetValue(0, r) == y->GetValue(0, r) && x->GetValue(1, r) == y->GetValue(1, r));
		}
	}
	{ // a filter on another column as well: not rewritten, still correct
		const idx_t before = CubitRewriteCount();
		auto a = Run(con, "SELECT count(*) FROM t WHERE q = 24 AND disc = 3");
		auto b = Run(con, "SELECT count(*) FROM t_plain WHERE q = 24 AND disc = 3");
		REQUIRE(CubitRewriteCount() == before && a->GetValue(0, 0) == b->GetValue(0, 0));
		auto plan = Run(con, "EXPLAIN SELECT price FROM t WHERE q BETWEEN 10 AND 19");
		REQUIRE(plan->GetValue(1, 0).ToString().find("CUBIT_SCAN") != string::npos);
		auto aplan = Run(con, "EXPLAIN SELECT sum(price) FROM t WHERE q BETWEEN 10 AND 19");
		REQUIRE(aplan->GetValue(1, 0).ToString().find("CUBIT_AGG_PUSHDOWN") != string::npos); // no aggregate operator left
		REQUIRE(aplan->GetValue(1, 0).ToString().find("AGGREGATE") == string::npos);
	}
	{ // aggregate push-down: ungrouped COUNT / SUM / SUM(a*b) directly on a rewritable scan → one row from the GPU
		const char *aggs[] = {"count(*)", "sum(price)", "count(*), sum(price), count(price), sum(q)", "sum(price * disc)",
		                      "sum(price) + 1, count(*) * 2", "sum(disc), sum(price * q), sum(price)"};
		for (auto &p : preds) {
			const string where = p[2];
			for (auto a : aggs) {
				const idx_t before = CubitAggPushdownCount();
				auto x = Run(con, string("SELECT ") + a + " FROM t WHERE " + where);
				auto y = Run(con, string("SELECT ") + a + " FROM t_plain WHERE " + where);
				auto any = Run(con, "SELECT count(*) FROM t_plain WHERE " + where)->GetValue(0, 0).GetValue<int64_t>() > 0;
				if (CubitAggPushdownCount() != before + (any ? 1 : 0)) {
					fprintf(stderr, "aggregate push-down expectation failed for: SELECT %s FROM t WHERE %s\n", a, where.c_str());
				}
				REQUIRE(CubitAggPushdownCount() == before + (any ? 1 : 0));
				REQUIRE(x->ColumnCount() == y->ColumnCount() && x->RowCount() == 1 && y->RowCount() == 1);
				for (idx_t c = 0; c < x->ColumnCount(); c++) {
					REQUIRE(x->types[c] == y->types[c]);
					REQUIRE(x->GetValue(c, 0).ToString() == y->GetValue(c, 0).ToString());
				}
			}
		}
		// not pushed: grouped, DISTINCT, FILTER, unsupported aggregates, expressions other than a product of columns
		const char *no[] = {"SELECT q, sum(price) FROM t WHERE q < 4 GROUP BY q ORDER BY q", "SELECT count(DISTINCT price) FROM t WHERE q = 24",
		                    "SELECT sum(price) FILTER (WHERE disc > 3) FROM t WHERE q = 24", "SELECT avg(price), min(price) FROM t WHERE q = 24",
		                    "SELECT sum(price + disc) FROM t WHERE q = 24", "SELECT sum(price * disc * 2) FROM t WHERE q = 24"};
		for (auto q : no) {
			const idx_t before = CubitAggPushdownCount(), rw = CubitRewriteCount();
			string sql = q, plain = q;
			plain.replace(plain.find("FROM t "), 7, "FROM t_plain ");
			auto x = Run(con, sql);
			auto y = Run(con, plain);
			REQUIRE(CubitAggPushdownCount() == before && CubitRewriteCount() == rw + 1); // the scan underneath still is
			REQUIRE(x->RowCount() == y->RowCount());
			for (idx_t r = 0; r < x->RowCount(); r++) {
				for (idx_t c = 0; c < x->ColumnCount(); c++) {
					REQUIRE(x->GetValue(c, r).ToString() == y->GetValue(c, r).ToString());
				}
			}
		}
		// the product overflows int64 on both sides: an error, not a wrong number (arithmetic.cpp:766-795)
		Run(con, "CREATE TABLE tov AS SELECT (i % 3)::BIGINT AS k, (4611686018427387904 + i)::BIGINT AS a, 4::BIGINT AS b FROM range(1000) r(i)");
		Run(con, "CALL cubit_load('tov', 'k', 0, 3)");
		REQUIRE(con.Query("SELECT sum(a * b) FROM tov WHERE k = 1")->HasError());
		printf("aggregate push-down ok\n");
	}
	{ // a second index on the same table: conjunctions across indexed columns become AND of OR groups
		auto l2 = Run(con, "CALL cubit_load('t', 'disc', 0, 11)");
		REQUIRE(l2->GetValue(0, 0).GetValue<int64_t>() == 300000);
		const char *conj[] = {"q BETWEEN 10 AND 19 AND disc BETWEEN 5 AND 7", "q = 24 AND disc = 3", "q < 24 AND disc >= 9",
		                      "disc = 0", "q >= 40 AND disc <= 1 AND disc >= 0", "q = 24 AND disc = 30"};
		for (auto w : conj) {
			const string where = w;
			const idx_t before = CubitRewriteCount();
			auto a = Run(con, "SELECT count(*), sum(price), sum(price * disc), min(rowid), max(rowid) FROM t WHERE " + where);
			auto b = Run(con, "SELECT count(*), sum(price), sum(price * disc), min(rowid), max(rowid) FROM t_plain WHERE " + where);
			REQUIRE(CubitRewriteCount() == before + (b->GetValue(0, 0).GetValue<int64_t>() > 0 ? 1 : 0));
			for (idx_t c = 0; c < 5; c++) {
				REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
			}
			auto x = Run(con, "SELECT rowid, q, price, disc FROM t WHERE " + where + " ORDER BY rowid");
			auto y = Run(con, "SELECT rowid, q, price, disc FROM t_plain WHERE " + where + " ORDER BY rowid");
			REQUIRE(x->RowCount() == y->RowCount());
			for (idx_t r = 0; r < x->RowCount(); r += 53) {
				for (idx_t c = 0; c < 4; c++) {
					REQUIRE(x->GetValue(c, r) == y->GetValue(c, r));
				}
			}
		}
		// a range that would read more than CUBIT_MAX_STREAMS bitvectors keeps the vanilla scan
		Run(con, "CREATE TABLE tw AS SELECT (i % 200)::BIGINT AS k, i::BIGINT AS v FROM range(50000) r(i)");
		Run(con, "CALL cubit_load('tw', 'k', 0, 200)");
		const idx_t before = CubitRewriteCount();
		auto wide = Run(con, "SELECT count(*), sum(v) FROM tw WHERE k BETWEEN 10 AND 120");
		REQUIRE(CubitRewriteCount() == before && wide->GetValue(0, 0).GetValue<int64_t>() == 111 * 250);
		auto narrow = Run(con, "SELECT count(*) FROM tw WHERE k BETWEEN 10 AND 70");
		REQUIRE(CubitRewriteCount() == before + 1 && narrow->GetValue(0, 0).GetValue<int64_t>() == 61 * 250);
		printf("multi-index conjunctions ok\n");
	}
	{ // binned indexes: DATE keys by month and DECIMAL keys by unit — TPC-H Q6 as written is answered by the GPU
		Run(con, "CREATE TABLE lq AS SELECT CAST((i * 7919 % 50 + 1) AS DECIMAL(15,2)) AS l_quantity, "
		         "CAST((i * 104729 % 100003) / 100.0 AS DECIMAL(15,2)) AS l_extendedprice, CAST((i % 11) / 100.0 AS DECIMAL(15,2)) AS l_discount, "
		         "DATE '1992-01-02' + CAST((i * 31 % 2526) AS INTEGER) AS l_shipdate FROM range(400000) r(i)");
		Run(con, "CREATE TABLE lq_plain AS SELECT * FROM lq");
		Run(con, "CALL cubit_load('lq', 'l_quantity', 100, 50, bin = '100')");   // raw cents 100 .. 5099 in 50 bins of 1.00
		Run(con, "CALL cubit_load('lq', 'l_discount', 0, 11)");                   // raw cents 0 .. 10, one bitvector per value
		Run(con, "CALL cubit_load('lq', 'l_shipdate', 0, 0, bin = 'month')");
		const string q6 = "SELECT sum(l_extendedprice * l_discount) AS revenue FROM %s WHERE l_shipdate >= CAST('1994-01-01' AS date) "
		                  "AND l_shipdate < CAST('1995-01-01' AS date) AND l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24";
		auto on = [](string q, const char *t) { q.replace(q.find("%s"), 2, t); return q; };
		idx_t before = CubitAggPushdownCount();
		auto a = Run(con, on(q6, "lq"));
		auto b = Run(con, on(q6, "lq_plain"));
		REQUIRE(CubitAggPushdownCount() == before + 1);
		REQUIRE(a->types[0] == b->types[0] && a->GetValue(0, 0).ToString() == b->GetValue(0, 0).ToString() && !a->GetValue(0, 0).IsNull());
		// rows, not aggregates, through the same three indexes
		before = CubitRewriteCount();
		auto x = Run(con, "SELECT rowid, l_extendedprice, l_shipdate FROM lq WHERE l_shipdate >= DATE '1993-03-01' AND l_shipdate < DATE '1993-06-01' AND l_quantity >= 10 AND l_quantity < 12 ORDER BY rowid");
		auto y = Run(con, "SELECT rowid, l_extendedprice, l_shipdate FROM lq_plain WHERE l_shipdate >= DATE '1993-03-01' AND l_shipdate < DATE '1993-06-01' AND l_quantity >= 10 AND l_quantity < 12 ORDER BY rowid");
		REQUIRE(CubitRewriteCount() == before); // l_shipdate is projected and is not an INT64 column: vanilla scan
		auto x2 = Run(con, "SELECT rowid, l_extendedprice FROM lq WHERE l_shipdate >= DATE '1993-03-01' AND l_shipdate < DATE '1993-06-01' AND l_quantity >= 10 AND l_quantity < 12 ORDER BY rowid");
		REQUIRE(CubitRewriteCount() == before + 1 && x2->RowCount() == y->RowCount() && x->RowCount() == y->RowCount() && y->RowCount() > 0);
		for (idx_t r = 0; r < y->RowCount(); r++) {
			REQUIRE(x2->GetValue(0, r) == y->GetValue(0, r) && x2->GetValue(1, r) == y->GetValue(1, r));
		}
		// a range that does not end on a bin boundary keeps the vanilla scan (and its answer)
		before = CubitRewriteCount();
		auto m1 = Run(con, "SELECT count(*), sum(l_extendedprice) FROM lq WHERE l_shipdate >= DATE '1994-01-15' AND l_shipdate < DATE '1995-01-01'");
		auto m2 = Run(con, "SELECT count(*), sum(l_extendedprice) FROM lq_plain WHERE l_shipdate >= DATE '1994-01-15' AND l_shipdate < DATE '1995-01-01'");
		auto m3 = Run(con, "SELECT count(*) FROM lq WHERE l_quantity < 24.5");
		auto m4 = Run(con, "SELECT count(*) FROM lq_plain WHERE l_quantity < 24.5");
		REQUIRE(CubitRewriteCount() == before && m1->GetValue(0, 0) == m2->GetValue(0, 0) && m1->GetValue(1, 0) == m2->GetValue(1, 0) && m3->GetValue(0, 0) == m4->GetValue(0, 0));
		// open-ended and out-of-domain ranges are aligned by construction
		const char *open_ended[] = {"l_shipdate >= DATE '1998-01-01'", "l_shipdate < DATE '1992-02-01'", "l_shipdate >= DATE '2001-01-01'",
		                            "l_quantity >= 49", "l_quantity <= 1", "l_shipdate < DATE '1990-01-01'"};
		for (auto w : open_ended) {
			auto c1 = Run(con, string("SELECT count(*), sum(l_extendedprice) FROM lq WHERE ") + w);
			auto c2 = Run(con, string("SELECT count(*), sum(l_extendedprice) FROM lq_plain WHERE ") + w);
			REQUIRE(c1->GetValue(0, 0) == c2->GetValue(0, 0) && c1->GetValue(1, 0).ToString() == c2->GetValue(1, 0).ToString());
		}
		// keys outside the declared domain are refused at load time (they would silently drop out of scans)
		Run(con, "CREATE TABLE dom AS SELECT (i % 60)::BIGINT AS k FROM range(1000) r(i)");
		REQUIRE(con.Query("CALL cubit_load('dom', 'k', 0, 50)")->HasError());
		printf("binned indexes ok\n");
	}
	{ // NULLs: in projected columns (validity masks on the DataChunk vectors), in aggregate inputs (skipped;
	  // SUM over only-NULL inputs is NULL) and in the key (NULL keys are not indexed)
		Run(con, "CREATE TABLE tn AS SELECT CASE WHEN i % 13 = 0 THEN NULL ELSE (i * 7919 % 50 + 1) END::BIGINT AS q, "
		         "CASE WHEN i % 7 = 3 OR (i >= 4096 AND i < 8300) THEN NULL ELSE (i * 104729 % 1000003 - 500000) END::BIGINT AS price, "
		         "CASE WHEN i % 11 = 5 THEN NULL ELSE i % 11 END::BIGINT AS disc FROM range(200000) r(i)");
		Run(con, "CREATE TABLE tn_plain AS SELECT * FROM tn");
		Run(con, "CALL cubit_load('tn', 'q', 1, 50)");
		for (auto &p : preds) {
			const string lo = p[0], hi = p[1], where = p[2];
			auto a = Run(con, "SELECT count(*), count(price), sum(price), count(price * disc), sum(price * disc) FROM cubit_scan('tn', " + lo + ", " + hi + ")");
			auto b = Run(con, "SELECT count(*), count(price), sum(price), count(price * disc), sum(price * disc) FROM tn_plain WHERE " + where);
			for (idx_t c = 0; c < 5; c++) {
				REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
			}
			auto x = Run(con, "SELECT q, price, disc FROM cubit_scan('tn', " + lo + ", " + hi + ")");
			auto y = Run(con, "SELECT q, price, disc FROM tn_plain WHERE " + where + " ORDER BY rowid");
			REQUIRE(x->RowCount() == y->RowCount());
			idx_t nulls_seen = 0;
			for (idx_t r = 0; r < x->RowCount(); r++) {
				for (idx_t c = 0; c < 3; c++) {
					REQUIRE(x->GetValue(c, r).ToString() == y->GetValue(c, r).ToString());
				}
				nulls_seen += x->GetValue(1, r).IsNull();
			}
			REQUIRE(x->RowCount() == 0 || nulls_seen > 0);
			const idx_t before = CubitRewriteCount();
			auto ra = Run(con, "SELECT count(*), count(price), sum(price), sum(price * disc) FROM tn WHERE " + where);
			// (statistics let the optimizer drop a scan whose predicate cannot match: no rewrite then)
			REQUIRE(CubitRewriteCount() == before + (b->GetValue(0, 0).GetValue<int64_t>() > 0 ? 1 : 0));
			for (idx_t c = 0; c < 4; c++) {
				REQUIRE(ra->GetValue(c, 0).ToString() == b->GetValue(c == 3 ? 4 : c, 0).ToString());
			}
			auto g = Run(con, "SELECT * FROM cubit_agg('tn', " + lo + ", " + hi + ", 'price')");
			REQUIRE(g->GetValue(0, 0).ToString() == b->GetValue(0, 0).ToString());
			REQUIRE(g->GetValue(1, 0).ToString() == b->GetValue(2, 0).ToString()); // NULL when no non-NULL input
		}
		// a selection whose aggregate inputs are ALL NULL → SUM is NULL, COUNT(*) is not 0
		Run(con, "CREATE TABLE tz AS SELECT (i % 5 + 1)::BIGINT AS q, CASE WHEN i % 5 = 2 THEN NULL ELSE i END::BIGINT AS price FROM range(5000) r(i)");
		Run(con, "CALL cubit_load('tz', 'q', 1, 5)");
		auto z = Run(con, "SELECT * FROM cubit_agg('tz', 3, 3, 'price')");
		REQUIRE(z->GetValue(0, 0).GetValue<int64_t>() == 1000 && z->GetValue(1, 0).IsNull());
		auto z2 = Run(con, "SELECT count(*), sum(price) FROM tz WHERE q = 3");
		REQUIRE(z2->GetValue(0, 0).GetValue<int64_t>() == 1000 && z2->GetValue(1, 0).IsNull());
		printf("null semantics ok\n");
	}
	{ // DML reaches the GPU copy through the table's CUBIT index (BoundIndex::Append / Delete, bound_index.hpp:71-97):
	  // DELETE and UPDATE become pending deltas, INSERT appends, and the SAME SELECTs keep being answered by the GPU
	  // with the answers of the vanilla scan (SURVEY §8c "delta semantics oracle": rowids are stable, updated rows of
	  // an indexed column move to the end, new rows take the next row ids)
		const vector<string> dml_wheres = {"q BETWEEN 10 AND 19", "q = 24", "q < 4", "q >= 48 AND q <= 49", "q = 24 AND disc = 3",
		                                   "q > 40 AND disc BETWEEN 2 AND 5", "disc = 4"};
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		const idx_t pairs0 = CubitDmlDeltaPairs(), app0 = CubitDmlAppendedRows();
		auto del = Run(con, "DELETE FROM t WHERE q = 24 AND price % 3 = 0");
		REQUIRE(del->GetValue(0, 0).GetValue<int64_t>() > 0);
		REQUIRE(CubitDmlDeltaPairs() >= pairs0 + 2 * NumericCast<idx_t>(del->GetValue(0, 0).GetValue<int64_t>())); // two indexes
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		auto plan = Run(con, "EXPLAIN SELECT price FROM t WHERE q BETWEEN 10 AND 19");
		REQUIRE(plan->GetValue(1, 0).ToString().find("CUBIT_SCAN") != string::npos); // still the GPU scan
		// UPDATE of the key column and of a payload column: both are index columns → DELETE + INSERT
		auto up1 = Run(con, "UPDATE t SET q = q % 50 + 1 WHERE disc = 4 AND price > 0");
		auto up2 = Run(con, "UPDATE t SET price = price + 7 WHERE q = 11");
		REQUIRE(up1->GetValue(0, 0).GetValue<int64_t>() > 0 && up2->GetValue(0, 0).GetValue<int64_t>() > 0);
		REQUIRE(CubitDmlAppendedRows() == app0 + NumericCast<idx_t>(up1->GetValue(0, 0).GetValue<int64_t>() + up2->GetValue(0, 0).GetValue<int64_t>()));
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		// INSERT: past several segment boundaries of the bitvectors
		Run(con, "INSERT INTO t SELECT (i * 31 % 50 + 1)::BIGINT, (i * 17 - 40000)::BIGINT, (i % 11)::BIGINT FROM range(150000) r(i)");
		REQUIRE(CubitDmlAppendedRows() >= app0 + 150000);
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		// inside a transaction with uncommitted changes only the vanilla scan can see them: no rewrite, right answers
		Run(con, "BEGIN");
		Run(con, "DELETE FROM t WHERE q = 3");
		const idx_t rw = CubitRewriteCount();
		auto in_txn = Run(con, "SELECT count(*) FROM t WHERE q < 4");
		REQUIRE(CubitRewriteCount() == rw);
		Run(con, "ROLLBACK");
		auto after = Run(con, "SELECT count(*) FROM t WHERE q < 4");
		REQUIRE(CubitRewriteCount() == rw + 1);
		REQUIRE(after->GetValue(0, 0).GetValue<int64_t>() > in_txn->GetValue(0, 0).GetValue<int64_t>()); // the delete was rolled back
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		Run(con, "BEGIN");
		Run(con, "INSERT INTO t VALUES (24, 123456, 3), (25, -5, 4)");
		Run(con, "DELETE FROM t WHERE q = 49 AND disc = 0");
		Run(con, "COMMIT");
		CompareWithVanilla(con, "t", "q, price, disc", dml_wheres);
		// binned indexes get appended rows as pending deltas (their bin-id source column is temporary)
		const vector<string> lq_wheres = {"l_quantity < 24 AND l_discount BETWEEN 0.05 AND 0.07", "l_discount = 0.03",
		                                  "l_shipdate >= DATE '1994-01-01' AND l_shipdate < DATE '1995-01-01' AND l_quantity >= 10 AND l_quantity < 12"};
		Run(con, "INSERT INTO lq SELECT CAST((i % 50 + 1) AS DECIMAL(15,2)), CAST(i / 7.0 AS DECIMAL(15,2)), CAST((i % 11) / 100.0 AS DECIMAL(15,2)), "
		         "DATE '1993-06-15' + CAST((i * 7 % 700) AS INTEGER) FROM range(90000) r(i)");
		Run(con, "DELETE FROM lq WHERE l_quantity = 11 AND l_discount = 0.06");
		for (auto &w : lq_wheres) {
			const string sql = "SELECT count(*), sum(l_extendedprice), sum(l_extendedprice * l_discount), min(rowid), max(rowid) FROM lq WHERE " + w;
			const idx_t before = CubitRewriteCount();
			auto a = Run(con, sql);
			auto b = RunVanilla(con, sql);
			REQUIRE(CubitRewriteCount() == before + 1);
			for (idx_t c = 0; c < 5; c++) {
				REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
			}
		}
		// a key outside the indexed domain cannot be represented: the GPU copy steps aside (vanilla answers), reload fixes it
		Run(con, "INSERT INTO t VALUES (77, 1, 1)");
		const idx_t rw2 = CubitRewriteCount();
		auto gone = Run(con, "SELECT count(*) FROM t WHERE q = 24");
		REQUIRE(CubitRewriteCount() == rw2);
		Run(con, "DELETE FROM t WHERE q = 77");
		Run(con, "CALL cubit_load('t', 'q', 1, 50)");
		auto back = Run(con, "SELECT count(*) FROM t WHERE q = 24");
		REQUIRE(CubitRewriteCount() == rw2 + 1 && back->GetValue(0, 0) == gone->GetValue(0, 0));
		CompareWithVanilla(con, "t", "q, price, disc", {"q BETWEEN 10 AND 19", "q = 24"}); // (positions of deleted rows hold no key)
		// DROP TABLE + CREATE of the same name: a different table, no GPU copy, vanilla answers (no name-keyed registry)
		Run(con, "CREATE TABLE dropme AS SELECT (i % 5)::BIGINT AS k, i::BIGINT AS v FROM range(10000) r(i)");
		Run(con, "CALL cubit_load('dropme', 'k', 0, 5)");
		Run(con, "DROP TABLE dropme");
		Run(con, "CREATE TABLE dropme AS SELECT (i % 5)::BIGINT AS k, (i * 2)::BIGINT AS v FROM range(20000) r(i)");
		const idx_t rw3 = CubitRewriteCount();
		auto fresh = Run(con, "SELECT count(*), sum(v) FROM dropme WHERE k = 2");
		REQUIRE(CubitRewriteCount() == rw3 && fresh->GetValue(0, 0).GetValue<int64_t>() == 4000);
		REQUIRE(con.Query("SELECT * FROM cubit_scan('dropme', 1, 2)")->HasError());
		// DECIMAL keys narrower than 64 bits keep their unscaled value (a numeric cast would round 0.05 to 0)
		Run(con, "CREATE TABLE dn AS SELECT CAST((i % 11) / 100.0 AS DECIMAL(4,2)) AS d, i::BIGINT AS v FROM range(50000) r(i)");
		Run(con, "CALL cubit_load('dn', 'd', 0, 11)");
		const idx_t rw4 = CubitRewriteCount();
		auto dn = Run(con, "SELECT count(*), sum(v) FROM dn WHERE d BETWEEN 0.05 AND 0.07");
		auto dv = RunVanilla(con, "SELECT count(*), sum(v) FROM dn WHERE d BETWEEN 0.05 AND 0.07");
		REQUIRE(CubitRewriteCount() == rw4 + 1 && dn->GetValue(0, 0) == dv->GetValue(0, 0) && dn->GetValue(1, 0).ToString() == dv->GetValue(1, 0).ToString());
		REQUIRE(dn->GetValue(0, 0).GetValue<int64_t>() > 10000);
		printf("dml through the index ok\n");
	}
	// CUBIT_GPU_DEVICES with more than one device: every table is cut over several shards behind one handle.  All
	// answers must stay the same; only the per-shard features (compressed-segment upload, index images) step aside.
	bool sharded = false;
	if (const char *devs = getenv("CUBIT_GPU_DEVICES")) {
		int have = 0;
		cubit_gpu_device_count(&have);
		sharded = string(devs).find(',') != string::npos || (string(devs) == "all" ? have : std::min(atoi(devs), have)) > 1;
	}
	if (argc > 2 && string(argv[1]) == "--db") {
		// Storage route: a FILE-backed, checkpointed table's columns reach the C-ABI as the compressed segments
		// the reference wrote (BitPacking), lifted from the buffer manager — not as decoded rows.
		DuckDB fdb(argv[2]);
		Connection fcon(fdb);
		RegisterCubitGpuFunctions(*fdb.instance);
		Run(fcon, "CREATE TABLE ft AS SELECT (i * 7919 % 50 + 1)::BIGINT AS q, (i * 104729 % 1000003 - 500000)::BIGINT AS price, "
		          "(i * 3)::BIGINT AS seq, CAST((i % 11) AS DECIMAL(15,2)) AS disc FROM range(400000) r(i)");
		Run(fcon, "CHECKPOINT");
		const idx_t seg_before = CubitSegmentRouteCount();
		auto fl = Run(fcon, "CALL cubit_load('ft', 'q', 1, 50)");
		REQUIRE(fl->GetValue(0, 0).GetValue<int64_t>() == 400000);
		REQUIRE(CubitSegmentRouteCount() == seg_before + (sharded ? 0 : 4)); // all four columns went through the segment route
		Run(fcon, "CREATE TABLE ft_plain AS SELECT * FROM ft");
		for (auto w : wheres) {
			const string where = w;
			auto a = Run(fcon, "SELECT count(*), sum(price), sum(seq), min(rowid), max(rowid) FROM ft WHERE " + where);
			auto b = Run(fcon, "SELECT count(*), sum(price), sum(seq), min(rowid), max(rowid) FROM ft_plain WHERE " + where);
			for (idx_t c = 0; c < 5; c++) {
				REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
			}
			auto x = Run(fcon, "SELECT q, price, seq, disc FROM cubit_scan('ft', 10, 19)");
			auto y = Run(fcon, "SELECT q, price, seq, CAST(disc * 100 AS BIGINT) FROM ft_plain WHERE q BETWEEN 10 AND 19 ORDER BY rowid");
			REQUIRE(x->RowCount() == y->RowCount());
			for (idx_t r = 0; r < x->RowCount(); r += 101) {
				for (idx_t c = 0; c < 4; c++) {
					REQUIRE(x->GetValue(c, r).ToString() == y->GetValue(c, r).ToString());
				}
			}
		}
		// an UPDATE of a resident column reaches the index as DELETE + INSERT: the GPU copy follows it; the rows it
		// deleted stay in the table's segments, so the next load cannot use the segment route (physical position =
		// row id only holds for the decoded, rowid-placed rows) and falls back for every column
		Run(fcon, "SET wal_autocheckpoint='10GB'"); // keep the UPDATE un-checkpointed
		auto upd = Run(fcon, "UPDATE ft SET price = price + 1 WHERE q = 7");
		REQUIRE(upd->GetValue(0, 0).GetValue<int64_t>() == 8000);
		Run(fcon, "UPDATE ft_plain SET price = price + 1 WHERE q = 7");
		const idx_t rw_before = CubitRewriteCount();
		auto v1 = Run(fcon, "SELECT sum(price), count(*) FROM ft WHERE q = 7");
		REQUIRE(CubitRewriteCount() == rw_before + 1);
		auto u2 = Run(fcon, "SELECT sum(price), count(*) FROM ft_plain WHERE q = 7");
		REQUIRE(v1->GetValue(0, 0).ToString() == u2->GetValue(0, 0).ToString() && v1->GetValue(1, 0) == u2->GetValue(1, 0));
		const idx_t seg_mid = CubitSegmentRouteCount();
		Run(fcon, "CALL cubit_load('ft', 'q', 1, 50)");
		REQUIRE(CubitSegmentRouteCount() == seg_mid);
		auto u1 = Run(fcon, "SELECT sum(price) FROM cubit_scan('ft', 7, 7)");
		REQUIRE(u1->GetValue(0, 0).ToString() == v1->GetValue(0, 0).ToString());
		{ // RLE-compressed columns take the same route (CUBIT_SEG_RLE)
			Run(fcon, "PRAGMA force_compression='rle'");
			Run(fcon, "CREATE TABLE fr AS SELECT ((i // 3) * 7919 % 50 + 1)::BIGINT AS q, ((i // 40) * 104729 % 1000003 - 500000)::BIGINT AS price, "
			          "(i // 70000)::BIGINT AS big FROM range(300000) r(i)");
			Run(fcon, "CHECKPOINT");
			Run(fcon, "PRAGMA force_compression='auto'");
			auto comp = Run(fcon, "SELECT count(*) FROM pragma_storage_info('fr') WHERE segment_type <> 'VALIDITY' AND compression = 'RLE'");
			REQUIRE(comp->GetValue(0, 0).GetValue<int64_t>() >= 3);
			const idx_t seg_rle = CubitSegmentRouteCount();
			Run(fcon, "CALL cubit_load('fr', 'q', 1, 50)");
			REQUIRE(CubitSegmentRouteCount() == seg_rle + (sharded ? 0 : 3));
			Run(fcon, "CREATE TABLE fr_plain AS SELECT * FROM fr");
			for (auto w : wheres) {
				const string where = w;
				auto a = Run(fcon, "SELECT count(*), sum(price), sum(big), min(rowid), max(rowid) FROM fr WHERE " + where);
				auto b = Run(fcon, "SELECT count(*), sum(price), sum(big), min(rowid), max(rowid) FROM fr_plain WHERE " + where);
				for (idx_t c = 0; c < 5; c++) {
					REQUIRE(a->GetValue(c, 0).ToString() == b->GetValue(c, 0).ToString());
				}
			}
		}
		printf("storage route ok\n");
		{ // the CUBIT index is a catalog object of a registered index type: it is checkpointed with the table (its
		  // images written to blocks by GetStorageInfo), bound again through create_instance when the file is
		  // re-opened, and the first statement that can use it materialises the GPU copy from the images
			Run(fcon, "CREATE TABLE pt AS SELECT (i * 7919 % 50 + 1)::BIGINT AS q, (i * 104729 % 1000003 - 500000)::BIGINT AS price, "
			          "(i % 11)::BIGINT AS disc FROM range(250000) r(i)");
			Run(fcon, "CALL cubit_load('pt', 'q', 1, 50)");
			Run(fcon, "CALL cubit_load('pt', 'disc', 0, 11)");
			Run(fcon, "DELETE FROM pt WHERE q = 24 AND disc = 3"); // pending deltas travel in the image
			auto want = Run(fcon, "SELECT count(*), sum(price), sum(price * disc) FROM pt WHERE q BETWEEN 20 AND 29 AND disc < 5");
			auto idx = Run(fcon, "SELECT index_name FROM duckdb_indexes() WHERE table_name = 'pt'");
			REQUIRE(idx->RowCount() == 1 && idx->GetValue(0, 0).ToString() == "cubit_pt");
			Run(fcon, "CHECKPOINT");
		}
	}
	if (argc > 2 && string(argv[1]) == "--db") {
		DuckDB rdb(argv[2]); // (the first instance went out of scope: the file is closed and re-opened)
		Connection rcon(rdb);
		RegisterCubitGpuFunctions(*rdb.instance);
		const idx_t img0 = CubitImageLoads(), rw0 = CubitRewriteCount();
		auto got = Run(rcon, "SELECT count(*), sum(price), sum(price * disc) FROM pt WHERE q BETWEEN 20 AND 29 AND disc < 5");
		REQUIRE(CubitRewriteCount() == rw0 + 1);
		REQUIRE(CubitImageLoads() == img0 + (sharded ? 0 : 2)); // both indexes came from their checkpointed images (sharded: rebuilt from the table)
		auto van = RunVanilla(rcon, "SELECT count(*), sum(price), sum(price * disc) FROM pt WHERE q BETWEEN 20 AND 29 AND disc < 5");
		for (idx_t c = 0; c < 3; c++) {
			REQUIRE(got->GetValue(c, 0).ToString() == van->GetValue(c, 0).ToString());
		}
		auto d24 = Run(rcon, "SELECT count(*) FROM pt WHERE q = 24 AND disc = 3");
		REQUIRE(d24->GetValue(0, 0).GetValue<int64_t>() == 0);
		// DML keeps working on the re-attached index
		Run(rcon, "INSERT INTO pt VALUES (24, 5, 3), (24, 6, 3)");
		auto d2 = Run(rcon, "SELECT count(*), sum(price) FROM pt WHERE q = 24 AND disc = 3");
		if (d2->GetValue(0, 0).GetValue<int64_t>() != 2) {
			fprintf(stderr, "after re-attach + INSERT: count %s sum %s\n", d2->GetValue(0, 0).ToString().c_str(), d2->GetValue(1, 0).ToString().c_str());
		}
		REQUIRE(d2->GetValue(0, 0).GetValue<int64_t>() == 2 && d2->GetValue(1, 0).GetValue<int64_t>() == 11);
		Run(rcon, "DROP TABLE pt"); // CommitDrop releases the GPU memory and the image blocks
		Run(rcon, "CHECKPOINT");
		printf("index persistence ok\n");
	}
	auto err = con.Query("SELECT * FROM cubit_scan('nope', 1, 2)");
	REQUIRE(err->HasError());
	printf("duckdb_sql_test ok%s\n", sharded ? " (tables sharded behind one handle)" : "");
	return 0;
}
