// duckdb_cubit_extension.cpp — the DuckDB-side binding of the B200 CUBIT scan (reference glue).
//
// This file is compiled against the REAL reference headers (/root/reference/src/include) — it is the
// code a DuckDB maintainer adds to make the GPU path a drop-in behind PhysicalTableScan::GetData.
// It only talks to the GPU library through include/cubit_gpu.h (extern "C").  It is syntax-checked
// against the reference headers by __graft_entry__.build() whenever /root/reference is present and is
// reproduced, annotated, in INTEGRATION.md.  Nothing here is copied from the reference; it USES its
// public extension API:
//   ExtensionUtil::RegisterFunction          src/include/duckdb/main/extension_util.hpp:34
//   TableFunction{bind, init_global, function} src/include/duckdb/function/table_function.hpp:184-301
//   DataChunk / FlatVector                   src/include/duckdb/common/types/{data_chunk,vector}.hpp
//
// SQL surface:
//   CALL cubit_load('lineitem', 'l_quantity', 1, 50);      -- build the GPU index + upload BIGINT columns
//   SELECT sum(l_extendedprice) FROM cubit_scan('lineitem', 24, 24);   -- rowid + the uploaded columns
//   SELECT * FROM cubit_agg('lineitem', 24, 24, 'l_extendedprice');    -- aggregate push-down: one row
// and, transparently (OptimizerExtension, src/include/duckdb/optimizer/optimizer_extension.hpp:31-43): a plain
//   SELECT sum(l_extendedprice) FROM lineitem WHERE l_quantity BETWEEN 10 AND 19;
// whose pushed-down filters touch only the indexed column is re-pointed at the GPU scan, the way
// TableScanPushdownComplexFilter re-points a seq_scan at ART's index_scan (src/function/table/table_scan.cpp:296-370).
#include "duckdb.hpp"
#include "duckdb/function/table_function.hpp"
#include "duckdb/main/extension_util.hpp"
#include "duckdb/common/types/data_chunk.hpp"
#include "duckdb/common/types/date.hpp"
#include "duckdb/common/types/vector.hpp"
#include "duckdb/catalog/catalog_entry/table_catalog_entry.hpp"
#include "duckdb/main/config.hpp"
#include "duckdb/optimizer/optimizer_extension.hpp"
#include "duckdb/planner/filter/conjunction_filter.hpp"
#include "duckdb/planner/filter/constant_filter.hpp"
#include "duckdb/planner/expression/bound_aggregate_expression.hpp"
#include "duckdb/planner/expression/bound_cast_expression.hpp"
#include "duckdb/planner/expression/bound_columnref_expression.hpp"
#include "duckdb/planner/expression/bound_function_expression.hpp"
#include "duckdb/planner/operator/logical_aggregate.hpp"
#include "duckdb/planner/operator/logical_delete.hpp"
#include "duckdb/planner/operator/logical_get.hpp"
#include "duckdb/planner/operator/logical_insert.hpp"
#include "duckdb/planner/operator/logical_update.hpp"
#include "duckdb/storage/block_manager.hpp"
#include "duckdb/storage/buffer_manager.hpp"
#include "duckdb/storage/data_table.hpp"
#include "duckdb/storage/table_io_manager.hpp"
#include "duckdb/storage/table_storage_info.hpp"

#include "cubit_gpu.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace duckdb {

// ---------------------------------------------------------------- registry of GPU-resident tables
// one CUBIT index of a GPU-resident table: value v of `key_table_column` ↔ bitvector v - base_value
struct CubitGpuIndex {
	int32_t index_id = -1;
	int64_t base_value = 0;
	uint32_t cardinality = 0;
	idx_t key_table_column = 0; // table column index of the indexed column
	// binned index (cubit_load(..., bin := 'month' | '<width>')): bitvector b covers the key values in
	// [bin_lo[b], bin_lo[b + 1]) (cardinality + 1 boundaries); a range predicate is answered by it only when both
	// of its ends fall on boundaries.  Empty = one bitvector per value.
	vector<int64_t> bin_lo;
};

struct CubitGpuTable {
	cubit_gpu_table *handle = nullptr;
	vector<CubitGpuIndex> indexes; // [0] = the one cubit_scan(table, lo, hi) / cubit_agg address; CALL cubit_load again
	                               // with another key column to add more (conjunctions across them are rewritten)
	vector<string> column_names; // uploaded BIGINT columns, column id = position
	vector<idx_t> table_column;  // table column index of every uploaded column
	idx_t row_count = 0;
	// page-locked staging buffers (one DataChunk window each) recycled across scans: allocating and freeing
	// pinned memory costs milliseconds, far more than a scan
	std::mutex pool_lock;
	vector<void *> host_pool;
	void *AcquireWindow(uint64_t bytes) {
		{
			std::lock_guard<std::mutex> lk(pool_lock);
			if (!host_pool.empty()) {
				void *p = host_pool.back();
				host_pool.pop_back();
				return p;
			}
		}
		void *p = nullptr;
		if (cubit_gpu_alloc_host(bytes, &p) != CUBIT_OK) {
			throw InvalidInputException("cubit_gpu: %s", cubit_gpu_last_error());
		}
		return p;
	}
	void ReleaseWindow(void *p) {
		if (p) {
			std::lock_guard<std::mutex> lk(pool_lock);
			host_pool.push_back(p);
		}
	}
	~CubitGpuTable() {
		for (auto p : host_pool) {
			cubit_gpu_free_host(p);
		}
		cubit_gpu_destroy(handle);
	}
};

static std::mutex cubit_registry_lock;
static std::unordered_map<string, shared_ptr<CubitGpuTable>> cubit_registry;

static void CubitCheck(int rc) {
	if (rc != CUBIT_OK) {
		// errors cross the C-ABI as codes; here they become the exceptions DuckDB expects
		throw InvalidInputException("cubit_gpu: %s", cubit_gpu_last_error());
	}
}

static shared_ptr<CubitGpuTable> CubitLookup(const string &name) {
	std::lock_guard<std::mutex> lk(cubit_registry_lock);
	auto it = cubit_registry.find(name);
	if (it == cubit_registry.end()) {
		throw InvalidInputException("no CUBIT GPU index loaded for table \"%s\" (CALL cubit_load first)", name);
	}
	return it->second;
}

// ---------------------------------------------------------------- cubit_load(table, key, base, cardinality)
struct CubitLoadBindData : public TableFunctionData {
	string table, key;
	int64_t base = 0;
	uint32_t cardinality = 0;
	string bin; // "" = one bitvector per value; "month" (DATE keys); or a positive integer bin width
	bool done = false;
};

static unique_ptr<FunctionData> CubitLoadBind(ClientContext &, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitLoadBindData>();
	bind->table = input.inputs[0].GetValue<string>();
	bind->key = input.inputs[1].GetValue<string>();
	bind->base = input.inputs[2].GetValue<int64_t>();
	bind->cardinality = NumericCast<uint32_t>(input.inputs[3].GetValue<int64_t>());
	auto bin = input.named_parameters.find("bin");
	if (bin != input.named_parameters.end()) {
		bind->bin = bin->second.GetValue<string>();
	}
	return_types.emplace_back(LogicalType::BIGINT);
	names.emplace_back("rows_indexed");
	return std::move(bind);
}

// ---- storage route: hand a column's ON-DISK segments to the GPU as they are.
// GetColumnSegmentInfo is what pragma_storage_info shows (src/function/table/system/pragma_storage_info.cpp):
// per segment its first row, row count, compression, and where its bytes live (block_id, block_offset).  The
// block is pinned through the buffer manager exactly as BitpackingScanState does (bitpacking.cpp:627-636) and
// the bytes go to cubit_gpu_upload_column_segments untouched; the GPU decodes them.  Returns false (and the
// caller falls back to pulling decoded rows through a query) unless EVERY segment of the column is a
// persistent, un-updated BitPacking, RLE or Uncompressed INT64 segment without NULLs.
static std::atomic<idx_t> cubit_segment_columns {0};
idx_t CubitSegmentRouteCount() {
	return cubit_segment_columns.load();
}

static bool CubitUploadColumnSegments(ClientContext &context, const string &table, idx_t table_column,
                                      const LogicalType &type, cubit_gpu_table *handle, int32_t gpu_col, idx_t row_count) {
	if (type.InternalType() != PhysicalType::INT64) {
		return false;
	}
	auto &entry = Catalog::GetEntry<TableCatalogEntry>(context, INVALID_CATALOG, DEFAULT_SCHEMA, table);
	if (!entry.IsDuckTable()) {
		return false;
	}
	auto &storage = entry.GetStorage();
	auto &block_manager = TableIOManager::Get(storage).GetBlockManagerForRowData();
	if (block_manager.InMemory()) {
		return false;
	}
	vector<ColumnSegmentInfo> mine;
	for (auto &info : entry.GetColumnSegmentInfo()) {
		if (info.column_id != table_column) {
			continue;
		}
		if (info.segment_type == "VALIDITY") {
			if (info.segment_stats.find("Has Null: false") == string::npos) {
				return false; // NULL keys/values are not handled on this route
			}
			continue;
		}
		if (!info.persistent || info.has_updates || info.block_id < 0 ||
		    (info.compression_type != "BitPacking" && info.compression_type != "Uncompressed" &&
		     info.compression_type != "RLE")) {
			return false;
		}
		mine.push_back(info);
	}
	std::sort(mine.begin(), mine.end(),
	          [](const ColumnSegmentInfo &a, const ColumnSegmentInfo &b) { return a.segment_start < b.segment_start; });
	vector<BufferHandle> pins; // keep every block resident until the upload returns
	vector<cubit_column_segment> segs;
	idx_t next_row = 0;
	for (auto &info : mine) {
		if (info.segment_start != next_row) {
			return false;
		}
		next_row += info.segment_count;
		auto block = block_manager.RegisterBlock(info.block_id);
		pins.push_back(block_manager.buffer_manager.Pin(block));
		auto ptr = pins.back().Ptr() + info.block_offset;
		cubit_column_segment seg;
		seg.reserved = 0;
		seg.row_start = info.segment_start;
		seg.count = info.segment_count;
		seg.data = ptr;
		if (info.compression_type == "BitPacking") {
			seg.kind = CUBIT_SEG_BITPACKING;
			seg.bytes = Load<idx_t>(ptr); // offset of the end of the metadata = segment size (bitpacking.cpp:524-544)
			if (info.block_offset + seg.bytes > block_manager.GetBlockSize()) {
				return false;
			}
		} else if (info.compression_type == "RLE") {
			seg.kind = CUBIT_SEG_RLE;
			// an RLE segment does not store its size (rle.cpp:190-205): hand over everything up to the end of
			// the block, the library validates the runs it needs against that bound
			seg.bytes = block_manager.GetBlockSize() - info.block_offset;
		} else {
			seg.kind = CUBIT_SEG_UNCOMPRESSED;
			seg.bytes = info.segment_count * sizeof(int64_t);
		}
		segs.push_back(seg);
	}
	if (next_row != row_count || segs.empty()) {
		return false;
	}
	cubit_decode_info dinfo;
	CubitCheck(cubit_gpu_upload_column_segments(handle, gpu_col, 8, segs.data(), NumericCast<uint32_t>(segs.size()), &dinfo));
	cubit_segment_columns++;
	return true;
}

// Build one index over GPU column `gcol` of a resident table.  The key values come back from the GPU (the
// uploaded raw int64, NULL rows = INT64_MIN): every non-NULL key must fall into the indexed domain — a key the
// index does not cover would silently drop rows from rewritten scans — and binned indexes need the bin id of
// every row (computed here, uploaded as a temporary column, indexed on the GPU, dropped).
static CubitGpuIndex CubitBuildIndex(CubitGpuTable &gpu, idx_t gcol, const CubitLoadBindData &bind, const LogicalType &key_type) {
	CubitGpuIndex ix;
	ix.key_table_column = gpu.table_column[gcol];
	vector<int64_t> vals(gpu.row_count);
	if (gpu.row_count) {
		CubitCheck(cubit_gpu_download_column(gpu.handle, NumericCast<int32_t>(gcol), vals.data(), 8, gpu.row_count));
	}
	const int64_t null_key = NumericLimits<int64_t>::Minimum();
	if (bind.bin.empty()) {
		ix.base_value = bind.base;
		ix.cardinality = bind.cardinality;
		for (auto v : vals) {
			if (v != null_key && (v < ix.base_value || v >= ix.base_value + ix.cardinality)) {
				throw InvalidInputException("cubit_load: key %lld of \"%s\" lies outside the indexed domain [%lld, %lld)",
				                            (long long)v, bind.key, (long long)ix.base_value,
				                            (long long)(ix.base_value + ix.cardinality));
			}
		}
		CubitCheck(cubit_gpu_index_create(gpu.handle, ix.cardinality, &ix.index_id));
		CubitCheck(cubit_gpu_index_build(gpu.handle, ix.index_id, NumericCast<int32_t>(gcol), ix.base_value));
		return ix;
	}
	// ---- binned
	vector<int64_t> bins(vals.size(), null_key);
	if (bind.bin == "month") {
		if (key_type.id() != LogicalTypeId::DATE) {
			throw InvalidInputException("cubit_load: bin := 'month' needs a DATE key column");
		}
		int64_t lo = NumericLimits<int64_t>::Maximum(), hi = NumericLimits<int64_t>::Minimum();
		for (auto v : vals) {
			if (v != null_key) {
				lo = MinValue(lo, v);
				hi = MaxValue(hi, v);
			}
		}
		if (lo > hi) {
			lo = hi = 0;
		}
		int32_t y0, m0, d0, y1, m1, d1;
		Date::Convert(date_t(NumericCast<int32_t>(lo)), y0, m0, d0);
		Date::Convert(date_t(NumericCast<int32_t>(hi)), y1, m1, d1);
		const int32_t n_bins = (y1 - y0) * 12 + (m1 - m0) + 1;
		for (int32_t b = 0; b <= n_bins; b++) {
			const int32_t mm = m0 - 1 + b;
			ix.bin_lo.push_back(Date::FromDate(y0 + mm / 12, mm % 12 + 1, 1).days);
		}
		for (idx_t r = 0; r < vals.size(); r++) {
			if (vals[r] != null_key) {
				int32_t y, m, d;
				Date::Convert(date_t(NumericCast<int32_t>(vals[r])), y, m, d);
				bins[r] = (y - y0) * 12 + (m - m0);
			}
		}
	} else {
		int64_t width = 0;
		try {
			width = std::stoll(bind.bin);
		} catch (...) {
			width = 0;
		}
		if (width <= 0 || bind.cardinality == 0) {
			throw InvalidInputException("cubit_load: bin must be 'month' or a positive integer width (with base and cardinality)");
		}
		for (uint32_t b = 0; b <= bind.cardinality; b++) {
			ix.bin_lo.push_back(bind.base + NumericCast<int64_t>(b) * width);
		}
		for (idx_t r = 0; r < vals.size(); r++) {
			if (vals[r] == null_key) {
				continue;
			}
			if (vals[r] < ix.bin_lo.front() || vals[r] >= ix.bin_lo.back()) {
				throw InvalidInputException("cubit_load: key %lld of \"%s\" lies outside the binned domain [%lld, %lld)",
				                            (long long)vals[r], bind.key, (long long)ix.bin_lo.front(), (long long)ix.bin_lo.back());
			}
			bins[r] = (vals[r] - bind.base) / width;
		}
	}
	ix.base_value = 0;
	ix.cardinality = NumericCast<uint32_t>(ix.bin_lo.size() - 1);
	const int32_t tmp_col = NumericCast<int32_t>(gpu.column_names.size()); // first unused GPU column id
	CubitCheck(cubit_gpu_upload_column(gpu.handle, tmp_col, bins.data(), 8, gpu.row_count));
	CubitCheck(cubit_gpu_index_create(gpu.handle, ix.cardinality, &ix.index_id));
	const int rc = cubit_gpu_index_build(gpu.handle, ix.index_id, tmp_col, 0);
	cubit_gpu_drop_column(gpu.handle, tmp_col);
	CubitCheck(rc);
	return ix;
}

static void CubitLoadFunction(ClientContext &context, TableFunctionInput &data_p, DataChunk &output) {
	auto &bind = data_p.bind_data->CastNoConst<CubitLoadBindData>();
	if (bind.done) {
		return;
	}
	{ // the table is already resident (any DML would have dropped it): just index one more of its columns
		shared_ptr<CubitGpuTable> have;
		{
			std::lock_guard<std::mutex> lk(cubit_registry_lock);
			auto it = cubit_registry.find(bind.table);
			if (it != cubit_registry.end()) {
				have = it->second;
			}
		}
		if (have) {
			idx_t gcol = 0;
			while (gcol < have->column_names.size() && have->column_names[gcol] != bind.key) {
				gcol++;
			}
			bool known = gcol == have->column_names.size();
			for (auto &ix : have->indexes) {
				known |= ix.key_table_column == have->table_column[gcol < have->table_column.size() ? gcol : 0];
			}
			if (!known) {
				auto &entry = Catalog::GetEntry<TableCatalogEntry>(context, INVALID_CATALOG, DEFAULT_SCHEMA, bind.table);
				auto ix = CubitBuildIndex(*have, gcol, bind, entry.GetColumns().GetColumn(LogicalIndex(have->table_column[gcol])).Type());
				{
					std::lock_guard<std::mutex> lk(cubit_registry_lock);
					have->indexes.push_back(ix);
				}
				output.SetValue(0, 0, Value::BIGINT(NumericCast<int64_t>(have->row_count)));
				output.SetCardinality(1);
				bind.done = true;
				return;
			}
			// same key again (or an unknown column): reload from scratch below
		}
	}
	// Pull the key and every BIGINT-castable column in row order through a second connection
	// (rows come back in insertion order: physical_result_collector.cpp:21-45).
	// (only the columns the GPU can hold are selected: integral, DECIMAL and DATE ones)
	auto gpu = make_shared_ptr<CubitGpuTable>();
	vector<idx_t> int_cols; // result column of every uploaded column (= its position in the select list)
	vector<LogicalType> col_types;
	idx_t key_col = DConstants::INVALID_INDEX;
	string select_list;
	{
		auto &entry = Catalog::GetEntry<TableCatalogEntry>(context, INVALID_CATALOG, DEFAULT_SCHEMA, bind.table);
		for (auto &col : entry.GetColumns().Logical()) {
			auto &t = col.Type();
			if (t.IsIntegral() || t.id() == LogicalTypeId::DECIMAL || t.id() == LogicalTypeId::DATE) {
				if (col.Name() == bind.key) {
					key_col = int_cols.size();
				}
				select_list += (select_list.empty() ? "" : ", ") + KeywordHelper::WriteOptionallyQuoted(col.Name());
				int_cols.push_back(int_cols.size());
				col_types.push_back(t);
				gpu->column_names.push_back(col.Name());
				gpu->table_column.push_back(col.Logical().index);
			}
		}
	}
	if (key_col == DConstants::INVALID_INDEX) {
		throw InvalidInputException("cubit_load: key column \"%s\" not found or not integral", bind.key);
	}
	Connection con(*context.db);
	auto res = con.Query("SELECT " + select_list + " FROM " + KeywordHelper::WriteOptionallyQuoted(bind.table));
	if (res->HasError()) {
		throw InvalidInputException("cubit_load: %s", res->GetError());
	}
	vector<vector<int64_t>> cols(int_cols.size());
	// NULLs: one validity mask per column in the reference's own layout (ValidityMask words), built only for
	// columns that hold a NULL; NULL keys are not indexed (plan_create_index.cpp:60-78 filters them out)
	vector<vector<uint64_t>> valid(int_cols.size());
	idx_t rows_seen = 0;
	for (auto &chunk : res->Collection().Chunks()) {
		for (idx_t k = 0; k < int_cols.size(); k++) {
			auto &vec = chunk.data[int_cols[k]];
			// DECIMAL(15,2) is stored as int64 cents (dbgen.cpp:48-50); cast everything to the raw BIGINT
			Vector as_bigint(LogicalType::BIGINT);
			if (vec.GetType().id() == LogicalTypeId::DECIMAL && vec.GetType().InternalType() == PhysicalType::INT64) {
				as_bigint.Reinterpret(vec);
			} else if (vec.GetType().id() == LogicalTypeId::DATE) {
				Vector as_int(LogicalType::INTEGER); // days since 1970-01-01 (date_t)
				as_int.Reinterpret(vec);
				VectorOperations::Cast(context, as_int, as_bigint, chunk.size());
			} else {
				VectorOperations::Cast(context, vec, as_bigint, chunk.size());
			}
			as_bigint.Flatten(chunk.size());
			auto ptr = FlatVector::GetData<int64_t>(as_bigint);
			cols[k].insert(cols[k].end(), ptr, ptr + chunk.size());
			auto &mask = FlatVector::Validity(as_bigint);
			if (!mask.AllValid()) {
				for (idx_t r = 0; r < chunk.size(); r++) {
					if (mask.RowIsValid(r)) {
						continue;
					}
					if (valid[k].empty()) {
						valid[k].assign((res->RowCount() + 63) / 64, ~uint64_t(0));
					}
					const idx_t row = rows_seen + r;
					valid[k][row / 64] &= ~(uint64_t(1) << (row % 64));
					// NULL rows carry a value outside every indexed domain (the probe never shows it: the
					// validity mask does), so any column can become an index key later
					cols[k][row] = NumericLimits<int64_t>::Minimum();
				}
			}
		}
		rows_seen += chunk.size();
	}
	gpu->row_count = cols.empty() ? 0 : cols[0].size();
	CubitCheck(cubit_gpu_create(0, gpu->row_count, 0, 65536, &gpu->handle));
	for (idx_t k = 0; k < cols.size(); k++) {
		// compressed segments straight from the buffer manager when the column qualifies, decoded rows otherwise
		// (a column with NULLs goes the decoded way: its NULL rows must carry an out-of-domain value)
		const bool null_keys = !valid[k].empty();
		if (null_keys || !CubitUploadColumnSegments(context, bind.table, gpu->table_column[k], col_types[k],
		                                            gpu->handle, NumericCast<int32_t>(k), gpu->row_count)) {
			CubitCheck(cubit_gpu_upload_column(gpu->handle, NumericCast<int32_t>(k), cols[k].data(), 8, gpu->row_count));
		}
		if (!valid[k].empty()) {
			CubitCheck(cubit_gpu_upload_column_validity(gpu->handle, NumericCast<int32_t>(k), valid[k].data(),
			                                            valid[k].size()));
		}
	}
	gpu->indexes.push_back(CubitBuildIndex(*gpu, key_col, bind, col_types[key_col]));
	{
		std::lock_guard<std::mutex> lk(cubit_registry_lock);
		cubit_registry[bind.table] = gpu;
	}
	output.SetValue(0, 0, Value::BIGINT(NumericCast<int64_t>(gpu->row_count)));
	output.SetCardinality(1);
	bind.done = true;
}

// ---------------------------------------------------------------- cubit_scan(table, lo, hi)
struct CubitScanBindData : public TableFunctionData {
	shared_ptr<CubitGpuTable> gpu;
	// AND of one range per index: lo <= key <= hi → OR over the value bitvectors in [lo, hi]
	struct Range {
		idx_t index_slot;
		int64_t lo, hi;
	};
	vector<Range> ranges;
	// aggregate push-down (cubit_agg)
	int32_t agg_col = -1;
	// set by the optimizer rewrite: column_ids are TABLE column indexes and must be mapped to GPU column ids
	bool table_column_ids = false;
};

struct CubitScanGlobalState : public GlobalTableFunctionState {
	cubit_gpu_result *result = nullptr;
	vector<column_t> column_ids;
	vector<idx_t> out_slot; // which entry of column_ids every output vector shows (projection_ids applied)
	idx_t row_count = 0, offset = 0;
	uint64_t sum_lo = 0;
	int64_t sum_hi = 0;
	uint64_t agg_rows = 0; // non-NULL inputs of the pushed-down SUM
	bool agg_emitted = false;
	// host staging window: result rows [win_begin, win_end) fetched with ONE device→host copy per column and
	// served to the executor 2048 rows at a time (a copy + synchronise per DataChunk costs ~30 us, i.e. more
	// than the whole scan for a few hundred thousand rows)
	static constexpr idx_t WINDOW_ROWS = 64 * STANDARD_VECTOR_SIZE;
	idx_t win_begin = 0, win_end = 0;
	// page-locked, WINDOW_ROWS int64 each, borrowed from the table's pool on first use
	shared_ptr<CubitGpuTable> pool_owner;
	int64_t *win_rowids = nullptr;
	vector<int64_t *> win_cols;             // one per projected value column
	vector<vector<uint64_t>> win_validity;  // ValidityMask words of the window, empty = no NULL in the window
	~CubitScanGlobalState() override {
		cubit_gpu_free_result(result);
		if (pool_owner) {
			pool_owner->ReleaseWindow(win_rowids);
			for (auto p : win_cols) {
				pool_owner->ReleaseWindow(p);
			}
		}
	}
	idx_t MaxThreads() const override {
		return 1; // like index_scan (table_scan.cpp:213-225)
	}
};

static unique_ptr<FunctionData> CubitScanBind(ClientContext &, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(input.inputs[0].GetValue<string>());
	bind->ranges.push_back({0, input.inputs[1].GetValue<int64_t>(), input.inputs[2].GetValue<int64_t>()});
	for (auto &n : bind->gpu->column_names) {
		return_types.emplace_back(LogicalType::BIGINT);
		names.emplace_back(n);
	}
	return std::move(bind);
}

static unique_ptr<GlobalTableFunctionState> CubitRunQuery(const CubitScanBindData &bind,
                                                          const vector<column_t> &column_ids_p,
                                                          const vector<idx_t> &projection_ids) {
	auto state = make_uniq<CubitScanGlobalState>();
	state->pool_owner = bind.gpu;
	auto &gpu = *bind.gpu;
	// the columns that actually leave the scan (filter-only columns are pruned: projection_ids)
	vector<column_t> column_ids;
	for (idx_t i = 0; i < (projection_ids.empty() ? column_ids_p.size() : projection_ids.size()); i++) {
		column_t c = column_ids_p[projection_ids.empty() ? i : projection_ids[i]];
		if (bind.table_column_ids && c != COLUMN_IDENTIFIER_ROW_ID) {
			idx_t g = 0;
			while (g < gpu.table_column.size() && gpu.table_column[g] != c) {
				g++;
			}
			if (g == gpu.table_column.size()) {
				throw InternalException("cubit: column %llu is not resident on the GPU", c);
			}
			c = g;
		}
		column_ids.push_back(c);
	}
	state->column_ids = column_ids;
	// one OR group per range, AND across the groups (Q = AND_j OR_{v in [lo_j, hi_j]} B_v)
	vector<vector<cubit_bv_ref>> refs(bind.ranges.size());
	vector<cubit_pred_group> groups;
	for (idx_t j = 0; j < bind.ranges.size(); j++) {
		auto &ix = gpu.indexes[bind.ranges[j].index_slot];
		const int64_t lo = MaxValue<int64_t>(bind.ranges[j].lo, ix.base_value);
		const int64_t hi = MinValue<int64_t>(bind.ranges[j].hi, ix.base_value + ix.cardinality - 1);
		if (lo > hi) {
			return std::move(state); // an empty range empties the conjunction
		}
		for (int64_t v = lo; v <= hi; v++) {
			refs[j].push_back(cubit_bv_ref {ix.index_id, NumericCast<uint32_t>(v - ix.base_value)});
		}
		groups.push_back(cubit_pred_group {NumericCast<uint32_t>(refs[j].size()), refs[j].data()});
	}
	vector<int32_t> cols;
	bool want_rowid = false;
	for (auto c : column_ids) {
		if (c == COLUMN_IDENTIFIER_ROW_ID) {
			want_rowid = true;
		} else {
			cols.push_back(NumericCast<int32_t>(c));
		}
	}
	cubit_query q {};
	q.n_groups = NumericCast<uint32_t>(groups.size());
	q.groups = groups.data();
	if (bind.agg_col >= 0) {
		q.agg_kind = CUBIT_AGG_SUM;
		q.agg_col_a = bind.agg_col;
	} else {
		q.flags = (want_rowid ? CUBIT_Q_ROWIDS : 0u) | (cols.empty() ? 0u : CUBIT_Q_VALUES);
		q.n_cols = NumericCast<uint32_t>(cols.size());
		q.cols = cols.data();
	}
	CubitCheck(cubit_gpu_query(gpu.handle, &q, &state->result));
	cubit_result_info info;
	CubitCheck(cubit_gpu_result_get(state->result, &info));
	state->row_count = info.count;
	state->sum_lo = info.sum_lo;
	state->sum_hi = info.sum_hi;
	state->agg_rows = info.agg_rows;
	return std::move(state);
}

static unique_ptr<GlobalTableFunctionState> CubitScanInitGlobal(ClientContext &, TableFunctionInitInput &input) {
	return CubitRunQuery(input.bind_data->Cast<CubitScanBindData>(), input.column_ids, input.projection_ids);
}

static void CubitScanFunction(ClientContext &, TableFunctionInput &data_p, DataChunk &output) {
	auto &state = data_p.global_state->Cast<CubitScanGlobalState>();
	if (state.offset >= state.row_count) {
		return; // chunk.size() == 0 → PhysicalTableScan::GetData returns FINISHED
	}
	if (state.offset >= state.win_end) { // refill the staging window
		const idx_t n = MinValue<idx_t>(CubitScanGlobalState::WINDOW_ROWS, state.row_count - state.offset);
		state.win_begin = state.offset;
		state.win_end = state.offset + n;
		bool want_rowid = false;
		idx_t n_value_cols = 0;
		for (auto c : state.column_ids) {
			want_rowid |= c == COLUMN_IDENTIFIER_ROW_ID;
			n_value_cols += c != COLUMN_IDENTIFIER_ROW_ID;
		}
		state.win_validity.resize(n_value_cols);
		const uint64_t win_bytes = CubitScanGlobalState::WINDOW_ROWS * sizeof(int64_t);
		while (state.win_cols.size() < n_value_cols) {
			state.win_cols.push_back(static_cast<int64_t *>(state.pool_owner->AcquireWindow(win_bytes)));
		}
		if (want_rowid && !state.win_rowids) {
			state.win_rowids = static_cast<int64_t *>(state.pool_owner->AcquireWindow(win_bytes));
		}
		vector<void *> ptrs(state.win_cols.begin(), state.win_cols.end());
		CubitCheck(cubit_gpu_fetch(state.result, state.win_begin, n, want_rowid ? state.win_rowids : nullptr,
		                           NumericCast<uint32_t>(ptrs.size()), ptrs.data()));
		// NULLs: the validity mask of every projected value (StandardColumnData::FetchRow = validity + data)
		for (idx_t c = 0; c < n_value_cols; c++) {
			int all_valid = 1;
			state.win_validity[c].assign((n + 63) / 64, 0);
			CubitCheck(cubit_gpu_fetch_validity(state.result, NumericCast<uint32_t>(c), state.win_begin, n,
			                                    state.win_validity[c].data(), &all_valid));
			if (all_valid) {
				state.win_validity[c].clear();
			}
		}
	}
	const idx_t scan_count = MinValue<idx_t>(STANDARD_VECTOR_SIZE, state.win_end - state.offset);
	const idx_t rel = state.offset - state.win_begin; // a multiple of 2048: word aligned in the window's masks
	idx_t value_col = 0;
	for (idx_t i = 0; i < state.column_ids.size(); i++) {
		auto dst = FlatVector::GetData<int64_t>(output.data[i]);
		if (state.column_ids[i] == COLUMN_IDENTIFIER_ROW_ID) {
			memcpy(dst, state.win_rowids + rel, scan_count * sizeof(int64_t));
			continue;
		}
		memcpy(dst, state.win_cols[value_col] + rel, scan_count * sizeof(int64_t));
		auto &words = state.win_validity[value_col];
		if (!words.empty()) {
			auto &mask = FlatVector::Validity(output.data[i]);
			for (idx_t r = 0; r < scan_count; r++) {
				if (!((words[(rel + r) / 64] >> ((rel + r) % 64)) & 1)) {
					mask.SetInvalid(r);
				}
			}
		}
		value_col++;
	}
	output.SetCardinality(scan_count);
	state.offset += scan_count;
}

// ---------------------------------------------------------------- cubit_agg(table, lo, hi, column) → (count, sum)
static unique_ptr<FunctionData> CubitAggBind(ClientContext &, TableFunctionBindInput &input,
                                             vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(input.inputs[0].GetValue<string>());
	bind->ranges.push_back({0, input.inputs[1].GetValue<int64_t>(), input.inputs[2].GetValue<int64_t>()});
	auto col = input.inputs[3].GetValue<string>();
	for (idx_t c = 0; c < bind->gpu->column_names.size(); c++) {
		if (bind->gpu->column_names[c] == col) {
			bind->agg_col = NumericCast<int32_t>(c);
		}
	}
	if (bind->agg_col < 0) {
		throw InvalidInputException("cubit_agg: column \"%s\" is not resident on the GPU", col);
	}
	return_types = {LogicalType::BIGINT, LogicalType::HUGEINT};
	names = {"count", "sum"};
	return std::move(bind);
}

static void CubitAggFunction(ClientContext &, TableFunctionInput &data_p, DataChunk &output) {
	auto &state = data_p.global_state->Cast<CubitScanGlobalState>();
	if (state.agg_emitted) {
		return;
	}
	output.SetValue(0, 0, Value::BIGINT(NumericCast<int64_t>(state.row_count)));
	hugeint_t sum;
	sum.lower = state.sum_lo;
	sum.upper = state.sum_hi;
	// SUM over no non-NULL input is NULL (sum.cpp: the state is only "set" by a valid row)
	output.SetValue(1, 0, state.agg_rows ? Value::HUGEINT(sum) : Value(LogicalType::HUGEINT));
	output.SetCardinality(1);
	state.agg_emitted = true;
}

// ---------------------------------------------------------------- transparent rewrite of seq_scan → cubit_scan
static std::atomic<idx_t> cubit_rewrite_count {0};
idx_t CubitRewriteCount() {
	return cubit_rewrite_count.load();
}

static TableFunction CubitScanTableFunction();

// lo <= key <= hi from the pushed-down filter of ONE column; false if the shape is not supported
static bool CubitBoundsFromFilter(const TableFilter &filter, int64_t &lo, int64_t &hi) {
	switch (filter.filter_type) {
	case TableFilterType::IS_NOT_NULL:
		return true; // NULL keys are not indexed
	case TableFilterType::CONJUNCTION_AND: {
		for (auto &child : filter.Cast<ConjunctionAndFilter>().child_filters) {
			if (!CubitBoundsFromFilter(*child, lo, hi)) {
				return false;
			}
		}
		return true;
	}
	case TableFilterType::CONSTANT_COMPARISON: {
		auto &cf = filter.Cast<ConstantFilter>();
		int64_t c;
		if (cf.constant.type().id() == LogicalTypeId::DATE) {
			c = cf.constant.GetValue<date_t>().days; // DATE keys are indexed as days since 1970-01-01
		} else if (!cf.constant.type().IsIntegral() && cf.constant.type().id() != LogicalTypeId::DECIMAL) {
			return false;
		} else if (!Hugeint::TryCast(IntegralValue::Get(cf.constant), c)) { // raw integer (DECIMAL: unscaled cents)
			return false;
		}
		switch (cf.comparison_type) {
		case ExpressionType::COMPARE_EQUAL:
			lo = MaxValue(lo, c);
			hi = MinValue(hi, c);
			return true;
		case ExpressionType::COMPARE_GREATERTHAN:
			lo = MaxValue(lo, c + 1);
			return true;
		case ExpressionType::COMPARE_GREATERTHANOREQUALTO:
			lo = MaxValue(lo, c);
			return true;
		case ExpressionType::COMPARE_LESSTHAN:
			hi = MinValue(hi, c - 1);
			return true;
		case ExpressionType::COMPARE_LESSTHANOREQUALTO:
			hi = MinValue(hi, c);
			return true;
		default:
			return false;
		}
	}
	default:
		return false;
	}
}

// Can this seq_scan's pushed-down filters be answered by the table's GPU indexes?  Fills the table and one
// (index, lo, hi) range per filtered column.  Nothing is modified.
static bool CubitPlanGet(LogicalGet &get, shared_ptr<CubitGpuTable> &gpu, vector<CubitScanBindData::Range> &ranges) {
#undef CUBIT_WHY
#define CUBIT_WHY(msg)                                                                                                  \
	do {                                                                                                               \
		if (getenv("CUBIT_DEBUG_REWRITE")) {                                                                           \
			fprintf(stderr, "cubit rewrite skipped: %s\n", msg);                                                      \
		}                                                                                                              \
		return false;                                                                                                  \
	} while (0)
	if (get.function.name != "seq_scan") {
		CUBIT_WHY(get.function.name.c_str());
	}
	if (get.table_filters.filters.empty()) {
		CUBIT_WHY("no pushed-down filter");
	}
	auto table = get.GetTable();
	if (!table) {
		CUBIT_WHY("no table");
	}
	{
		std::lock_guard<std::mutex> lk(cubit_registry_lock);
		auto it = cubit_registry.find(table->name);
		if (it == cubit_registry.end()) {
			CUBIT_WHY("table has no GPU index");
		}
		gpu = it->second;
	}
	// EVERY pushed-down filter must sit on an indexed column (keys of LogicalGet::table_filters are table column
	// indexes: filter_combiner.cpp:438-480, plan_get.cpp:15-33); each becomes one OR group over the value
	// bitvectors of its range, the groups are ANDed — the Q6-style conjunction of range predicates
	idx_t n_streams = 0;
	for (auto &entry : get.table_filters.filters) {
		idx_t slot = 0;
		while (slot < gpu->indexes.size() && gpu->indexes[slot].key_table_column != entry.first) {
			slot++;
		}
		if (slot == gpu->indexes.size()) {
			CUBIT_WHY("a filter sits on a column without a GPU index");
		}
		int64_t lo = NumericLimits<int64_t>::Minimum() + 1, hi = NumericLimits<int64_t>::Maximum() - 1;
		if (!CubitBoundsFromFilter(*entry.second, lo, hi)) {
			CUBIT_WHY("unsupported filter shape");
		}
		auto &ix = gpu->indexes[slot];
		if (!ix.bin_lo.empty()) {
			// binned index: exact only when both ends of the range fall on bin boundaries (every key lies inside
			// the binned domain — checked at load — so a range that starts before / ends after it is aligned too)
			int64_t first = 0, last = NumericCast<int64_t>(ix.cardinality) - 1;
			if (lo > ix.bin_lo.front()) {
				auto it = std::lower_bound(ix.bin_lo.begin(), ix.bin_lo.end(), lo);
				if (it == ix.bin_lo.end() || *it != lo) {
					if (lo >= ix.bin_lo.back()) {
						first = last + 1; // empty
					} else {
						CUBIT_WHY("range start is not a bin boundary of the binned index");
					}
				} else {
					first = it - ix.bin_lo.begin();
				}
			}
			if (hi < ix.bin_lo.back() - 1) {
				auto it = std::lower_bound(ix.bin_lo.begin(), ix.bin_lo.end(), hi + 1);
				if (it == ix.bin_lo.end() || *it != hi + 1) {
					if (hi < ix.bin_lo.front()) {
						last = -1; // empty
					} else {
						CUBIT_WHY("range end is not a bin boundary of the binned index");
					}
				} else {
					last = (it - ix.bin_lo.begin()) - 1;
				}
			}
			lo = first; // ranges are kept in the index's key units: bin ids here
			hi = last;
		}
		const int64_t clo = MaxValue<int64_t>(lo, ix.base_value), chi = MinValue<int64_t>(hi, ix.base_value + ix.cardinality - 1);
		n_streams += chi >= clo ? NumericCast<idx_t>(chi - clo + 1) : 0;
		ranges.push_back({slot, lo, hi});
	}
	if (n_streams > CUBIT_MAX_STREAMS) {
		CUBIT_WHY("predicate reads more value bitvectors than one scan merges");
	}
	return true;
}

// GPU column id of table column `c` if it is resident and physically INT64 (BIGINT, DECIMAL(≤18)), else -1
static int32_t CubitResidentColumn(const LogicalGet &get, const CubitGpuTable &gpu, column_t c) {
	if (c == COLUMN_IDENTIFIER_ROW_ID || c >= get.returned_types.size() ||
	    get.returned_types[c].InternalType() != PhysicalType::INT64) {
		return -1;
	}
	for (idx_t g = 0; g < gpu.table_column.size(); g++) {
		if (gpu.table_column[g] == c) {
			return NumericCast<int32_t>(g);
		}
	}
	return -1;
}

static void CubitRewriteGet(LogicalGet &get) {
	shared_ptr<CubitGpuTable> gpu;
	vector<CubitScanBindData::Range> ranges;
	if (!CubitPlanGet(get, gpu, ranges)) {
		return;
	}
	// every column that leaves the scan must be GPU resident and physically int64 (BIGINT, DECIMAL(≤18))
	for (idx_t i = 0; i < (get.projection_ids.empty() ? get.column_ids.size() : get.projection_ids.size()); i++) {
		const column_t c = get.column_ids[get.projection_ids.empty() ? i : get.projection_ids[i]];
		if (c == COLUMN_IDENTIFIER_ROW_ID) {
			continue;
		}
		if (CubitResidentColumn(get, *gpu, c) < 0) {
			if (getenv("CUBIT_DEBUG_REWRITE")) {
				fprintf(stderr, "cubit rewrite skipped: projected column is not a GPU-resident INT64 column\n");
			}
			return;
		}
	}
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = gpu;
	bind->ranges = std::move(ranges);
	bind->table_column_ids = true;
	get.function = CubitScanTableFunction();
	get.bind_data = std::move(bind);
	get.table_filters.filters.clear(); // applied exactly by the bitmap scan
	cubit_rewrite_count++;
}

// ---------------------------------------------------------------- aggregate push-down
// An ungrouped aggregate of COUNT(*) / COUNT(col) / SUM(col) / SUM(a * b) sitting directly on a rewritable scan is
// answered on the GPU in ONE row: no row ever crosses PCIe.  The aggregate node is replaced by
// a table function that carries the aggregate's own table index and return types, so every reference above it
// stays valid.  SUM semantics are the reference's: int64 inputs into a 128-bit sum (sum.cpp:172-199), int64 product
// with overflow error (arithmetic.cpp:766-795), NULL inputs skipped, NULL over no input.
struct CubitAggSpec {
	enum Kind : uint8_t { COUNT_STAR, COUNT_COL, SUM_COL, SUM_PROD } kind;
	int32_t col_a = -1, col_b = -1;
};

struct CubitAggMultiBindData : public TableFunctionData {
	shared_ptr<CubitGpuTable> gpu;
	vector<CubitScanBindData::Range> ranges;
	vector<CubitAggSpec> specs;
	vector<LogicalType> types;
};

struct CubitAggMultiState : public GlobalTableFunctionState {
	bool emitted = false;
};

static std::atomic<idx_t> cubit_agg_pushdown_count {0};
idx_t CubitAggPushdownCount() {
	return cubit_agg_pushdown_count.load();
}

static unique_ptr<GlobalTableFunctionState> CubitAggMultiInit(ClientContext &, TableFunctionInitInput &) {
	return make_uniq<CubitAggMultiState>();
}

static void CubitAggMultiFunction(ClientContext &, TableFunctionInput &data_p, DataChunk &output) {
	auto &state = data_p.global_state->Cast<CubitAggMultiState>();
	if (state.emitted) {
		return;
	}
	state.emitted = true;
	auto &bind = data_p.bind_data->Cast<CubitAggMultiBindData>();
	auto &gpu = *bind.gpu;
	vector<vector<cubit_bv_ref>> refs(bind.ranges.size());
	vector<cubit_pred_group> groups;
	bool empty = false;
	for (idx_t j = 0; j < bind.ranges.size(); j++) {
		auto &ix = gpu.indexes[bind.ranges[j].index_slot];
		const int64_t lo = MaxValue<int64_t>(bind.ranges[j].lo, ix.base_value);
		const int64_t hi = MinValue<int64_t>(bind.ranges[j].hi, ix.base_value + ix.cardinality - 1);
		empty |= lo > hi;
		for (int64_t v = lo; v <= hi; v++) {
			refs[j].push_back(cubit_bv_ref {ix.index_id, NumericCast<uint32_t>(v - ix.base_value)});
		}
		groups.push_back(cubit_pred_group {NumericCast<uint32_t>(refs[j].size()), refs[j].data()});
	}
	// one GPU query per aggregate (merge + bit-driven probe with the fused SUM; a few tens of microseconds each);
	// COUNT(*) rides on any of them (every query reports the selection's row count)
	bool have_count = false;
	uint64_t row_count = 0;
	vector<idx_t> order;
	for (idx_t a = 0; a < bind.specs.size(); a++) {
		if (bind.specs[a].kind != CubitAggSpec::COUNT_STAR) {
			order.push_back(a);
		}
	}
	for (idx_t a = 0; a < bind.specs.size(); a++) {
		if (bind.specs[a].kind == CubitAggSpec::COUNT_STAR) {
			order.push_back(a);
		}
	}
	for (auto a : order) {
		auto &spec = bind.specs[a];
		cubit_result_info info;
		memset(&info, 0, sizeof(info));
		if (spec.kind == CubitAggSpec::COUNT_STAR && have_count) {
			info.count = row_count;
		} else if (!empty) {
			cubit_query q {};
			q.n_groups = NumericCast<uint32_t>(groups.size());
			q.groups = groups.data();
			if (spec.kind != CubitAggSpec::COUNT_STAR) {
				q.agg_kind = spec.kind == CubitAggSpec::SUM_PROD ? CUBIT_AGG_SUM_PROD : CUBIT_AGG_SUM;
				q.agg_col_a = spec.col_a;
				q.agg_col_b = spec.col_b;
			}
			cubit_gpu_result *res = nullptr;
			CubitCheck(cubit_gpu_query(gpu.handle, &q, &res));
			const int rc = cubit_gpu_result_get(res, &info);
			cubit_gpu_free_result(res);
			CubitCheck(rc);
			have_count = true;
			row_count = info.count;
		}
		auto &vec = output.data[a];
		if (spec.kind == CubitAggSpec::COUNT_STAR || spec.kind == CubitAggSpec::COUNT_COL) {
			const uint64_t n = spec.kind == CubitAggSpec::COUNT_STAR ? info.count : info.agg_rows;
			FlatVector::GetData<int64_t>(vec)[0] = NumericCast<int64_t>(n);
		} else if (info.agg_rows == 0) {
			FlatVector::SetNull(vec, 0, true); // SUM over no non-NULL input
		} else if (bind.types[a].InternalType() == PhysicalType::INT128) {
			hugeint_t sum;
			sum.lower = info.sum_lo;
			sum.upper = info.sum_hi;
			FlatVector::GetData<hugeint_t>(vec)[0] = sum;
		} else { // an INT64-typed sum (sum_no_overflow's narrow form): the statistics guarantee that it fits
			if ((info.sum_hi != 0 || (info.sum_lo >> 63)) && (info.sum_hi != -1 || !(info.sum_lo >> 63))) {
				throw OutOfRangeException("cubit: SUM does not fit the INT64 result type");
			}
			FlatVector::GetData<int64_t>(vec)[0] = static_cast<int64_t>(info.sum_lo);
		}
	}
	output.SetCardinality(1);
}

// the GPU column an aggregate input reads: a column reference of the scan, possibly under casts that keep the
// stored int64 as it is (DECIMAL(15,2) → DECIMAL(18,2): same scale, same physical type)
static int32_t CubitAggInputColumn(const Expression &expr, const LogicalGet &get, const CubitGpuTable &gpu) {
	const Expression *e = &expr;
	while (e->expression_class == ExpressionClass::BOUND_CAST) {
		auto &cast = e->Cast<BoundCastExpression>();
		const LogicalType &from = cast.child->return_type;
		const LogicalType &to = cast.return_type;
		const bool same_scale = (from.id() == LogicalTypeId::DECIMAL ? DecimalType::GetScale(from) : 0) ==
		                        (to.id() == LogicalTypeId::DECIMAL ? DecimalType::GetScale(to) : 0);
		if (from.InternalType() != PhysicalType::INT64 || to.InternalType() != PhysicalType::INT64 || !same_scale ||
		    (!from.IsIntegral() && from.id() != LogicalTypeId::DECIMAL) || (!to.IsIntegral() && to.id() != LogicalTypeId::DECIMAL)) {
			return -1;
		}
		e = cast.child.get();
	}
	if (e->expression_class != ExpressionClass::BOUND_COLUMN_REF) {
		return -1;
	}
	auto &ref = e->Cast<BoundColumnRefExpression>();
	if (ref.binding.table_index != get.table_index || ref.depth != 0) {
		return -1;
	}
	// a binding's column_index is a position in column_ids, with or without projection_ids
	// (LogicalGet::GetColumnBindings, src/planner/operator/logical_get.cpp)
	if (ref.binding.column_index >= get.column_ids.size()) {
		return -1;
	}
	return CubitResidentColumn(get, gpu, get.column_ids[ref.binding.column_index]);
}

static bool CubitTryAggregatePushdown(unique_ptr<LogicalOperator> &op) {
	if (op->type != LogicalOperatorType::LOGICAL_AGGREGATE_AND_GROUP_BY) {
		return false;
	}
	auto &aggr = op->Cast<LogicalAggregate>();
	if (!aggr.groups.empty() || !aggr.grouping_functions.empty() || aggr.grouping_sets.size() > 1 ||
	    aggr.children.size() != 1 || aggr.children[0]->type != LogicalOperatorType::LOGICAL_GET || aggr.expressions.empty()) {
		return false;
	}
	auto &get = aggr.children[0]->Cast<LogicalGet>();
	shared_ptr<CubitGpuTable> gpu;
	vector<CubitScanBindData::Range> ranges;
	if (!CubitPlanGet(get, gpu, ranges)) {
		return false;
	}
	auto bind = make_uniq<CubitAggMultiBindData>();
	vector<string> names;
	for (auto &expr : aggr.expressions) {
		if (expr->expression_class != ExpressionClass::BOUND_AGGREGATE) {
			return false;
		}
		auto &a = expr->Cast<BoundAggregateExpression>();
		if (a.IsDistinct() || a.filter || a.order_bys) {
			return false;
		}
		CubitAggSpec spec;
		const auto &fn = a.function.name;
		const auto phys = a.return_type.InternalType();
		if (fn == "count_star" && a.children.empty()) {
			spec.kind = CubitAggSpec::COUNT_STAR;
		} else if (fn == "count" && a.children.size() == 1) {
			spec.kind = CubitAggSpec::COUNT_COL;
			spec.col_a = CubitAggInputColumn(*a.children[0], get, *gpu);
		} else if ((fn == "sum" || fn == "sum_no_overflow") && a.children.size() == 1 &&
		           (phys == PhysicalType::INT128 || phys == PhysicalType::INT64) &&
		           (a.return_type.id() == LogicalTypeId::DECIMAL || a.return_type.id() == LogicalTypeId::HUGEINT ||
		            a.return_type.id() == LogicalTypeId::BIGINT)) {
			auto &child = *a.children[0];
			if (child.expression_class == ExpressionClass::BOUND_FUNCTION && child.Cast<BoundFunctionExpression>().function.name == "*" &&
			    child.Cast<BoundFunctionExpression>().children.size() == 2 && child.return_type.InternalType() == PhysicalType::INT64) {
				auto &mul = child.Cast<BoundFunctionExpression>();
				spec.kind = CubitAggSpec::SUM_PROD; // int64 product, overflow is an error on both sides
				spec.col_a = CubitAggInputColumn(*mul.children[0], get, *gpu);
				spec.col_b = CubitAggInputColumn(*mul.children[1], get, *gpu);
				if (spec.col_b < 0) {
					return false;
				}
			} else {
				spec.kind = CubitAggSpec::SUM_COL;
				spec.col_a = CubitAggInputColumn(child, get, *gpu);
			}
		} else {
			if (getenv("CUBIT_DEBUG_REWRITE")) {
				fprintf(stderr, "cubit aggregate push-down skipped: unsupported aggregate %s -> %s\n", a.ToString().c_str(),
				        a.return_type.ToString().c_str());
			}
			return false;
		}
		if (spec.kind != CubitAggSpec::COUNT_STAR && spec.col_a < 0) {
			if (getenv("CUBIT_DEBUG_REWRITE")) {
				fprintf(stderr, "cubit aggregate push-down skipped: input of %s is not a GPU-resident INT64 column\n",
				        a.ToString().c_str());
			}
			return false;
		}
		bind->specs.push_back(spec);
		bind->types.push_back(a.return_type);
		names.push_back(a.ToString());
	}
	bind->gpu = gpu;
	bind->ranges = std::move(ranges);
	TableFunction fn("cubit_agg_pushdown", {}, CubitAggMultiFunction, nullptr, CubitAggMultiInit);
	auto types = bind->types;
	auto replacement = make_uniq<LogicalGet>(aggr.aggregate_index, fn, std::move(bind), types, names);
	for (idx_t i = 0; i < types.size(); i++) {
		replacement->column_ids.push_back(i);
	}
	op = std::move(replacement);
	cubit_rewrite_count++;
	cubit_agg_pushdown_count++;
	return true;
}

static void CubitRewritePlan(unique_ptr<LogicalOperator> &op) {
	if (CubitTryAggregatePushdown(op)) {
		return;
	}
	if (op->type == LogicalOperatorType::LOGICAL_GET) {
		CubitRewriteGet(op->Cast<LogicalGet>());
	}
	for (auto &child : op->children) {
		CubitRewritePlan(child);
	}
}

// DML on an indexed table: this glue does not sit in the index-maintenance path (that needs a registered index
// type, bound_index.hpp:71-97 → cubit_gpu_set_delta / cubit_gpu_append_rows), so the GPU copy would go stale.
// The statement itself keeps the vanilla scan, and the table's GPU index is dropped: later scans fall back to
// the vanilla path until cubit_load is called again.
static bool CubitInvalidateOnDml(LogicalOperator &op) {
	bool dml = false;
	const TableCatalogEntry *target = nullptr;
	switch (op.type) {
	case LogicalOperatorType::LOGICAL_UPDATE:
		target = &op.Cast<LogicalUpdate>().table;
		break;
	case LogicalOperatorType::LOGICAL_DELETE:
		target = &op.Cast<LogicalDelete>().table;
		break;
	case LogicalOperatorType::LOGICAL_INSERT:
		target = &op.Cast<LogicalInsert>().table;
		break;
	default:
		break;
	}
	if (target) {
		dml = true;
		std::lock_guard<std::mutex> lk(cubit_registry_lock);
		cubit_registry.erase(target->name);
	}
	for (auto &child : op.children) {
		dml |= CubitInvalidateOnDml(*child);
	}
	return dml;
}

static void CubitOptimize(OptimizerExtensionInput &, unique_ptr<LogicalOperator> &plan) {
	if (CubitInvalidateOnDml(*plan)) {
		return;
	}
	if (getenv("CUBIT_NO_AGG_PUSHDOWN")) { // (tests: compare the row-returning scan with the pushed-down aggregate)
		std::function<void(LogicalOperator &)> walk = [&](LogicalOperator &o) {
			if (o.type == LogicalOperatorType::LOGICAL_GET) {
				CubitRewriteGet(o.Cast<LogicalGet>());
			}
			for (auto &child : o.children) {
				walk(*child);
			}
		};
		walk(*plan);
		return;
	}
	CubitRewritePlan(plan);
}

static TableFunction CubitScanTableFunction() {
	TableFunction scan("cubit_scan", {LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT}, CubitScanFunction,
	                   CubitScanBind, CubitScanInitGlobal);
	scan.projection_pushdown = true; // column_ids tell the GPU which columns to probe
	scan.filter_prune = true;        // columns used only by the (absorbed) filter are not produced
	return scan;
}

// ---------------------------------------------------------------- registration
void RegisterCubitGpuFunctions(DatabaseInstance &db) {
	TableFunction load("cubit_load", {LogicalType::VARCHAR, LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT},
	                   CubitLoadFunction, CubitLoadBind);
	load.named_parameters["bin"] = LogicalType::VARCHAR;
	ExtensionUtil::RegisterFunction(db, load);

	ExtensionUtil::RegisterFunction(db, CubitScanTableFunction());

	TableFunction agg("cubit_agg", {LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT, LogicalType::VARCHAR},
	                  CubitAggFunction, CubitAggBind, CubitScanInitGlobal);
	ExtensionUtil::RegisterFunction(db, agg);

	OptimizerExtension rewrite;
	rewrite.optimize_function = CubitOptimize;
	DBConfig::GetConfig(db).optimizer_extensions.push_back(rewrite);
}

} // namespace duckdb

extern "C" {
// loadable-extension entry points (src/main/extension/extension_load.cpp:23-24)
DUCKDB_EXTENSION_API void cubit_gpu_init(duckdb::DatabaseInstance &db) {
	duckdb::RegisterCubitGpuFunctions(db);
}
DUCKDB_EXTENSION_API const char *cubit_gpu_version() {
	return duckdb::DuckDB::LibraryVersion();
}
}
