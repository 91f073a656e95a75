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
#include "duckdb.hpp"
#include "duckdb/function/table_function.hpp"
#include "duckdb/main/extension_util.hpp"
#include "duckdb/common/types/data_chunk.hpp"
#include "duckdb/common/types/vector.hpp"

#include "cubit_gpu.h"

#include <mutex>
#include <unordered_map>

namespace duckdb {

// ---------------------------------------------------------------- registry of GPU-resident tables
struct CubitGpuTable {
	cubit_gpu_table *handle = nullptr;
	int32_t index_id = -1;
	int64_t base_value = 0;
	uint32_t cardinality = 0;
	vector<string> column_names; // uploaded BIGINT columns, column id = position
	idx_t row_count = 0;
	~CubitGpuTable() {
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
	bool done = false;
};

static unique_ptr<FunctionData> CubitLoadBind(ClientContext &, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitLoadBindData>();
	bind->table = input.inputs[0].GetValue<string>();
	bind->key = input.inputs[1].GetValue<string>();
	bind->base = input.inputs[2].GetValue<int64_t>();
	bind->cardinality = NumericCast<uint32_t>(input.inputs[3].GetValue<int64_t>());
	return_types.emplace_back(LogicalType::BIGINT);
	names.emplace_back("rows_indexed");
	return std::move(bind);
}

static void CubitLoadFunction(ClientContext &context, TableFunctionInput &data_p, DataChunk &output) {
	auto &bind = data_p.bind_data->CastNoConst<CubitLoadBindData>();
	if (bind.done) {
		return;
	}
	// Pull the key and every BIGINT-castable column in row order through a second connection
	// (rows come back in insertion order: physical_result_collector.cpp:21-45).
	Connection con(*context.db);
	auto res = con.Query("SELECT * FROM " + KeywordHelper::WriteOptionallyQuoted(bind.table));
	if (res->HasError()) {
		throw InvalidInputException("cubit_load: %s", res->GetError());
	}
	auto gpu = make_shared_ptr<CubitGpuTable>();
	vector<idx_t> int_cols;
	idx_t key_col = DConstants::INVALID_INDEX;
	for (idx_t c = 0; c < res->ColumnCount(); c++) {
		auto &t = res->types[c];
		if (t.IsIntegral() || t.id() == LogicalTypeId::DECIMAL || t.id() == LogicalTypeId::DATE) {
			if (res->names[c] == bind.key) {
				key_col = int_cols.size();
			}
			int_cols.push_back(c);
			gpu->column_names.push_back(res->names[c]);
		}
	}
	if (key_col == DConstants::INVALID_INDEX) {
		throw InvalidInputException("cubit_load: key column \"%s\" not found or not integral", bind.key);
	}
	vector<vector<int64_t>> cols(int_cols.size());
	for (auto &chunk : res->Collection().Chunks()) {
		for (idx_t k = 0; k < int_cols.size(); k++) {
			auto &vec = chunk.data[int_cols[k]];
			// DECIMAL(15,2) is stored as int64 cents (dbgen.cpp:48-50); cast everything to the raw BIGINT
			Vector as_bigint(LogicalType::BIGINT);
			if (vec.GetType().id() == LogicalTypeId::DECIMAL && vec.GetType().InternalType() == PhysicalType::INT64) {
				as_bigint.Reinterpret(vec);
			} else {
				VectorOperations::Cast(context, vec, as_bigint, chunk.size());
			}
			as_bigint.Flatten(chunk.size());
			auto ptr = FlatVector::GetData<int64_t>(as_bigint);
			cols[k].insert(cols[k].end(), ptr, ptr + chunk.size());
		}
	}
	gpu->row_count = cols.empty() ? 0 : cols[0].size();
	gpu->base_value = bind.base;
	gpu->cardinality = bind.cardinality;
	CubitCheck(cubit_gpu_create(0, gpu->row_count, 0, 65536, &gpu->handle));
	for (idx_t k = 0; k < cols.size(); k++) {
		CubitCheck(cubit_gpu_upload_column(gpu->handle, NumericCast<int32_t>(k), cols[k].data(), 8, gpu->row_count));
	}
	CubitCheck(cubit_gpu_index_create(gpu->handle, gpu->cardinality, &gpu->index_id));
	CubitCheck(cubit_gpu_index_build(gpu->handle, gpu->index_id, NumericCast<int32_t>(key_col), gpu->base_value));
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
	int64_t lo = 0, hi = 0;
	// aggregate push-down (cubit_agg)
	int32_t agg_col = -1;
};

struct CubitScanGlobalState : public GlobalTableFunctionState {
	cubit_gpu_result *result = nullptr;
	vector<column_t> column_ids;
	idx_t row_count = 0, offset = 0;
	uint64_t sum_lo = 0;
	int64_t sum_hi = 0;
	bool agg_emitted = false;
	~CubitScanGlobalState() override {
		cubit_gpu_free_result(result);
	}
	idx_t MaxThreads() const override {
		return 1; // like index_scan (table_scan.cpp:213-225)
	}
};

static unique_ptr<FunctionData> CubitScanBind(ClientContext &, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(input.inputs[0].GetValue<string>());
	bind->lo = input.inputs[1].GetValue<int64_t>();
	bind->hi = input.inputs[2].GetValue<int64_t>();
	for (auto &n : bind->gpu->column_names) {
		return_types.emplace_back(LogicalType::BIGINT);
		names.emplace_back(n);
	}
	return std::move(bind);
}

static unique_ptr<GlobalTableFunctionState> CubitRunQuery(const CubitScanBindData &bind,
                                                          const vector<column_t> &column_ids) {
	auto state = make_uniq<CubitScanGlobalState>();
	state->column_ids = column_ids;
	auto &gpu = *bind.gpu;
	const int64_t lo = MaxValue<int64_t>(bind.lo, gpu.base_value);
	const int64_t hi = MinValue<int64_t>(bind.hi, gpu.base_value + gpu.cardinality - 1);
	if (lo > hi) {
		return std::move(state);
	}
	vector<cubit_bv_ref> refs;
	for (int64_t v = lo; v <= hi; v++) {
		refs.push_back(cubit_bv_ref {gpu.index_id, NumericCast<uint32_t>(v - gpu.base_value)});
	}
	cubit_pred_group group {NumericCast<uint32_t>(refs.size()), refs.data()};
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
	q.n_groups = 1;
	q.groups = &group;
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
	return std::move(state);
}

static unique_ptr<GlobalTableFunctionState> CubitScanInitGlobal(ClientContext &, TableFunctionInitInput &input) {
	return CubitRunQuery(input.bind_data->Cast<CubitScanBindData>(), input.column_ids);
}

static void CubitScanFunction(ClientContext &, TableFunctionInput &data_p, DataChunk &output) {
	auto &state = data_p.global_state->Cast<CubitScanGlobalState>();
	if (state.offset >= state.row_count) {
		return; // chunk.size() == 0 → PhysicalTableScan::GetData returns FINISHED
	}
	const idx_t scan_count = MinValue<idx_t>(STANDARD_VECTOR_SIZE, state.row_count - state.offset);
	int64_t *rowids = nullptr;
	vector<void *> col_ptrs;
	for (idx_t i = 0; i < state.column_ids.size(); i++) {
		auto ptr = FlatVector::GetData<int64_t>(output.data[i]);
		if (state.column_ids[i] == COLUMN_IDENTIFIER_ROW_ID) {
			rowids = ptr;
		} else {
			col_ptrs.push_back(ptr);
		}
	}
	CubitCheck(cubit_gpu_fetch(state.result, state.offset, scan_count, rowids, NumericCast<uint32_t>(col_ptrs.size()),
	                           col_ptrs.data()));
	output.SetCardinality(scan_count);
	state.offset += scan_count;
}

// ---------------------------------------------------------------- cubit_agg(table, lo, hi, column) → (count, sum)
static unique_ptr<FunctionData> CubitAggBind(ClientContext &, TableFunctionBindInput &input,
                                             vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(input.inputs[0].GetValue<string>());
	bind->lo = input.inputs[1].GetValue<int64_t>();
	bind->hi = input.inputs[2].GetValue<int64_t>();
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
	output.SetValue(1, 0, Value::HUGEINT(sum));
	output.SetCardinality(1);
	state.agg_emitted = true;
}

// ---------------------------------------------------------------- registration
void RegisterCubitGpuFunctions(DatabaseInstance &db) {
	TableFunction load("cubit_load", {LogicalType::VARCHAR, LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT},
	                   CubitLoadFunction, CubitLoadBind);
	ExtensionUtil::RegisterFunction(db, load);

	TableFunction scan("cubit_scan", {LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT}, CubitScanFunction,
	                   CubitScanBind, CubitScanInitGlobal);
	scan.projection_pushdown = true; // column_ids tell the GPU which columns to probe
	ExtensionUtil::RegisterFunction(db, scan);

	TableFunction agg("cubit_agg", {LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT, LogicalType::VARCHAR},
	                  CubitAggFunction, CubitAggBind, CubitScanInitGlobal);
	ExtensionUtil::RegisterFunction(db, agg);
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
