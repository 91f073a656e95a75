// cubit_scan.hpp — the table function behind PhysicalTableScan::GetData for the bitmap scan.
//
// Mirrors, name for name, what the reference's scan operator drives:
//   PhysicalTableScan::GetData → function.function(context, TableFunctionInput, DataChunk&)
//        src/execution/operator/scan/physical_table_scan.cpp:82-103
//   bind / init_global / function typedefs        src/include/duckdb/function/table_function.hpp:184-222
//   IndexScanGlobalState / IndexScanFunction      src/function/table/table_scan.cpp:213-273
//        (row ids → DataTable::Fetch, ≤ STANDARD_VECTOR_SIZE rows per call, single threaded:
//         MaxThreads() = 1, table_function.hpp:53-55)
//   predicates arrive as per-column ConstantFilter conjunctions (lo <= col AND col <= hi)
//        src/optimizer/filter_combiner.cpp:438-480, planner/filter/constant_filter.hpp:17-27
// The GPU does merge → decode → probe once in init_global; function() hands the result out
// in DataChunks.  With aggregate push-down the scan returns ONE row (COUNT, SUM lower, SUM upper).
#pragma once
#include "cubit_index.hpp"

namespace cubit_host {

// lo <= indexed column <= hi  — one OR group over the value bitvectors in [lo, hi]
struct CubitPredicate {
	CubitIndex *index;
	int64_t lo;
	int64_t hi;
};

enum class CubitAggregate : uint8_t { NONE = 0, SUM = 1, SUM_PRODUCT = 2 };

struct CubitScanBindData { // FunctionData: immutable during execution (table_function.hpp:236-237)
	CubitTable *table = nullptr;
	std::vector<CubitPredicate> predicates; // AND across entries
	CubitAggregate aggregate = CubitAggregate::NONE;
	column_t agg_column_a = 0;
	column_t agg_column_b = 0;
};

struct CubitScanGlobalState { // GlobalTableFunctionState
	~CubitScanGlobalState();
	idx_t MaxThreads() const {
		return 1;
	}
	cubit_gpu_result *result = nullptr;
	std::vector<column_t> column_ids; // projected columns; COLUMN_IDENTIFIER_ROW_ID = rowid
	std::vector<LogicalTypeId> types;
	idx_t row_count = 0;  // rows selected
	idx_t offset = 0;     // rows already handed out
	hugeint_t sum;        // aggregate push-down result
	bool aggregate_done = false;
	bool finished = false;
	// host staging window: rows [win_begin, win_end) of the result, fetched in one D2H into page-locked buffers
	// (cubit_gpu_alloc_host; kWindowRows * 8 bytes each, allocated on first use, freed with the state)
	idx_t win_begin = 0, win_end = 0;
	row_t *win_rowids = nullptr;
	std::vector<uint8_t *> win_cols; // one slot per column_ids entry (nullptr for the rowid slot)
	// the same window in the narrow wire format (include/cubit_gpu_wire.h): the device writes per-DataChunk frames
	// of base + 1/2/4/8-byte deltas into it and GetData widens one chunk straight into the output vectors.  On a
	// sharded table the shard that holds the window writes it; a window that straddles two shards uses the copies above.
	void *win_wire = nullptr;
	uint64_t win_wire_bytes = 0;
	bool narrow_wire = false; // try the wire first
	bool win_is_wire = false; // the current window arrived as a wire (one that straddles two shards does not)
};

std::unique_ptr<CubitScanBindData> CubitScanBind(CubitTable &table, std::vector<CubitPredicate> predicates,
                                                 CubitAggregate aggregate = CubitAggregate::NONE,
                                                 column_t agg_column_a = 0, column_t agg_column_b = 0);

// runs the query on the GPU (merge + delta XOR → decode → probe); returns the scan state
std::unique_ptr<CubitScanGlobalState> CubitScanInitGlobal(const CubitScanBindData &bind,
                                                         const std::vector<column_t> &column_ids);

// the table function: fills `output` with the next ≤ STANDARD_VECTOR_SIZE rows (0 rows = exhausted)
void CubitScanFunction(const CubitScanBindData &bind, CubitScanGlobalState &gstate, DataChunk &output);

// PhysicalTableScan::GetData: call the function, FINISHED when the chunk comes back empty
SourceResultType CubitScanGetData(const CubitScanBindData &bind, CubitScanGlobalState &gstate, DataChunk &chunk);

std::vector<LogicalTypeId> CubitScanReturnTypes(const CubitScanBindData &bind, const std::vector<column_t> &column_ids);

} // namespace cubit_host
