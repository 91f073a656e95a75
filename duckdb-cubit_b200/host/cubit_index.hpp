// cubit_index.hpp — host-side CUBIT index and table objects over the C-ABI (include/cubit_gpu.h).
//
// CubitIndex mirrors the surface a DuckDB index exposes on this path:
//   BoundIndex::Append / Delete / (update = delete+insert) / MergeIndexes
//        src/include/duckdb/execution/index/bound_index.hpp:67-126
//   ART::Scan(..., max_count, row_ids) → false when more than max_count rows match,
//        otherwise SORTED, duplicate-free row ids
//        src/execution/index/art/art.cpp:918-986
// Deletes/updates are recorded as pending bit flips (UpBit/CUBIT deltas) and XOR-ed at query
// time on the GPU; MergeDeltas() folds them back into the value bitvectors.
// All errors are C++ exceptions, as in the reference's operator code.
#pragma once
#include "duckdb_mirror.hpp"

#include <map>
#include <memory>
#include <mutex>

struct cubit_gpu_table;
struct cubit_gpu_result;

namespace cubit_host {

class CubitTable {
public:
	CubitTable(idx_t n_rows, row_t row_base = 0, uint32_t seg_bits = 65536, int device = 0);
	~CubitTable();
	CubitTable(const CubitTable &) = delete;
	CubitTable &operator=(const CubitTable &) = delete;

	void AddColumn(column_t col, const int64_t *values);
	void AddColumn(column_t col, const int32_t *values);
	// NULLs of a resident column: its ValidityMask words (ceil(rows / 64); nullptr = no NULLs)
	void SetValidity(column_t col, const uint64_t *words);
	// INSERT: n_new rows appended at the end (they take the next row ids); `values` holds one array of n_new
	// elements per resident column.  Indexes built from a column are extended on the GPU
	// (BoundIndex::Append, bound_index.hpp:71-75).
	void Append(idx_t n_new, const std::map<column_t, const void *> &values);
	LogicalTypeId ColumnType(column_t col) const;
	idx_t RowCount() const {
		return n_rows;
	}
	row_t RowBase() const {
		return row_base;
	}
	cubit_gpu_table *Handle() const {
		return handle;
	}

private:
	cubit_gpu_table *handle = nullptr;
	idx_t n_rows;
	row_t row_base;
	std::map<column_t, LogicalTypeId> col_types;
};

class CubitIndex {
public:
	// index over `column` whose values lie in [base_value, base_value + cardinality)
	CubitIndex(CubitTable &table, column_t column, int64_t base_value, uint32_t cardinality);

	void Build();                                               // CREATE INDEX: bitvectors from the column (GPU)
	void Delete(row_t row_id, int64_t current_value);           // BoundIndex::Delete
	void Update(row_t row_id, int64_t old_value, int64_t new_value); // delete + insert of the key
	void CommitDeltas();                                        // publish pending flips to the GPU (XOR at query time)
	void MergeDeltas();                                         // merge-back: B ^= D (BoundIndex::MergeIndexes analog)
	idx_t PendingDeltaRows() const;

	// ART::Scan analog for lo <= key <= hi.  Returns false (and leaves row_ids untouched) when
	// more than max_count rows match.
	bool Scan(int64_t lo, int64_t hi, idx_t max_count, std::vector<row_t> &row_ids);

	int32_t Id() const {
		return index_id;
	}
	int64_t BaseValue() const {
		return base_value;
	}
	uint32_t Cardinality() const {
		return cardinality;
	}
	column_t Column() const {
		return column;
	}
	CubitTable &Table() const {
		return table;
	}

private:
	uint32_t ValueId(int64_t v) const;
	CubitTable &table;
	column_t column;
	int64_t base_value;
	uint32_t cardinality;
	int32_t index_id = -1;
	std::vector<uint32_t> staged_values; // (value id, local row) flips recorded since the last CommitDeltas
	std::vector<row_t> staged_rows;
	idx_t committed = 0;                 // flips handed to the GPU and still pending there
	mutable std::mutex mu;
};

[[noreturn]] void ThrowLastError(int rc);

} // namespace cubit_host
