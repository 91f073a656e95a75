#include "cubit_index.hpp"

#include "cubit_gpu.h"

namespace cubit_host {

void ThrowLastError(int rc) {
	std::string msg = cubit_gpu_last_error();
	if (rc == CUBIT_EINVAL) {
		throw InvalidInputException(msg);
	}
	throw InternalException("cubit_gpu: " + msg);
}

static inline void Check(int rc) {
	if (rc != CUBIT_OK) {
		ThrowLastError(rc);
	}
}

CubitTable::CubitTable(idx_t n_rows_p, row_t row_base_p, uint32_t seg_bits, int device)
    : n_rows(n_rows_p), row_base(row_base_p) {
	Check(cubit_gpu_create(device, n_rows, row_base, seg_bits, &handle));
}

CubitTable::~CubitTable() {
	cubit_gpu_destroy(handle);
}

void CubitTable::AddColumn(column_t col, const int64_t *values) {
	Check(cubit_gpu_upload_column(handle, (int32_t)col, values, 8, n_rows));
	col_types[col] = LogicalTypeId::BIGINT;
}

void CubitTable::AddColumn(column_t col, const int32_t *values) {
	Check(cubit_gpu_upload_column(handle, (int32_t)col, values, 4, n_rows));
	col_types[col] = LogicalTypeId::INTEGER;
}

void CubitTable::SetValidity(column_t col, const uint64_t *words) {
	if (!col_types.count(col)) {
		throw InvalidInputException("Table does not have column " + std::to_string(col));
	}
	Check(cubit_gpu_upload_column_validity(handle, (int32_t)col, words, words ? (n_rows + 63) / 64 : 0));
}

void CubitTable::Append(idx_t n_new, const std::map<column_t, const void *> &values) {
	std::vector<cubit_append_column> cols;
	for (auto &kv : values) {
		auto it = col_types.find(kv.first);
		if (it == col_types.end()) {
			throw InvalidInputException("Table does not have column " + std::to_string(kv.first));
		}
		cols.push_back(cubit_append_column {(int32_t)kv.first, it->second == LogicalTypeId::BIGINT ? 8u : 4u, kv.second});
	}
	Check(cubit_gpu_append_rows(handle, n_new, cols.data(), (uint32_t)cols.size()));
	n_rows += n_new;
}

LogicalTypeId CubitTable::ColumnType(column_t col) const {
	if (col == COLUMN_IDENTIFIER_ROW_ID) {
		return LogicalTypeId::BIGINT;
	}
	auto it = col_types.find(col);
	if (it == col_types.end()) {
		throw InvalidInputException("Table does not have column " + std::to_string(col));
	}
	return it->second;
}

CubitIndex::CubitIndex(CubitTable &table_p, column_t column_p, int64_t base_value_p, uint32_t cardinality_p)
    : table(table_p), column(column_p), base_value(base_value_p), cardinality(cardinality_p) {
	Check(cubit_gpu_index_create(table.Handle(), cardinality, &index_id));
}

uint32_t CubitIndex::ValueId(int64_t v) const {
	if (v < base_value || v >= base_value + (int64_t)cardinality) {
		throw InvalidInputException("key " + std::to_string(v) + " outside the indexed domain");
	}
	return (uint32_t)(v - base_value);
}

void CubitIndex::Build() {
	Check(cubit_gpu_index_build(table.Handle(), index_id, (int32_t)column, base_value));
}

void CubitIndex::Delete(row_t row_id, int64_t current_value) {
	std::lock_guard<std::mutex> lk(mu);
	staged_values.push_back(ValueId(current_value));
	staged_rows.push_back(row_id - table.RowBase());
}

void CubitIndex::Update(row_t row_id, int64_t old_value, int64_t new_value) {
	if (old_value == new_value) {
		return;
	}
	std::lock_guard<std::mutex> lk(mu);
	const uint32_t a = ValueId(old_value), b = ValueId(new_value);
	staged_values.push_back(a);
	staged_rows.push_back(row_id - table.RowBase());
	staged_values.push_back(b);
	staged_rows.push_back(row_id - table.RowBase());
}

// INCREMENTAL: only the flips staged since the last commit cross to the GPU, where they are merged into the
// index's pending-delta lists on the device (cubit_gpu_add_delta_pairs); earlier pending flips stay as they are
void CubitIndex::CommitDeltas() {
	std::lock_guard<std::mutex> lk(mu);
	if (staged_rows.empty()) {
		return;
	}
	Check(cubit_gpu_add_delta_pairs(table.Handle(), index_id, staged_values.data(), staged_rows.data(), staged_rows.size()));
	committed += staged_rows.size();
	staged_values.clear();
	staged_rows.clear();
}

void CubitIndex::MergeDeltas() {
	CommitDeltas();
	std::lock_guard<std::mutex> lk(mu);
	Check(cubit_gpu_merge_deltas(table.Handle(), index_id));
	committed = 0;
}

idx_t CubitIndex::PendingDeltaRows() const {
	std::lock_guard<std::mutex> lk(mu);
	// (the library may have folded committed flips back on its own: the merge-back threshold)
	cubit_index_info info;
	if (cubit_gpu_index_info(table.Handle(), index_id, &info) == CUBIT_OK) {
		return staged_rows.size() + info.delta_entries;
	}
	return staged_rows.size() + committed;
}

bool CubitIndex::Scan(int64_t lo, int64_t hi, idx_t max_count, std::vector<row_t> &row_ids) {
	lo = lo < base_value ? base_value : lo;
	hi = hi >= base_value + (int64_t)cardinality ? base_value + (int64_t)cardinality - 1 : hi;
	if (lo > hi) {
		row_ids.clear();
		return true;
	}
	std::vector<cubit_bv_ref> refs;
	for (int64_t v = lo; v <= hi; v++) {
		refs.push_back(cubit_bv_ref {index_id, (uint32_t)(v - base_value)});
	}
	if (refs.size() > CUBIT_MAX_STREAMS) {
		throw InvalidInputException("range spans more than CUBIT_MAX_STREAMS value bitvectors");
	}
	cubit_pred_group grp {(uint32_t)refs.size(), refs.data()};
	cubit_query q {};
	q.n_groups = 1;
	q.groups = &grp;
	q.flags = CUBIT_Q_ROWIDS;
	cubit_gpu_result *res = nullptr;
	Check(cubit_gpu_query(table.Handle(), &q, &res));
	cubit_result_info info;
	int rc = cubit_gpu_result_get(res, &info);
	if (rc == CUBIT_OK && info.count > max_count) {
		cubit_gpu_free_result(res);
		return false; // ART::Scan gives up past max_count (art.cpp:955-972)
	}
	if (rc == CUBIT_OK) {
		row_ids.resize(info.count);
		rc = cubit_gpu_fetch(res, 0, info.count, row_ids.data(), 0, nullptr);
	}
	cubit_gpu_free_result(res);
	Check(rc);
	return true;
}

} // namespace cubit_host
