#include "cubit_scan.hpp"

#include "cubit_gpu.h"
#include "cubit_gpu_wire.h"

#include <algorithm>

namespace cubit_host {

static constexpr idx_t kWindowRows = 128 * STANDARD_VECTOR_SIZE; // rows per device→host transfer

static inline void Check(int rc) {
	if (rc != CUBIT_OK) {
		ThrowLastError(rc);
	}
}

CubitScanGlobalState::~CubitScanGlobalState() {
	if (result) {
		cubit_gpu_free_result(result);
	}
	cubit_gpu_free_host(win_wire);
	cubit_gpu_free_host(win_rowids);
	for (auto p : win_cols) {
		cubit_gpu_free_host(p);
	}
}

std::unique_ptr<CubitScanBindData> CubitScanBind(CubitTable &table, std::vector<CubitPredicate> predicates,
                                                 CubitAggregate aggregate, column_t agg_a, column_t agg_b) {
	if (predicates.empty()) {
		throw InvalidInputException("cubit_scan needs at least one predicate on an indexed column");
	}
	for (auto &p : predicates) {
		if (!p.index || &p.index->Table() != &table) {
			throw InvalidInputException("cubit_scan predicate refers to an index of another table");
		}
	}
	auto bind = std::make_unique<CubitScanBindData>();
	bind->table = &table;
	bind->predicates = std::move(predicates);
	bind->aggregate = aggregate;
	bind->agg_column_a = agg_a;
	bind->agg_column_b = agg_b;
	return bind;
}

std::vector<LogicalTypeId> CubitScanReturnTypes(const CubitScanBindData &bind, const std::vector<column_t> &column_ids) {
	if (bind.aggregate != CubitAggregate::NONE) {
		return {LogicalTypeId::BIGINT, LogicalTypeId::BIGINT, LogicalTypeId::BIGINT}; // COUNT, SUM lower, SUM upper
	}
	std::vector<LogicalTypeId> types;
	for (auto c : column_ids) {
		types.push_back(bind.table->ColumnType(c));
	}
	return types;
}

std::unique_ptr<CubitScanGlobalState> CubitScanInitGlobal(const CubitScanBindData &bind,
                                                         const std::vector<column_t> &column_ids) {
	auto state = std::make_unique<CubitScanGlobalState>();
	state->column_ids = column_ids;
	state->types = CubitScanReturnTypes(bind, column_ids);

	// predicate → AND of OR groups over value bitvectors
	std::vector<std::vector<cubit_bv_ref>> refs(bind.predicates.size());
	std::vector<cubit_pred_group> groups;
	bool empty_result = false;
	for (size_t g = 0; g < bind.predicates.size(); g++) {
		const auto &p = bind.predicates[g];
		const int64_t base = p.index->BaseValue(), card = p.index->Cardinality();
		const int64_t lo = std::max(p.lo, base), hi = std::min(p.hi, base + card - 1);
		if (lo > hi) {
			empty_result = true;
			break;
		}
		for (int64_t v = lo; v <= hi; v++) {
			refs[g].push_back(cubit_bv_ref {p.index->Id(), (uint32_t)(v - base)});
		}
		groups.push_back(cubit_pred_group {(uint32_t)refs[g].size(), refs[g].data()});
	}
	if (empty_result) {
		state->row_count = 0;
		return state;
	}
	std::vector<int32_t> cols;
	bool want_rowid = false;
	for (auto c : column_ids) {
		if (c == COLUMN_IDENTIFIER_ROW_ID) {
			want_rowid = true;
		} else {
			cols.push_back((int32_t)c);
		}
	}
	cubit_query q {};
	q.n_groups = (uint32_t)groups.size();
	q.groups = groups.data();
	if (bind.aggregate != CubitAggregate::NONE) {
		q.flags = 0; // aggregate push-down: no row IDs are materialised at all
		q.agg_kind = (int32_t)bind.aggregate;
		q.agg_col_a = (int32_t)bind.agg_column_a;
		q.agg_col_b = (int32_t)bind.agg_column_b;
	} else {
		q.flags = (want_rowid ? CUBIT_Q_ROWIDS : 0u) | (cols.empty() ? 0u : CUBIT_Q_VALUES);
		q.n_cols = (uint32_t)cols.size();
		q.cols = cols.data();
	}
	Check(cubit_gpu_query(bind.table->Handle(), &q, &state->result));
	cubit_result_info info;
	Check(cubit_gpu_result_get(state->result, &info));
	state->row_count = info.count;
	{
		uint32_t n_shards = 1;
		cubit_gpu_shard_count(bind.table->Handle(), &n_shards);
		(void)n_shards;
		state->narrow_wire = bind.aggregate == CubitAggregate::NONE && !column_ids.empty();
	}
	state->sum.lower = info.sum_lo;
	state->sum.upper = info.sum_hi;
	return state;
}

static void FillWindow(CubitScanGlobalState &st) {
	const idx_t n = std::min(kWindowRows, st.row_count - st.offset);
	st.win_begin = st.offset;
	st.win_end = st.offset + n;
	bool want_rowid = false;
	std::vector<void *> ptrs;
	if (st.narrow_wire) {
		uint32_t n_value_cols = 0;
		for (auto c : st.column_ids) {
			want_rowid = want_rowid || c == COLUMN_IDENTIFIER_ROW_ID;
			n_value_cols += c == COLUMN_IDENTIFIER_ROW_ID ? 0 : 1;
		}
		if (!st.win_wire) {
			st.win_wire_bytes = cubit_wire_bytes(kWindowRows, (want_rowid ? 1u : 0u) + n_value_cols);
			Check(cubit_gpu_alloc_host(st.win_wire_bytes, &st.win_wire));
		}
		cubit_gpu_fetch_ticket *ticket = nullptr;
		const int rc = cubit_gpu_fetch_wire_async(st.result, st.win_begin, n, want_rowid ? 1 : 0, n_value_cols, st.win_wire,
		                                          st.win_wire_bytes, &ticket);
		if (rc == CUBIT_OK) {
			Check(cubit_gpu_fetch_wait(ticket));
			st.win_is_wire = true;
			return;
		}
		if (rc != CUBIT_ESTATE) { // ESTATE: the window straddles two shards of a sharded table → 8-byte copies
			Check(rc);
		}
		want_rowid = false;
	}
	st.win_is_wire = false;
	st.win_cols.resize(st.column_ids.size(), nullptr);
	auto pinned = [](void *&p) {
		if (!p) {
			Check(cubit_gpu_alloc_host(kWindowRows * 8, &p));
		}
	};
	for (size_t i = 0; i < st.column_ids.size(); i++) {
		if (st.column_ids[i] == COLUMN_IDENTIFIER_ROW_ID) {
			want_rowid = true;
		} else {
			void *p = st.win_cols[i];
			pinned(p);
			st.win_cols[i] = static_cast<uint8_t *>(p);
			ptrs.push_back(p);
		}
	}
	if (want_rowid) {
		void *p = st.win_rowids;
		pinned(p);
		st.win_rowids = static_cast<row_t *>(p);
	}
	Check(cubit_gpu_fetch(st.result, st.win_begin, n, want_rowid ? st.win_rowids : nullptr, (uint32_t)ptrs.size(),
	                      ptrs.data()));
}

void CubitScanFunction(const CubitScanBindData &bind, CubitScanGlobalState &st, DataChunk &output) {
	output.Reset();
	if (st.finished) {
		return;
	}
	if (bind.aggregate != CubitAggregate::NONE) {
		if (!st.aggregate_done) {
			if (output.ColumnCount() != 3) {
				throw InternalException("aggregate push-down returns (COUNT, SUM lower, SUM upper)");
			}
			output.data[0].GetData<int64_t>()[0] = (int64_t)st.row_count;
			output.data[1].GetData<int64_t>()[0] = (int64_t)st.sum.lower;
			output.data[2].GetData<int64_t>()[0] = st.sum.upper;
			output.SetCardinality(1);
			st.aggregate_done = true;
		} else {
			st.finished = true;
		}
		return;
	}
	if (st.offset >= st.row_count) {
		st.finished = true;
		return;
	}
	if (output.ColumnCount() != st.column_ids.size()) {
		throw InternalException("output chunk does not match the projected column list");
	}
	if (st.offset >= st.win_end) {
		FillWindow(st);
	}
	// IndexScanFunction: scan_count = min(STANDARD_VECTOR_SIZE, remaining)  (table_scan.cpp:258-261)
	const idx_t scan_count = std::min<idx_t>(STANDARD_VECTOR_SIZE, st.win_end - st.offset);
	const idx_t rel = st.offset - st.win_begin;
	uint32_t value_col = 0;
	bool has_rowid = false;
	for (auto c : st.column_ids) {
		has_rowid = has_rowid || c == COLUMN_IDENTIFIER_ROW_ID;
	}
	auto unpack = [&](uint32_t stream, void *dst, uint32_t elem) {
		if (cubit_wire_unpack_chunk(st.win_wire, stream, rel / STANDARD_VECTOR_SIZE, dst, elem) != (int)scan_count) {
			throw InternalException("malformed wire window");
		}
	};
	for (size_t i = 0; i < st.column_ids.size(); i++) {
		output.data[i].all_valid = true;
		if (st.column_ids[i] == COLUMN_IDENTIFIER_ROW_ID) {
			if (st.win_is_wire) {
				unpack(0, output.data[i].Raw(), 8);
			} else {
				memcpy(output.data[i].Raw(), st.win_rowids + rel, scan_count * sizeof(row_t));
			}
		} else {
			const size_t w = (size_t)st.types[i];
			if (st.win_is_wire) {
				unpack((has_rowid ? 1u : 0u) + value_col, output.data[i].Raw(), (uint32_t)w);
			} else {
				memcpy(output.data[i].Raw(), st.win_cols[i] + rel * w, scan_count * w);
			}
			// validity of the probed values (StandardColumnData::FetchRow: validity.FetchRow + data)
			int all = 1;
			Check(cubit_gpu_fetch_validity(st.result, value_col++, st.offset, scan_count, output.data[i].validity, &all));
			output.data[i].all_valid = all != 0;
		}
	}
	output.SetCardinality(scan_count);
	st.offset += scan_count;
}

SourceResultType CubitScanGetData(const CubitScanBindData &bind, CubitScanGlobalState &gstate, DataChunk &chunk) {
	CubitScanFunction(bind, gstate, chunk);
	return chunk.size() == 0 ? SourceResultType::FINISHED : SourceResultType::HAVE_MORE_OUTPUT;
}

} // namespace cubit_host
