// duckdb_mirror.hpp — the handful of DuckDB runtime types the scan path touches, restated so
// the host layer compiles without the DuckDB tree (reference headers may not be copied).
// Names, fields and semantics follow the reference:
//   row_t / idx_t / sel_t            src/include/duckdb/common/typedefs.hpp:16,19,30
//   STANDARD_VECTOR_SIZE = 2048      src/include/duckdb/common/vector_size.hpp:16
//   COLUMN_IDENTIFIER_ROW_ID         src/common/constants.cpp:11
//   DataChunk (SetCardinality, size, Reset, data[])   src/include/duckdb/common/types/data_chunk.hpp:43-163
//   Vector (flat: typed array)                        src/include/duckdb/common/types/vector.hpp:242-256
//   exceptions thrown by table functions              src/main/capi/table_function-c.cpp:203-212
// INTEGRATION.md shows the same code against the real headers.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace cubit_host {

using idx_t = uint64_t;
using row_t = int64_t;
using column_t = uint64_t;
constexpr idx_t STANDARD_VECTOR_SIZE = 2048;
constexpr column_t COLUMN_IDENTIFIER_ROW_ID = (column_t)-1;

enum class LogicalTypeId : uint8_t { INTEGER = 4, BIGINT = 8 }; // value = physical width in bytes

struct InvalidInputException : std::runtime_error {
	using std::runtime_error::runtime_error;
};
struct InternalException : std::runtime_error {
	using std::runtime_error::runtime_error;
};

// FLAT vector: a typed array of up to STANDARD_VECTOR_SIZE values plus the ValidityMask (vector.hpp:242-256,
// validity_mask.hpp:50: u64 words, bit r%64 of word r/64 = 1 when row r is valid; "all valid" = no mask)
struct Vector {
	LogicalTypeId type = LogicalTypeId::BIGINT;
	std::vector<uint8_t> buffer;
	uint64_t validity[STANDARD_VECTOR_SIZE / 64];
	bool all_valid = true;
	void Initialize(LogicalTypeId t) {
		type = t;
		buffer.assign((size_t)t * STANDARD_VECTOR_SIZE, 0);
		all_valid = true;
	}
	bool RowIsValid(idx_t i) const { // ValidityMask::RowIsValid (validity_mask.hpp:163-168)
		return all_valid || ((validity[i / 64] >> (i % 64)) & 1);
	}
	template <class T>
	T *GetData() {
		return reinterpret_cast<T *>(buffer.data());
	}
	void *Raw() {
		return buffer.data();
	}
};

struct DataChunk {
	std::vector<Vector> data;
	idx_t count = 0;
	void Initialize(const std::vector<LogicalTypeId> &types) {
		data.resize(types.size());
		for (size_t i = 0; i < types.size(); i++) {
			data[i].Initialize(types[i]);
		}
		count = 0;
	}
	idx_t ColumnCount() const {
		return data.size();
	}
	idx_t size() const {
		return count;
	}
	void SetCardinality(idx_t n) {
		count = n;
	}
	void Reset() {
		count = 0;
	}
};

enum class SourceResultType : uint8_t { HAVE_MORE_OUTPUT, FINISHED }; // operator/physical_operator_states.hpp

// 128-bit SUM result (hugeint_t: lower u64, upper i64 — src/include/duckdb/common/hugeint.hpp)
struct hugeint_t {
	uint64_t lower = 0;
	int64_t upper = 0;
};

} // namespace cubit_host
