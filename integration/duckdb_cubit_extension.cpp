// duckdb_cubit_extension.cpp — the DuckDB-side binding of the B200 CUBIT scan (reference glue).
//
// This file is compiled against the REAL reference headers (/root/reference/src/include) — it is the
// code a DuckDB maintainer adds to make the GPU path a drop-in behind PhysicalTableScan::GetData.
// It only talks to the GPU library through include/cubit_gpu.h (extern "C").  It is syntax-checked
// against the reference headers by __graft_entry__.build() whenever /root/reference is present and is
// reproduced, annotated, in INTEGRATION.md.  Nothing here is copied from the reference; it USES its
// public extension API:
//   ExtensionUtil::RegisterFunction          src/include/duckdb/main/extension_util.hpp:34
//   TableFunction{bind, init_global, init_local, function, get_batch_index}
//                                            src/include/duckdb/function/table_function.hpp:184-301
//   IndexTypeSet::RegisterIndexType          src/execution/index/index_type_set.cpp:24-30
//   BoundIndex (Append / Delete / Insert / CommitDrop / GetStorageInfo)
//                                            src/include/duckdb/execution/index/bound_index.hpp:67-126
//   OptimizerExtension                       src/include/duckdb/optimizer/optimizer_extension.hpp:31-43
//
// The GPU-resident mirror of a table is a registered INDEX TYPE ("CUBIT"): CALL cubit_load(...) creates an index
// catalog entry of that type on the table and puts a CubitIndex (a BoundIndex) into the table's own index list, next
// to its ART indexes.  From then on DuckDB's DML path maintains it like any index —
//   INSERT            → DataTable::AppendToIndexes → CubitIndex::Append  → cubit_gpu_append_rows
//   DELETE            → cleanup of the committed delete → DataTable::RemoveFromIndexes → CubitIndex::Delete
//                       → cubit_gpu_add_delta_pairs (the rows' bits are flipped in the pending deltas D_v)
//   UPDATE of a resident column → the binder turns it into DELETE + INSERT because an index covers the column
//                       (BoundIndex::IndexIsUpdated) → both of the above
//   DROP TABLE / DROP INDEX → CommitDrop → the GPU memory is released
//   CHECKPOINT        → GetStorageInfo → the index images (cubit_gpu_index_serialize) are written to blocks
//   ATTACH / restart  → IndexType::create_instance → the images are read back and handed to
//                       cubit_gpu_index_deserialize when the table is first used
// — and there is no registry keyed by table name: the accelerator is found through the table's own index list, so
// dropping, re-creating, altering or shadowing a table can never route a query to another table's rows.
//
// SQL surface:
//   CALL cubit_load('lineitem', 'l_quantity', 1, 50);      -- build the GPU index + upload the GPU-eligible columns
//   SELECT sum(l_extendedprice) FROM cubit_scan('lineitem', 24, 24);   -- rowid + the uploaded columns
//   SELECT * FROM cubit_agg('lineitem', 24, 24, 'l_extendedprice');    -- aggregate push-down: one row
// and, transparently (OptimizerExtension): a plain
//   SELECT sum(l_extendedprice) FROM lineitem WHERE l_quantity BETWEEN 10 AND 19;
// whose pushed-down filters touch only indexed columns is re-pointed at the GPU scan, the way
// TableScanPushdownComplexFilter re-points a seq_scan at ART's index_scan (src/function/table/table_scan.cpp:296-370).
#include "duckdb.hpp"
#include <chrono>
#include <exception>
#include <thread>
#include "duckdb/catalog/catalog_entry/duck_index_entry.hpp"
#include "duckdb/catalog/catalog_entry/duck_table_entry.hpp"
#include "duckdb/catalog/catalog_entry/table_catalog_entry.hpp"
#include "duckdb/common/types/data_chunk.hpp"
#include "duckdb/common/types/date.hpp"
#include "duckdb/common/types/vector.hpp"
#include "duckdb/execution/index/bound_index.hpp"
#include "duckdb/execution/index/index_type.hpp"
#include "duckdb/execution/index/index_type_set.hpp"
#include "duckdb/function/table_function.hpp"
#include "duckdb/main/config.hpp"
#include "duckdb/main/extension_util.hpp"
#include "duckdb/optimizer/optimizer_extension.hpp"
#include "duckdb/parser/expression/columnref_expression.hpp"
#include "duckdb/parser/parsed_data/create_index_info.hpp"
#include "duckdb/planner/expression/bound_aggregate_expression.hpp"
#include "duckdb/planner/expression/bound_cast_expression.hpp"
#include "duckdb/planner/expression/bound_columnref_expression.hpp"
#include "duckdb/planner/expression/bound_function_expression.hpp"
#include "duckdb/planner/filter/conjunction_filter.hpp"
#include "duckdb/planner/filter/constant_filter.hpp"
#include "duckdb/planner/operator/logical_aggregate.hpp"
#include "duckdb/planner/operator/logical_get.hpp"
#include "duckdb/storage/block_manager.hpp"
#include "duckdb/storage/buffer_manager.hpp"
#include "duckdb/storage/data_table.hpp"
#include "duckdb/storage/table/data_table_info.hpp"
#include "duckdb/storage/table_io_manager.hpp"
#include "duckdb/storage/table_storage_info.hpp"
#include "duckdb/parser/qualified_name.hpp"
#include "duckdb/transaction/duck_transaction.hpp"
#include "duckdb/transaction/meta_transaction.hpp"

#include "cubit_gpu.h"
#include "cubit_gpu_wire.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace duckdb {

static constexpr const char *CUBIT_INDEX_TYPE = "CUBIT";

static void CubitCheck(int rc) {
	if (rc != CUBIT_OK) {
		// errors cross the C-ABI as codes; here they become the exceptions DuckDB expects
		throw InvalidInputException("cubit_gpu: %s", cubit_gpu_last_error());
	}
}

// ---------------------------------------------------------------- the GPU-resident mirror of one table
// one CUBIT index of a GPU-resident table: value v of `key_table_column` ↔ bitvector v - base_value
struct CubitGpuIndex {
	int32_t index_id = -1;
	int64_t base_value = 0;
	uint32_t cardinality = 0;
	idx_t key_table_column = 0; // table column index of the indexed column
	idx_t key_gpu_column = 0;   // ... and its GPU column id
	// binned index (cubit_load(..., bin := 'month' | '<width>')): bitvector b covers the key values in
	// [bin_lo[b], bin_lo[b + 1]) (cardinality + 1 boundaries); a range predicate is answered by it only when both
	// of its ends fall on boundaries.  Empty = one bitvector per value.
	vector<int64_t> bin_lo;
	// bitvector of raw key `v`, or -1 when the index does not cover it (NULL keys: INT64_MIN)
	int64_t ValueId(int64_t v) const {
		if (bin_lo.empty()) {
			return v >= base_value && v < base_value + NumericCast<int64_t>(cardinality) ? v - base_value : -1;
		}
		if (v < bin_lo.front() || v >= bin_lo.back()) {
			return -1;
		}
		return (std::upper_bound(bin_lo.begin(), bin_lo.end(), v) - bin_lo.begin()) - 1;
	}
};

// what CALL cubit_load asked for: kept in the index catalog entry's options, so it survives a restart
struct CubitIndexSpec {
	string key;
	int64_t base = 0;
	uint32_t cardinality = 0;
	string bin; // "" = one bitvector per value; "month" (DATE keys); or a positive integer bin width
};

static string CubitEncodeSpecs(const vector<CubitIndexSpec> &specs) {
	string out;
	for (auto &s : specs) {
		out += s.key + "\x1f" + std::to_string(s.base) + "\x1f" + std::to_string(s.cardinality) + "\x1f" + s.bin + "\x1e";
	}
	return out;
}

static vector<CubitIndexSpec> CubitDecodeSpecs(const string &text) {
	vector<CubitIndexSpec> out;
	for (auto &rec : StringUtil::Split(text, '\x1e')) {
		if (rec.empty()) {
			continue;
		}
		vector<string> f;
		idx_t at = 0;
		while (true) { // (StringUtil::Split drops empty fields: the last one — the bin — is often empty)
			auto next = rec.find('\x1f', at);
			f.push_back(rec.substr(at, next == string::npos ? string::npos : next - at));
			if (next == string::npos) {
				break;
			}
			at = next + 1;
		}
		if (f.size() != 4) {
			throw InvalidInputException("cubit: malformed index options");
		}
		CubitIndexSpec s;
		s.key = f[0];
		s.base = std::stoll(f[1]);
		s.cardinality = NumericCast<uint32_t>(std::stoll(f[2]));
		s.bin = f[3];
		out.push_back(s);
	}
	return out;
}

struct CubitGpuTable {
	cubit_gpu_table *handle = nullptr;
	vector<CubitGpuIndex> indexes; // [0] = the one cubit_scan(table, lo, hi) / cubit_agg address; CALL cubit_load again
	                               // with another key column to add more (conjunctions across them are rewritten)
	vector<string> column_names; // uploaded columns (raw int64 on the GPU), GPU column id = position
	vector<idx_t> table_column;  // table column index of every uploaded column
	vector<LogicalType> column_types;
	idx_t row_count = 0; // rows resident on the GPU = the table's row-id high-water mark at that point
	bool loaded = false;
	// false once DML produced something the GPU copy cannot represent (a key outside the indexed domain, row ids that
	// do not continue the table): scans then keep the vanilla path until the next cubit_load
	bool usable = true;
	string unusable_reason;
	std::mutex dml_lock; // DML callbacks and (re)loads
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
	// narrow-wire windows (their size depends on the number of streams)
	vector<std::pair<void *, uint64_t>> wire_pool;
	void *AcquireWire(uint64_t bytes) {
		{
			std::lock_guard<std::mutex> lk(pool_lock);
			for (idx_t i = 0; i < wire_pool.size(); i++) {
				if (wire_pool[i].second >= bytes) {
					void *p = wire_pool[i].first;
					wire_pool[i] = wire_pool.back();
					wire_pool.pop_back();
					return p;
				}
			}
		}
		void *p = nullptr;
		if (cubit_gpu_alloc_host(bytes, &p) != CUBIT_OK) {
			throw InvalidInputException("cubit_gpu: %s", cubit_gpu_last_error());
		}
		return p;
	}
	void ReleaseWire(void *p, uint64_t bytes) {
		if (p) {
			std::lock_guard<std::mutex> lk(pool_lock);
			wire_pool.emplace_back(p, bytes);
		}
	}
	void Unload() {
		for (auto p : host_pool) {
			cubit_gpu_free_host(p);
		}
		host_pool.clear();
		for (auto &w : wire_pool) {
			cubit_gpu_free_host(w.first);
		}
		wire_pool.clear();
		cubit_gpu_destroy(handle);
		handle = nullptr;
		indexes.clear();
		loaded = false;
		row_count = 0;
	}
	~CubitGpuTable() {
		Unload();
	}
};

static std::atomic<idx_t> cubit_dml_append_rows {0}, cubit_dml_delta_pairs {0}, cubit_image_loads {0};
idx_t CubitImageLoads() {
	return cubit_image_loads.load();
}
idx_t CubitDmlAppendedRows() {
	return cubit_dml_append_rows.load();
}
idx_t CubitDmlDeltaPairs() {
	return cubit_dml_delta_pairs.load();
}

// raw int64 of every row of `vec` as the GPU holds it: integers widened, DECIMALs as their unscaled value whatever
// their physical width (DECIMAL(15,2) is int64 cents, dbgen.cpp:48-50; DECIMAL(4,2) is an int16 — a numeric cast
// would ROUND 0.05 to 0), DATEs as days since 1970-01-01; NULL rows carry INT64_MIN (outside every indexed domain)
static void CubitRawInt64(Vector &vec, idx_t count, int64_t *out, uint64_t *valid_words,
                          idx_t valid_bit0, bool &any_null) {
	Vector as_bigint(LogicalType::BIGINT);
	const auto &type = vec.GetType();
	if (type.id() == LogicalTypeId::DECIMAL) {
		switch (type.InternalType()) {
		case PhysicalType::INT64:
			as_bigint.Reinterpret(vec);
			break;
		case PhysicalType::INT32: {
			Vector as_int(LogicalType::INTEGER);
			as_int.Reinterpret(vec);
			VectorOperations::DefaultCast(as_int, as_bigint, count);
			break;
		}
		case PhysicalType::INT16: {
			Vector as_small(LogicalType::SMALLINT);
			as_small.Reinterpret(vec);
			VectorOperations::DefaultCast(as_small, as_bigint, count);
			break;
		}
		default:
			throw InternalException("cubit: DECIMAL wider than 18 digits is not GPU eligible");
		}
	} else if (type.id() == LogicalTypeId::DATE) {
		Vector as_int(LogicalType::INTEGER); // days since 1970-01-01 (date_t)
		as_int.Reinterpret(vec);
		VectorOperations::DefaultCast(as_int, as_bigint, count);
	} else {
		VectorOperations::DefaultCast(vec, as_bigint, count);
	}
	as_bigint.Flatten(count);
	auto ptr = FlatVector::GetData<int64_t>(as_bigint);
	auto &mask = FlatVector::Validity(as_bigint);
	for (idx_t r = 0; r < count; r++) {
		if (mask.RowIsValid(r)) {
			out[r] = ptr[r];
		} else {
			out[r] = NumericLimits<int64_t>::Minimum();
			any_null = true;
			if (valid_words) {
				const idx_t bit = valid_bit0 + r;
				valid_words[bit / 64] &= ~(uint64_t(1) << (bit % 64));
			}
		}
	}
}

static bool CubitEligibleType(const LogicalType &t) {
	if (t.id() == LogicalTypeId::DECIMAL) {
		return t.InternalType() == PhysicalType::INT16 || t.InternalType() == PhysicalType::INT32 ||
		       t.InternalType() == PhysicalType::INT64;
	}
	if (t.id() == LogicalTypeId::DATE) {
		return true;
	}
	return t.IsIntegral() && t.InternalType() != PhysicalType::INT128 && t.InternalType() != PhysicalType::UINT128 &&
	       t.InternalType() != PhysicalType::UINT64;
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

static bool CubitUploadColumnSegments(ClientContext &, TableCatalogEntry &entry, idx_t table_column,
                                      const LogicalType &type, cubit_gpu_table *handle, int32_t gpu_col, idx_t row_count) {
	if (type.InternalType() != PhysicalType::INT64) {
		return false;
	}
	uint32_t n_shards = 1;
	cubit_gpu_shard_count(handle, &n_shards);
	if (n_shards > 1) {
		return false; // a compressed segment addresses one shard: a table cut over several GPUs takes the decoded rows
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

// host-side loops over a whole table (placing rows, domain checks, bin ids) on up to 16 threads; fn(begin, end) or,
// with n == 0, fn(thread, n_threads); the first exception is rethrown on the caller's thread
static idx_t CubitHostThreads() {
	return MinValue<idx_t>(MaxValue<idx_t>(1, std::thread::hardware_concurrency()), 16);
}
template <class F>
static void CubitParallelFor(idx_t n, F fn) {
	const idx_t nt = n < (idx_t(1) << 16) ? 1 : CubitHostThreads();
	if (nt == 1) {
		fn(idx_t(0), n);
		return;
	}
	std::exception_ptr err;
	std::mutex mu;
	vector<std::thread> th;
	for (idx_t i = 0; i < nt; i++) {
		th.emplace_back([&, i]() {
			try {
				fn(n * i / nt, n * (i + 1) / nt);
			} catch (...) {
				std::lock_guard<std::mutex> lk(mu);
				if (!err) {
					err = std::current_exception();
				}
			}
		});
	}
	for (auto &t : th) {
		t.join();
	}
	if (err) {
		std::rethrow_exception(err);
	}
}

// Build one index over GPU column `gcol` of a resident table.  The key values come back from the GPU (the
// uploaded raw int64, NULL rows = INT64_MIN): every non-NULL key must fall into the indexed domain — a key the
// index does not cover would silently drop rows from rewritten scans — and binned indexes need the bin id of
// every row (computed here, uploaded as a temporary column, indexed on the GPU, dropped).
static CubitGpuIndex CubitBuildIndex(CubitGpuTable &gpu, idx_t gcol, const CubitIndexSpec &spec, const LogicalType &key_type,
                                     const vector<uint8_t> *image) {
	CubitGpuIndex ix;
	ix.key_table_column = gpu.table_column[gcol];
	ix.key_gpu_column = gcol;
	const int64_t null_key = NumericLimits<int64_t>::Minimum();
	// ---- the domain (bin boundaries) is a function of the spec alone, except for month bins, which follow the data
	vector<int64_t> vals;
	auto fetch_vals = [&]() {
		if (vals.empty() && gpu.row_count) {
			vals.resize(gpu.row_count);
			CubitCheck(cubit_gpu_download_column(gpu.handle, NumericCast<int32_t>(gcol), vals.data(), 8, gpu.row_count));
		}
	};
	if (spec.bin.empty()) {
		ix.base_value = spec.base;
		ix.cardinality = spec.cardinality;
	} else if (spec.bin == "month") {
		if (key_type.id() != LogicalTypeId::DATE) {
			throw InvalidInputException("cubit_load: bin := 'month' needs a DATE key column");
		}
		fetch_vals();
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
		ix.base_value = 0;
		ix.cardinality = NumericCast<uint32_t>(ix.bin_lo.size() - 1);
	} else {
		int64_t width = 0;
		try {
			width = std::stoll(spec.bin);
		} catch (...) {
			width = 0;
		}
		if (width <= 0 || spec.cardinality == 0) {
			throw InvalidInputException("cubit_load: bin must be 'month' or a positive integer width (with base and cardinality)");
		}
		for (uint32_t b = 0; b <= spec.cardinality; b++) {
			ix.bin_lo.push_back(spec.base + NumericCast<int64_t>(b) * width);
		}
		ix.base_value = 0;
		ix.cardinality = spec.cardinality;
	}
	// ---- a checkpointed image of this index (GetStorageInfo) replaces the build when it describes this table
	if (image && !image->empty()) {
		if (cubit_gpu_index_deserialize(gpu.handle, image->data(), image->size(), &ix.index_id) == CUBIT_OK) {
			cubit_image_loads++;
			return ix;
		}
		ix.index_id = -1; // stale image (the table changed since the checkpoint): rebuild from the column
	}
	fetch_vals();
	// one pass on all host threads: the domain check and — binned indexes — the bin id of every row
	const bool binned = !ix.bin_lo.empty();
	unique_ptr<int64_t[]> bins(binned ? new int64_t[vals.size()] : nullptr);
	CubitParallelFor(vals.size(), [&](idx_t b, idx_t e) {
		for (idx_t r = b; r < e; r++) {
			const int64_t v = vals[r];
			int64_t id = null_key;
			if (v != null_key) {
				id = ix.ValueId(v);
				if (id < 0) {
					throw InvalidInputException("cubit_load: key %lld of \"%s\" lies outside the indexed domain", (long long)v, spec.key);
				}
			}
			if (binned) {
				bins[r] = id;
			}
		}
	});
	CubitCheck(cubit_gpu_index_create(gpu.handle, ix.cardinality, &ix.index_id));
	if (!binned) {
		CubitCheck(cubit_gpu_index_build(gpu.handle, ix.index_id, NumericCast<int32_t>(gcol), ix.base_value));
		return ix;
	}
	const int32_t tmp_col = NumericCast<int32_t>(gpu.column_names.size()); // first unused GPU column id
	CubitCheck(cubit_gpu_upload_column(gpu.handle, tmp_col, bins.get(), 8, gpu.row_count));
	const int rc = cubit_gpu_index_build(gpu.handle, ix.index_id, tmp_col, 0);
	cubit_gpu_drop_column(gpu.handle, tmp_col);
	CubitCheck(rc);
	return ix;
}

// ---------------------------------------------------------------- the registered index type
class CubitIndex : public BoundIndex {
public:
	CubitIndex(const string &name, const vector<column_t> &column_ids, TableIOManager &table_io_manager,
	           const vector<unique_ptr<Expression>> &unbound_expressions, AttachedDatabase &db, vector<CubitIndexSpec> specs_p)
	    : BoundIndex(name, CUBIT_INDEX_TYPE, IndexConstraintType::NONE, column_ids, table_io_manager, unbound_expressions, db),
	      gpu(make_shared_ptr<CubitGpuTable>()), specs(std::move(specs_p)) {
	}

	shared_ptr<CubitGpuTable> gpu;
	vector<CubitIndexSpec> specs;
	// index images: read from the database file at attach (consumed by the first load), refreshed at every checkpoint
	vector<vector<uint8_t>> images;
	vector<block_id_t> image_blocks; // blocks holding the images of the last checkpoint

	// IndexType::create_instance (index_type.hpp): binding after ATTACH / WAL replay.  The GPU copy is materialised
	// lazily, at the first statement that can use it (the optimizer hook has a ClientContext; this callback does not).
	static unique_ptr<BoundIndex> Create(CreateIndexInput &input) {
		vector<CubitIndexSpec> specs;
		auto opt = input.options.find("cubit_specs");
		if (opt != input.options.end()) {
			specs = CubitDecodeSpecs(opt->second.GetValue<string>());
		}
		auto index = make_uniq<CubitIndex>(input.name, input.column_ids, input.table_io_manager, input.unbound_expressions, input.db,
		                                   std::move(specs));
		index->ReadImages(input.storage_info); // (also recovers the specs: this version's IndexCatalogEntry::GetInfo
		                                       // drops `options` at checkpoint, index_catalog_entry.cpp:14-39)
		return std::move(index);
	}

	// ---- DML (bound_index.hpp:71-97)
	// INSERT: rows arrive in the table's column layout with their final row ids (DataTable::AppendToIndexes,
	// data_table.cpp:1000-1040; for transaction-local appends at commit, LocalStorage::AppendToIndexes)
	ErrorData Append(IndexLock &, DataChunk &entries, Vector &row_identifiers) override {
		std::lock_guard<std::mutex> lk(gpu->dml_lock);
		if (!gpu->loaded || entries.size() == 0) {
			return ErrorData(); // not resident (yet): the load reads the table as it is then
		}
		try {
			AppendLocked(entries, row_identifiers);
		} catch (std::exception &ex) {
			// the GPU copy is an accelerator: it never fails the user's statement, it steps aside
			gpu->usable = false;
			gpu->unusable_reason = ErrorData(ex).Message();
		}
		return ErrorData();
	}
	ErrorData Insert(IndexLock &lock, DataChunk &input, Vector &row_identifiers) override {
		return Append(lock, input, row_identifiers);
	}
	// DELETE (and the delete half of an UPDATE): called when the committed delete is cleaned up
	// (CleanupState::CleanupDelete → DataTable::RemoveFromIndexes → RowGroupCollection::RemoveFromIndexes,
	// row_group_collection.cpp:526-589) and when an append is reverted (data_table.cpp:962-980)
	void Delete(IndexLock &, DataChunk &entries, Vector &row_identifiers) override {
		std::lock_guard<std::mutex> lk(gpu->dml_lock);
		if (!gpu->loaded || entries.size() == 0) {
			return;
		}
		try {
			DeleteLocked(entries, row_identifiers);
		} catch (std::exception &ex) {
			gpu->usable = false;
			gpu->unusable_reason = ErrorData(ex).Message();
		}
	}
	void VerifyAppend(DataChunk &) override {
	}
	void VerifyAppend(DataChunk &, ConflictManager &) override {
	}
	void CheckConstraintsForChunk(DataChunk &, ConflictManager &) override {
	}
	void CommitDrop(IndexLock &) override {
		std::lock_guard<std::mutex> lk(gpu->dml_lock);
		gpu->Unload();
		auto &block_manager = table_io_manager.GetIndexBlockManager();
		for (auto id : image_blocks) {
			block_manager.MarkBlockAsModified(id);
		}
		image_blocks.clear();
		images.clear();
	}
	bool MergeIndexes(IndexLock &, BoundIndex &) override {
		return false;
	}
	void Vacuum(IndexLock &) override {
	}
	idx_t GetInMemorySize(IndexLock &) override {
		return 0; // device memory, not DuckDB's buffer pool
	}
	string VerifyAndToString(IndexLock &, const bool) override {
		return "CUBIT index (GPU resident)";
	}
	string GetConstraintViolationMessage(VerifyExistenceType, idx_t, DataChunk &) override {
		return "CUBIT indexes carry no constraint";
	}

	// ---- persistence (bound_index.hpp:117-118, IndexStorageInfo; WAL: write_ahead_log.cpp:260-273, replay
	// wal_replay.cpp:533-565).  One "allocator" per GPU index; its buffers are the index image
	// (cubit_gpu_index_serialize) cut into block-sized pieces.  get_buffers = true (WAL): the pieces are handed over
	// as memory; false (checkpoint): they are written to blocks here and the block pointers recorded.
	IndexStorageInfo GetStorageInfo(const bool get_buffers) override {
		std::lock_guard<std::mutex> lk(gpu->dml_lock);
		RefreshImages();
		IndexStorageInfo info(name);
		info.root = 0;
		auto &block_manager = table_io_manager.GetIndexBlockManager();
		const idx_t piece = block_manager.GetBlockSize();
		vector<block_id_t> new_blocks;
		for (idx_t i = 0; i < MaxValue<idx_t>(1, images.size()); i++) {
			FixedSizeAllocatorInfo alloc;
			alloc.segment_size = i < images.size() ? images[i].size() : 0; // image bytes
			vector<IndexBufferInfo> bufs;
			for (idx_t at = 0; i < images.size() && at < images[i].size(); at += piece) {
				const idx_t n = MinValue<idx_t>(piece, images[i].size() - at);
				alloc.buffer_ids.push_back(alloc.buffer_ids.size());
				alloc.segment_counts.push_back(1);
				alloc.allocation_sizes.push_back(n);
				if (get_buffers || block_manager.InMemory()) {
					alloc.block_pointers.emplace_back();
					bufs.emplace_back(images[i].data() + at, n);
				} else {
					shared_ptr<BlockHandle> block;
					auto handle = block_manager.buffer_manager.Allocate(MemoryTag::ART_INDEX, piece, false, &block);
					memcpy(handle.Ptr(), images[i].data() + at, n);
					const auto id = block_manager.GetFreeBlockId();
					block_manager.ConvertToPersistent(id, std::move(block));
					alloc.block_pointers.emplace_back(id, 0);
					new_blocks.push_back(id);
				}
			}
			info.allocator_infos.push_back(std::move(alloc));
			if (get_buffers) {
				info.buffers.push_back(std::move(bufs));
			}
		}
		if (!get_buffers && !block_manager.InMemory()) {
			for (auto id : image_blocks) { // the previous checkpoint's image
				block_manager.MarkBlockAsModified(id);
			}
			image_blocks = std::move(new_blocks);
		}
		// last "allocator": the specs (key, domain, binning of every GPU index), one character per buffer id
		FixedSizeAllocatorInfo spec_alloc;
		spec_alloc.segment_size = SPEC_MAGIC;
		for (auto ch : CubitEncodeSpecs(specs)) {
			spec_alloc.buffer_ids.push_back(static_cast<uint8_t>(ch));
		}
		info.allocator_infos.push_back(std::move(spec_alloc));
		if (get_buffers) {
			info.buffers.emplace_back();
		}
		return info;
	}

private:
	static constexpr idx_t SPEC_MAGIC = 0xC0B175BEC5ULL;

	void ReadImages(const IndexStorageInfo &info) {
		auto &block_manager = table_io_manager.GetIndexBlockManager();
		for (auto &alloc : info.allocator_infos) {
			if (alloc.segment_size == SPEC_MAGIC) {
				string text;
				for (auto ch : alloc.buffer_ids) {
					text.push_back(static_cast<char>(ch));
				}
				if (specs.empty()) {
					specs = CubitDecodeSpecs(text);
				}
				continue;
			}
			vector<uint8_t> image;
			for (idx_t j = 0; j < alloc.block_pointers.size() && j < alloc.allocation_sizes.size(); j++) {
				if (!alloc.block_pointers[j].IsValid()) {
					image.clear();
					break;
				}
				auto block = block_manager.RegisterBlock(alloc.block_pointers[j].block_id);
				auto handle = block_manager.buffer_manager.Pin(block);
				image.insert(image.end(), handle.Ptr(), handle.Ptr() + alloc.allocation_sizes[j]);
				image_blocks.push_back(alloc.block_pointers[j].block_id);
			}
			if (image.size() != alloc.segment_size) {
				image.clear();
			}
			images.push_back(std::move(image));
		}
	}

	// the images of the resident indexes (the checkpoint / WAL write what the GPU holds NOW, pending deltas included)
	void RefreshImages() {
		if (!gpu->loaded) {
			return; // never used since attach: what was read from the file is still current
		}
		images.assign(gpu->indexes.size(), {});
		for (idx_t i = 0; i < gpu->indexes.size(); i++) {
			void *img = nullptr;
			uint64_t bytes = 0;
			if (cubit_gpu_index_serialize(gpu->handle, gpu->indexes[i].index_id, &img, &bytes) == CUBIT_OK) {
				images[i].assign(static_cast<uint8_t *>(img), static_cast<uint8_t *>(img) + bytes);
				cubit_gpu_free_image(img);
			}
		}
	}

	void AppendLocked(DataChunk &entries, Vector &row_identifiers) {
		const idx_t n = entries.size();
		row_identifiers.Flatten(n);
		auto row_ids = FlatVector::GetData<row_t>(row_identifiers);
		if (row_ids[0] != NumericCast<row_t>(gpu->row_count) || row_ids[n - 1] != row_ids[0] + NumericCast<row_t>(n - 1)) {
			throw InvalidInputException("appended row ids [%lld, %lld] do not continue the %llu resident rows", (long long)row_ids[0],
			                            (long long)row_ids[n - 1], (unsigned long long)gpu->row_count);
		}
		vector<vector<int64_t>> cols(gpu->column_names.size(), vector<int64_t>(n));
		vector<cubit_append_column> ac;
		bool any_null = false;
		for (idx_t g = 0; g < cols.size(); g++) {
			Vector copy(entries.data[gpu->table_column[g]]);
			CubitRawInt64(copy, n, cols[g].data(), nullptr, 0, any_null);
			ac.push_back(cubit_append_column {NumericCast<int32_t>(g), 8, cols[g].data()});
		}
		CubitCheck(cubit_gpu_append_rows(gpu->handle, n, ac.data(), NumericCast<uint32_t>(ac.size())));
		gpu->row_count += n;
		cubit_dml_append_rows += n;
		if (any_null) {
			// NULLs in appended rows: the validity masks would have to grow row by row — rare in append-mostly
			// fact tables; the GPU copy steps aside until it is reloaded
			throw InvalidInputException("NULL in appended rows");
		}
		// indexes built from a resident column were extended by the library; binned indexes (built from a
		// temporary bin-id column) get the new rows' bits as pending deltas: 0 XOR 1
		for (auto &ix : gpu->indexes) {
			vector<uint32_t> values;
			vector<int64_t> rows;
			for (idx_t r = 0; r < n; r++) {
				const int64_t id = ix.ValueId(cols[ix.key_gpu_column][r]);
				if (id < 0) {
					throw InvalidInputException("appended key %lld lies outside the indexed domain", (long long)cols[ix.key_gpu_column][r]);
				}
				if (!ix.bin_lo.empty()) {
					values.push_back(NumericCast<uint32_t>(id));
					rows.push_back(row_ids[r]);
				}
			}
			if (!rows.empty()) {
				CubitCheck(cubit_gpu_add_delta_pairs(gpu->handle, ix.index_id, values.data(), rows.data(), rows.size()));
			}
		}
	}

	void DeleteLocked(DataChunk &entries, Vector &row_identifiers) {
		const idx_t n = entries.size();
		UnifiedVectorFormat rid;
		row_identifiers.ToUnifiedFormat(n, rid);
		auto row_ids = UnifiedVectorFormat::GetData<row_t>(rid);
		for (auto &ix : gpu->indexes) {
			Vector copy(entries.data[ix.key_table_column]);
			vector<int64_t> keys(n);
			bool any_null = false;
			CubitRawInt64(copy, n, keys.data(), nullptr, 0, any_null);
			vector<uint32_t> values;
			vector<int64_t> rows;
			for (idx_t r = 0; r < n; r++) {
				const row_t row = row_ids[rid.sel->get_index(r)];
				const int64_t id = ix.ValueId(keys[r]);
				if (id >= 0 && row >= 0 && NumericCast<idx_t>(row) < gpu->row_count) {
					values.push_back(NumericCast<uint32_t>(id)); // the row's bit in B_v is flipped off at query time
					rows.push_back(row);
				}
			}
			if (!rows.empty()) {
				CubitCheck(cubit_gpu_add_delta_pairs(gpu->handle, ix.index_id, values.data(), rows.data(), rows.size()));
				cubit_dml_delta_pairs += rows.size();
			}
		}
	}
};

// the table's CUBIT index, bound if necessary (after ATTACH it sits in the index list unbound until someone asks)
static optional_ptr<CubitIndex> CubitFindIndex(ClientContext &context, TableCatalogEntry &table) {
	if (!table.IsDuckTable()) {
		return nullptr;
	}
	auto info = table.GetStorage().GetDataTableInfo();
	bool present = false;
	info->GetIndexes().Scan([&](Index &index) {
		present |= index.GetIndexType() == CUBIT_INDEX_TYPE;
		return present;
	});
	if (!present) {
		return nullptr;
	}
	info->InitializeIndexes(context, CUBIT_INDEX_TYPE);
	optional_ptr<CubitIndex> found;
	info->GetIndexes().Scan([&](Index &index) {
		if (index.GetIndexType() == CUBIT_INDEX_TYPE && index.IsBound()) {
			found = &index.Cast<CubitIndex>();
		}
		return found != nullptr;
	});
	return found;
}

static TableCatalogEntry &CubitResolveTable(ClientContext &context, const string &name) {
	auto qname = QualifiedName::Parse(name);
	return Catalog::GetEntry<TableCatalogEntry>(context, qname.catalog.empty() ? INVALID_CATALOG : qname.catalog,
	                                            qname.schema.empty() ? DEFAULT_SCHEMA : qname.schema, qname.name);
}

static shared_ptr<CubitGpuTable> CubitLookup(ClientContext &context, const string &name) {
	auto index = CubitFindIndex(context, CubitResolveTable(context, name));
	if (!index || !index->gpu->loaded) {
		throw InvalidInputException("no CUBIT GPU index loaded for table \"%s\" (CALL cubit_load first)", name);
	}
	return index->gpu;
}

// devices named by CUBIT_GPU_DEVICES ("all", a count, or a comma-separated list; a device may repeat); empty = default
static vector<int> CubitDeviceList() {
	vector<int> out;
	const char *env = getenv("CUBIT_GPU_DEVICES");
	if (!env || !*env) {
		return out;
	}
	int have = 0;
	if (cubit_gpu_device_count(&have) != CUBIT_OK || have <= 0) {
		return out;
	}
	string spec(env);
	if (spec == "all") {
		for (int d = 0; d < have; d++) {
			out.push_back(d);
		}
		return out;
	}
	if (spec.find(',') == string::npos) {
		int n = 0;
		try {
			n = std::stoi(spec);
		} catch (...) {
			n = 0;
		}
		for (int d = 0; d < n && d < have; d++) {
			out.push_back(d);
		}
		return out;
	}
	size_t at = 0;
	while (at <= spec.size()) {
		const size_t comma = spec.find(',', at);
		const string tok = spec.substr(at, comma == string::npos ? string::npos : comma - at);
		try {
			const int d = std::stoi(tok);
			if (d >= 0 && d < have) {
				out.push_back(d);
			}
		} catch (...) {
		}
		if (comma == string::npos) {
			break;
		}
		at = comma + 1;
	}
	return out;
}

// ---------------------------------------------------------------- materialise the GPU copy
// Pulls rowid and every GPU-eligible column through a second connection and places every row at ITS ROW ID: the GPU
// row position is the table's row id (rowids are dense table positions, row_group.cpp:511-514, stable under
// DELETE / UPDATE), whatever order the chunks arrive in and whether or not rows were deleted before — positions of
// deleted rows hold a NULL key and are never selected.
static void CubitMaterialise(ClientContext &context, TableCatalogEntry &entry, CubitIndex &index) {
	auto &gpu = *index.gpu;
	const bool timing = getenv("CUBIT_LOAD_TIMING") != nullptr; // phase times on stderr
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (timing) {
			const auto now = std::chrono::steady_clock::now();
			fprintf(stderr, "cubit_load %s: %s %.1f ms\n", entry.name.c_str(), what,
			        std::chrono::duration<double, std::milli>(now - t_prev).count());
			t_prev = now;
		}
	};
	gpu.Unload();
	gpu.usable = true;
	gpu.unusable_reason.clear();
	gpu.column_names.clear();
	gpu.table_column.clear();
	gpu.column_types.clear();
	string select_list = "rowid";
	for (auto &col : entry.GetColumns().Logical()) {
		if (!col.Generated() && CubitEligibleType(col.Type())) {
			select_list += ", " + KeywordHelper::WriteOptionallyQuoted(col.Name());
			gpu.column_names.push_back(col.Name());
			gpu.table_column.push_back(col.Logical().index);
			gpu.column_types.push_back(col.Type());
		}
	}
	auto &storage = entry.GetStorage();
	const idx_t total = storage.GetTotalRows();
	if (total == 0) {
		throw InvalidInputException("cubit_load: table \"%s\" is empty", entry.name);
	}
	Connection con(*context.db);
	auto res = con.Query("SELECT " + select_list + " FROM " + KeywordHelper::WriteOptionallyQuoted(entry.ParentCatalog().GetName()) + "." +
	                     KeywordHelper::WriteOptionallyQuoted(entry.ParentSchema().name) + "." +
	                     KeywordHelper::WriteOptionallyQuoted(entry.name));
	if (res->HasError()) {
		throw InvalidInputException("cubit_load: %s", res->GetError());
	}
	lap("SELECT through a second connection");
	const idx_t n_cols = gpu.column_names.size();
	const int64_t null_key = NumericLimits<int64_t>::Minimum();
	// (uninitialised: every position is written by the scan below, or — rows deleted before the load — nulled after it)
	vector<unique_ptr<int64_t[]>> cols(n_cols);
	for (auto &c : cols) {
		c.reset(new int64_t[total]);
	}
	// NULLs: one validity mask per column in the reference's own layout (ValidityMask words), built only for
	// columns that hold a NULL; NULL keys are not indexed (plan_create_index.cpp:60-78 filters them out)
	vector<vector<uint64_t>> valid(n_cols);
	vector<uint64_t> seen((total + 63) / 64, 0); // row ids that arrived (the others were deleted before the load)
	std::atomic<idx_t> rows_seen_a {0};
	std::mutex valid_mu;
	// the result's chunks are placed by several threads at once (ColumnDataCollection's parallel scan): one thread
	// moved 0.18 G values per second, 3.7 s for SF10 lineitem
	auto &collection = res->Collection();
	ColumnDataParallelScanState pstate;
	collection.InitializeScan(pstate);
	auto place = [&](idx_t, idx_t) {
		ColumnDataLocalScanState lstate;
		DataChunk chunk;
		collection.InitializeScanChunk(chunk);
		vector<int64_t> tmp(STANDARD_VECTOR_SIZE);
		while (collection.Scan(pstate, lstate, chunk)) {
			const idx_t n = chunk.size();
			if (n == 0) {
				continue;
			}
			chunk.data[0].Flatten(n);
			auto rid = FlatVector::GetData<row_t>(chunk.data[0]);
			const bool dense = rid[n - 1] - rid[0] == NumericCast<row_t>(n - 1);
			if (rid[0] < 0 || NumericCast<idx_t>(rid[n - 1]) >= total) {
				throw InvalidInputException("cubit_load: the table changed while it was being read");
			}
			for (idx_t r = 0; r < n; r++) {
				__atomic_fetch_or(&seen[rid[r] / 64], uint64_t(1) << (rid[r] % 64), __ATOMIC_RELAXED);
			}
			for (idx_t k = 0; k < n_cols; k++) {
				bool any_null = false;
				int64_t *dst = dense ? cols[k].get() + rid[0] : tmp.data();
				CubitRawInt64(chunk.data[k + 1], n, dst, nullptr, 0, any_null);
				if (!dense) {
					for (idx_t r = 0; r < n; r++) {
						cols[k][rid[r]] = tmp[r];
					}
				}
				if (any_null) {
					{
						std::lock_guard<std::mutex> lk(valid_mu);
						if (valid[k].empty()) {
							valid[k].assign((total + 63) / 64, ~uint64_t(0));
						}
					}
					chunk.data[k + 1].Flatten(n);
					auto &mask = FlatVector::Validity(chunk.data[k + 1]);
					for (idx_t r = 0; r < n; r++) {
						if (!mask.RowIsValid(r)) {
							__atomic_fetch_and(&valid[k][rid[r] / 64], ~(uint64_t(1) << (rid[r] % 64)), __ATOMIC_RELAXED);
						}
					}
				}
			}
			rows_seen_a += n;
		}
	};
	{
		// one worker per host thread, all pulling chunks from the same parallel scan state
		const idx_t nt = total < (idx_t(1) << 18) ? 1 : CubitHostThreads();
		std::exception_ptr err;
		std::mutex mu;
		vector<std::thread> th;
		for (idx_t i = 1; i < nt; i++) {
			th.emplace_back([&]() {
				try {
					place(0, 0);
				} catch (...) {
					std::lock_guard<std::mutex> lk(mu);
					if (!err) {
						err = std::current_exception();
					}
				}
			});
		}
		try {
			place(0, 0);
		} catch (...) {
			std::lock_guard<std::mutex> lk(mu);
			if (!err) {
				err = std::current_exception();
			}
		}
		for (auto &t : th) {
			t.join();
		}
		if (err) {
			std::rethrow_exception(err);
		}
	}
	const idx_t rows_seen = rows_seen_a.load();
	if (rows_seen != total) { // positions of rows deleted before the load: a NULL key, never selected
		CubitParallelFor(total, [&](idx_t b, idx_t e) {
			for (idx_t r = b; r < e; r++) {
				if (!((seen[r / 64] >> (r % 64)) & 1)) {
					for (idx_t k = 0; k < n_cols; k++) {
						cols[k][r] = null_key;
					}
				}
			}
		});
	}
	lap("rows placed at their row ids (host)");
	const bool no_deleted_rows = rows_seen == total;
	gpu.row_count = total;
	// One GPU by default; CUBIT_GPU_DEVICES = "all" | a count | "0,1,3" cuts the table into contiguous row ranges of whole
	// segments over those devices behind ONE handle (cubit_gpu_create_sharded — the reference's parallel unit is the
	// row-group range of RowGroupCollection::NextParallelScan, row_group_collection.cpp:174-224).  Everything below
	// and every query / DML callback addresses that handle unchanged; the storage route and index images are per
	// shard and fall back to decoded rows / a rebuild at attach.
	auto devices = CubitDeviceList();
	if (devices.size() > 1) {
		CubitCheck(cubit_gpu_create_sharded(devices.data(), NumericCast<uint32_t>(devices.size()), gpu.row_count, 0, 65536,
		                                    &gpu.handle));
	} else {
		CubitCheck(cubit_gpu_create(devices.empty() ? 0 : devices[0], gpu.row_count, 0, 65536, &gpu.handle));
	}
	for (idx_t k = 0; k < n_cols; k++) {
		// compressed segments straight from the buffer manager when the column qualifies (physical position = row
		// id, so the route needs a table without deleted rows), decoded rows otherwise (a column with NULLs goes
		// the decoded way: its NULL rows must carry an out-of-domain value)
		const bool has_nulls = !valid[k].empty();
		if (has_nulls || !no_deleted_rows ||
		    !CubitUploadColumnSegments(context, entry, gpu.table_column[k], gpu.column_types[k], gpu.handle, NumericCast<int32_t>(k),
		                               gpu.row_count)) {
			CubitCheck(cubit_gpu_upload_column(gpu.handle, NumericCast<int32_t>(k), cols[k].get(), 8, gpu.row_count));
		}
		if (has_nulls) {
			CubitCheck(cubit_gpu_upload_column_validity(gpu.handle, NumericCast<int32_t>(k), valid[k].data(), valid[k].size()));
		}
	}
	lap("columns uploaded");
	for (idx_t i = 0; i < index.specs.size(); i++) {
		auto &spec = index.specs[i];
		idx_t gcol = 0;
		while (gcol < n_cols && gpu.column_names[gcol] != spec.key) {
			gcol++;
		}
		if (gcol == n_cols) {
			throw InvalidInputException("cubit_load: key column \"%s\" not found or not GPU eligible", spec.key);
		}
		gpu.indexes.push_back(CubitBuildIndex(gpu, gcol, spec, gpu.column_types[gcol], i < index.images.size() ? &index.images[i] : nullptr));
	}
	lap("indexes built");
	// NULL-free columns that pack to <= 32 bits per value also get a FOR-bit-packed form (the resident analog of the
	// BitPacking segments the table is stored in): dense selections are then probed by streaming width/8 bytes per row.
	// The raw form stays (index builds, appends — an append drops the packed form of the columns it extends).
	if (!getenv("CUBIT_NO_PACK")) {
		for (idx_t k = 0; k < n_cols; k++) {
			if (valid[k].empty()) {
				uint64_t packed_bytes = 0;
				CubitCheck(cubit_gpu_pack_column(gpu.handle, NumericCast<int32_t>(k), 2, &packed_bytes));
			}
		}
	}
	lap("columns packed");
	gpu.loaded = true;
}

// ---------------------------------------------------------------- cubit_load(table, key, base, cardinality)
struct CubitLoadBindData : public TableFunctionData {
	string table;
	CubitIndexSpec spec;
	bool done = false;
};

static unique_ptr<FunctionData> CubitLoadBind(ClientContext &, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitLoadBindData>();
	bind->table = input.inputs[0].GetValue<string>();
	bind->spec.key = input.inputs[1].GetValue<string>();
	bind->spec.base = input.inputs[2].GetValue<int64_t>();
	bind->spec.cardinality = NumericCast<uint32_t>(input.inputs[3].GetValue<int64_t>());
	auto bin = input.named_parameters.find("bin");
	if (bin != input.named_parameters.end()) {
		bind->spec.bin = bin->second.GetValue<string>();
	}
	return_types.emplace_back(LogicalType::BIGINT);
	names.emplace_back("rows_indexed");
	return std::move(bind);
}

static void CubitLoadFunction(ClientContext &context, TableFunctionInput &data_p, DataChunk &output) {
	auto &bind = data_p.bind_data->CastNoConst<CubitLoadBindData>();
	if (bind.done) {
		return;
	}
	auto &entry = CubitResolveTable(context, bind.table);
	if (!entry.IsDuckTable()) {
		throw InvalidInputException("cubit_load: \"%s\" is not a DuckDB table", bind.table);
	}
	auto &storage = entry.GetStorage();
	auto index = CubitFindIndex(context, entry);
	optional_ptr<DuckIndexEntry> catalog_entry;
	const string index_name = "cubit_" + entry.name;
	if (!index) {
		// ---- first load of this table: register a CUBIT index on it (catalog entry + BoundIndex in the table's index
		// list), the way PhysicalCreateARTIndex::Finalize does for ART (physical_create_art_index.cpp:155-189);
		// CREATE INDEX ... USING CUBIT itself is rejected by the planner for non-ART types (plan_create_index.cpp:33-38)
		MetaTransaction::Get(context).ModifyDatabase(entry.ParentCatalog().GetAttached());
		auto info = make_uniq<CreateIndexInfo>();
		info->catalog = entry.ParentCatalog().GetName();
		info->schema = entry.ParentSchema().name;
		info->table = entry.name;
		info->index_name = index_name;
		info->index_type = CUBIT_INDEX_TYPE;
		info->constraint_type = IndexConstraintType::NONE;
		vector<unique_ptr<Expression>> unbound;
		vector<column_t> column_ids;
		for (auto &col : entry.GetColumns().Logical()) {
			info->scan_types.push_back(col.Type());
			info->names.push_back(col.Name());
			if (col.Generated() || !CubitEligibleType(col.Type())) {
				continue;
			}
			// every GPU-resident column is an index column: DML that touches one of them reaches the index
			// (an UPDATE of such a column is planned as DELETE + INSERT, BoundIndex::IndexIsUpdated)
			info->expressions.push_back(make_uniq<ColumnRefExpression>(col.Name(), entry.name));
			info->parsed_expressions.push_back(make_uniq<ColumnRefExpression>(col.Name(), entry.name));
			unbound.push_back(make_uniq<BoundColumnRefExpression>(col.Name(), col.Type(), ColumnBinding(0, column_ids.size())));
			column_ids.push_back(col.StorageOid());
		}
		info->scan_types.emplace_back(LogicalType::ROW_TYPE);
		info->column_ids = column_ids;
		info->options["cubit_specs"] = Value(CubitEncodeSpecs({bind.spec}));
		auto &schema = entry.ParentSchema();
		auto created = schema.CreateIndex(schema.GetCatalogTransaction(context), *info, entry);
		if (!created) {
			throw InvalidInputException("cubit_load: could not register the index \"%s\"", index_name);
		}
		auto &dentry = created->Cast<DuckIndexEntry>();
		dentry.initial_index_size = 0;
		dentry.info = make_shared_ptr<IndexDataTableInfo>(storage.GetDataTableInfo(), dentry.name);
		for (auto &expr : info->parsed_expressions) {
			dentry.parsed_expressions.push_back(expr->Copy());
		}
		auto bound = make_uniq<CubitIndex>(index_name, column_ids, TableIOManager::Get(storage), unbound, storage.db,
		                                   vector<CubitIndexSpec> {bind.spec});
		index = bound.get();
		storage.AddIndex(std::move(bound));
	}
	auto &gpu = *index->gpu;
	std::lock_guard<std::mutex> lk(gpu.dml_lock);
	idx_t slot = 0;
	while (slot < index->specs.size() && index->specs[slot].key != bind.spec.key) {
		slot++;
	}
	const bool new_key = slot == index->specs.size();
	if (new_key) {
		index->specs.push_back(bind.spec);
	} else {
		index->specs[slot] = bind.spec; // same key again: new domain / binning
	}
	{ // the catalog entry carries the specs across restarts
		auto entry_ptr = entry.ParentSchema().GetEntry(entry.ParentSchema().GetCatalogTransaction(context), CatalogType::INDEX_ENTRY,
		                                               index_name);
		if (entry_ptr) {
			entry_ptr->Cast<IndexCatalogEntry>().options["cubit_specs"] = Value(CubitEncodeSpecs(index->specs));
		}
	}
	if (gpu.loaded && gpu.usable && new_key && gpu.row_count == storage.GetTotalRows()) {
		// the table is already resident: just index one more of its columns
		idx_t gcol = 0;
		while (gcol < gpu.column_names.size() && gpu.column_names[gcol] != bind.spec.key) {
			gcol++;
		}
		if (gcol == gpu.column_names.size()) {
			index->specs.pop_back();
			throw InvalidInputException("cubit_load: key column \"%s\" not found or not GPU eligible", bind.spec.key);
		}
		gpu.indexes.push_back(CubitBuildIndex(gpu, gcol, bind.spec, gpu.column_types[gcol], nullptr));
	} else {
		index->images.clear(); // an explicit load always rebuilds from the table
		try {
			CubitMaterialise(context, entry, *index);
		} catch (...) {
			if (new_key) {
				index->specs.pop_back();
			}
			throw;
		}
	}
	output.SetValue(0, 0, Value::BIGINT(NumericCast<int64_t>(gpu.row_count)));
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

// DataChunk hand-off.  The result stays on the GPU; its rows leave in WINDOWS of 64 DataChunks.
// Up to MaxThreads() workers drain windows in parallel: a worker claims the next window index (the batch index the
// ordered sinks sort by, the pattern of seq_scan's row-group batches, table_scan.cpp:179-189), and while it serves
// one window out of its page-locked buffers the copy of the NEXT window it claimed is already in flight
// (cubit_gpu_fetch_async on the library's copy streams) — two buffer sets per worker, no synchronise per chunk.
static_assert(STANDARD_VECTOR_SIZE == CUBIT_WIRE_CHUNK, "a wire frame is one DataChunk");
struct CubitScanGlobalState : public GlobalTableFunctionState {
	cubit_gpu_result *result = nullptr;
	vector<column_t> column_ids;
	idx_t row_count = 0;
	uint64_t sum_lo = 0;
	int64_t sum_hi = 0;
	uint64_t agg_rows = 0; // non-NULL inputs of the pushed-down SUM
	bool agg_emitted = false;
	bool want_rowid = false;
	idx_t n_value_cols = 0;
	vector<bool> col_has_nulls; // per projected value column: the result carries a validity mask for it
	// rows leave the GPU in the narrow wire format (cubit_gpu_wire.h: per DataChunk a base + 1/2/4/8-byte deltas the
	// device writes straight into the page-locked window; widened into the output vectors one chunk at a time).  On a
	// sharded table the shard that holds a window writes it; the rare window that straddles two shards keeps the
	// 8-byte copies.
	bool narrow_wire = false;
	uint64_t wire_bytes = 0;
	static constexpr idx_t WINDOW_ROWS = 64 * STANDARD_VECTOR_SIZE;
	std::atomic<idx_t> next_window {0};
	idx_t n_windows = 0;
	shared_ptr<CubitGpuTable> pool_owner;
	~CubitScanGlobalState() override {
		cubit_gpu_free_result(result);
	}
	idx_t MaxThreads() const override {
		// a window is 1-3 MB of PCIe traffic plus 64 memcpy'd DataChunks: worth a worker each, up to a handful
		return MaxValue<idx_t>(1, MinValue<idx_t>(16, n_windows / 2));
	}
};

struct CubitScanLocalState : public LocalTableFunctionState {
	struct Window {
		idx_t index = DConstants::INVALID_INDEX, begin = 0, end = 0;
		int64_t *rowids = nullptr;
		vector<int64_t *> cols;
		vector<vector<uint64_t>> validity; // ValidityMask words of the window, empty = no NULL in the window
		cubit_gpu_fetch_ticket *ticket = nullptr;
		void *wire = nullptr; // narrow-wire window (instead of rowids / cols)
		uint64_t wire_bytes = 0;
		bool narrow = false;  // this window arrived as a wire
	};
	Window win[2];
	int cur = 0;          // the window being served
	idx_t offset = 0;     // next result row to emit (inside win[cur])
	bool primed = false;
	shared_ptr<CubitGpuTable> pool_owner;
	~CubitScanLocalState() override {
		for (auto &w : win) {
			cubit_gpu_fetch_wait(w.ticket);
			if (pool_owner) {
				pool_owner->ReleaseWindow(w.rowids);
				for (auto p : w.cols) {
					pool_owner->ReleaseWindow(p);
				}
				pool_owner->ReleaseWire(w.wire, w.wire_bytes);
			}
		}
	}
};

static unique_ptr<FunctionData> CubitScanBind(ClientContext &context, TableFunctionBindInput &input,
                                              vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(context, input.inputs[0].GetValue<string>());
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
	state->want_rowid = want_rowid;
	state->n_value_cols = cols.size();
	for (idx_t c = 0; c < cols.size(); c++) {
		// (a sharded result has no single device pointer: ask the library per window instead)
		uint32_t n_shards = 1;
		cubit_gpu_shard_count(gpu.handle, &n_shards);
		state->col_has_nulls.push_back(n_shards > 1 || info.d_validity[c] != nullptr);
	}
	state->n_windows = (state->row_count + CubitScanGlobalState::WINDOW_ROWS - 1) / CubitScanGlobalState::WINDOW_ROWS;
	{
		uint32_t n_shards = 1;
		cubit_gpu_shard_count(gpu.handle, &n_shards);
		const uint32_t streams = NumericCast<uint32_t>((want_rowid ? 1 : 0) + cols.size());
		state->narrow_wire = streams > 0 && bind.agg_col < 0 && !getenv("CUBIT_WIDE_HANDOFF");
		(void)n_shards; // (a window that straddles two shards falls back to the 8-byte copies on its own)
		state->wire_bytes = cubit_wire_bytes(CubitScanGlobalState::WINDOW_ROWS, streams);
	}
	return std::move(state);
}

static unique_ptr<GlobalTableFunctionState> CubitScanInitGlobal(ClientContext &, TableFunctionInitInput &input) {
	return CubitRunQuery(input.bind_data->Cast<CubitScanBindData>(), input.column_ids, input.projection_ids);
}

static unique_ptr<LocalTableFunctionState> CubitScanInitLocal(ExecutionContext &, TableFunctionInitInput &,
                                                             GlobalTableFunctionState *global_state) {
	auto local = make_uniq<CubitScanLocalState>();
	local->pool_owner = global_state->Cast<CubitScanGlobalState>().pool_owner;
	return std::move(local);
}

// claim the next window of the result and start its device → host copy into `w`'s page-locked buffers
static void CubitClaimWindow(CubitScanGlobalState &state, CubitScanLocalState &local, CubitScanLocalState::Window &w) {
	w.index = state.next_window.fetch_add(1);
	w.ticket = nullptr;
	if (w.index >= state.n_windows) {
		w.index = DConstants::INVALID_INDEX;
		return;
	}
	w.begin = w.index * CubitScanGlobalState::WINDOW_ROWS;
	w.end = MinValue<idx_t>(w.begin + CubitScanGlobalState::WINDOW_ROWS, state.row_count);
	w.narrow = false;
	if (state.narrow_wire) {
		if (!w.wire) {
			w.wire = local.pool_owner->AcquireWire(state.wire_bytes);
			w.wire_bytes = state.wire_bytes;
		}
		const int rc = cubit_gpu_fetch_wire_async(state.result, w.begin, w.end - w.begin, state.want_rowid ? 1 : 0,
		                                          NumericCast<uint32_t>(state.n_value_cols), w.wire, w.wire_bytes,
		                                          &w.ticket);
		if (rc == CUBIT_OK) {
			w.narrow = true;
		} else if (rc != CUBIT_ESTATE) { // ESTATE: the window straddles two shards — this one travels as 8-byte copies
			CubitCheck(rc);
		}
	}
	if (!w.narrow) {
		const uint64_t win_bytes = CubitScanGlobalState::WINDOW_ROWS * sizeof(int64_t);
		while (w.cols.size() < state.n_value_cols) {
			w.cols.push_back(static_cast<int64_t *>(local.pool_owner->AcquireWindow(win_bytes)));
		}
		if (state.want_rowid && !w.rowids) {
			w.rowids = static_cast<int64_t *>(local.pool_owner->AcquireWindow(win_bytes));
		}
		vector<void *> ptrs(w.cols.begin(), w.cols.end());
		CubitCheck(cubit_gpu_fetch_async(state.result, w.begin, w.end - w.begin, state.want_rowid ? w.rowids : nullptr,
		                                 NumericCast<uint32_t>(ptrs.size()), ptrs.data(), &w.ticket));
	}
	// NULLs: the validity mask of every projected value (StandardColumnData::FetchRow = validity + data) — asked for
	// only when the query touched a NULL-bearing column at all
	w.validity.resize(state.n_value_cols);
	for (idx_t c = 0; c < state.n_value_cols; c++) {
		w.validity[c].clear();
		if (!state.col_has_nulls[c]) {
			continue;
		}
		int all_valid = 1;
		w.validity[c].assign((w.end - w.begin + 63) / 64, 0);
		CubitCheck(cubit_gpu_fetch_validity(state.result, NumericCast<uint32_t>(c), w.begin, w.end - w.begin,
		                                    w.validity[c].data(), &all_valid));
		if (all_valid) {
			w.validity[c].clear();
		}
	}
}

static void CubitScanFunction(ClientContext &, TableFunctionInput &data_p, DataChunk &output) {
	auto &state = data_p.global_state->Cast<CubitScanGlobalState>();
	auto &local = data_p.local_state->Cast<CubitScanLocalState>();
	if (!local.primed) { // two windows in flight per worker from the start
		local.primed = true;
		CubitClaimWindow(state, local, local.win[0]);
		CubitClaimWindow(state, local, local.win[1]);
		local.cur = 0;
		local.offset = local.win[0].begin;
		if (local.win[0].index != DConstants::INVALID_INDEX) {
			CubitCheck(cubit_gpu_fetch_wait(local.win[0].ticket));
			local.win[0].ticket = nullptr;
		}
	}
	auto *w = &local.win[local.cur];
	if (w->index != DConstants::INVALID_INDEX && local.offset >= w->end) {
		// this window is served: re-arm its buffers with the next unclaimed window, switch to the other (its copy has
		// been in flight all along)
		CubitClaimWindow(state, local, *w);
		local.cur ^= 1;
		w = &local.win[local.cur];
		if (w->index != DConstants::INVALID_INDEX) {
			CubitCheck(cubit_gpu_fetch_wait(w->ticket));
			w->ticket = nullptr;
			local.offset = w->begin;
		}
	}
	if (w->index == DConstants::INVALID_INDEX) {
		return; // chunk.size() == 0 → PhysicalTableScan::GetData returns FINISHED
	}
	const idx_t scan_count = MinValue<idx_t>(STANDARD_VECTOR_SIZE, w->end - local.offset);
	const idx_t rel = local.offset - w->begin; // a multiple of 2048: word aligned in the window's masks
	idx_t value_col = 0;
	const uint64_t wire_chunk = rel / STANDARD_VECTOR_SIZE; // windows are cut at DataChunk boundaries
	for (idx_t i = 0; i < state.column_ids.size(); i++) {
		auto dst = FlatVector::GetData<int64_t>(output.data[i]);
		if (state.column_ids[i] == COLUMN_IDENTIFIER_ROW_ID) {
			if (w->narrow) { // stream 0: widened straight into the output vector
				if (cubit_wire_unpack_chunk(w->wire, 0, wire_chunk, dst, 8) != int(scan_count)) {
					throw InternalException("cubit: malformed wire window");
				}
			} else {
				memcpy(dst, w->rowids + rel, scan_count * sizeof(int64_t));
			}
			continue;
		}
		if (w->narrow) {
			const uint32_t stream = NumericCast<uint32_t>((state.want_rowid ? 1 : 0) + value_col);
			if (cubit_wire_unpack_chunk(w->wire, stream, wire_chunk, dst, 8) != int(scan_count)) {
				throw InternalException("cubit: malformed wire window");
			}
		} else {
			memcpy(dst, w->cols[value_col] + rel, scan_count * sizeof(int64_t));
		}
		auto &words = w->validity[value_col];
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
	local.offset += scan_count;
}

// ordered sinks (batch collectors, ORDER-preserving inserts) sort the chunks of parallel workers by this
static idx_t CubitScanGetBatchIndex(ClientContext &, const FunctionData *, LocalTableFunctionState *local_state,
                                    GlobalTableFunctionState *) {
	auto &local = local_state->Cast<CubitScanLocalState>();
	auto &w = local.win[local.cur];
	return w.index == DConstants::INVALID_INDEX ? 0 : w.index;
}

// ---------------------------------------------------------------- cubit_agg(table, lo, hi, column) → (count, sum)
static unique_ptr<FunctionData> CubitAggBind(ClientContext &context, TableFunctionBindInput &input,
                                             vector<LogicalType> &return_types, vector<string> &names) {
	auto bind = make_uniq<CubitScanBindData>();
	bind->gpu = CubitLookup(context, input.inputs[0].GetValue<string>());
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
static bool CubitPlanGet(ClientContext &context, LogicalGet &get, shared_ptr<CubitGpuTable> &gpu,
                         vector<CubitScanBindData::Range> &ranges) {
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
	// the accelerator is the table's own CUBIT index (no name-keyed registry: another table, schema or database with
	// the same name has its own index list)
	auto index = CubitFindIndex(context, *table);
	if (!index) {
		CUBIT_WHY("table has no GPU index");
	}
	auto &storage = table->GetStorage();
	{
		std::lock_guard<std::mutex> lk(index->gpu->dml_lock);
		if (!index->gpu->loaded) {
			// bound from the catalog after ATTACH / restart: materialise the GPU copy now (index images from the
			// last checkpoint replace the index builds when they still describe the table)
			if (index->specs.empty()) {
				CUBIT_WHY("CUBIT index without specs");
			}
			try {
				CubitMaterialise(context, *table, *index);
			} catch (std::exception &ex) {
				index->gpu->usable = false;
				index->gpu->unusable_reason = ErrorData(ex).Message();
			}
		}
		if (!index->gpu->loaded || !index->gpu->usable) {
			if (getenv("CUBIT_DEBUG_REWRITE")) {
				fprintf(stderr, "cubit: GPU copy stepped aside: %s\n", index->gpu->unusable_reason.c_str());
			}
			CUBIT_WHY("GPU copy is not usable (reload with cubit_load)");
		}
		// the GPU copy must describe exactly the committed table: same row-id high-water mark (appends that bypassed
		// the index callbacks — none are known — or a reverted append would show here) ...
		if (index->gpu->row_count != storage.GetTotalRows()) {
			CUBIT_WHY("GPU copy and table differ in row count");
		}
	}
	// ... and this transaction must not have changes of its own: its uncommitted inserts live in its local storage and
	// its uncommitted deletes reach the index only at commit, so only the vanilla scan sees them
	if (DuckTransaction::Get(context, table->ParentCatalog()).ChangesMade()) {
		CUBIT_WHY("the transaction has uncommitted changes");
	}
	gpu = index->gpu;
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

static void CubitRewriteGet(ClientContext &context, LogicalGet &get) {
	shared_ptr<CubitGpuTable> gpu;
	vector<CubitScanBindData::Range> ranges;
	if (!CubitPlanGet(context, get, gpu, ranges)) {
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

static bool CubitTryAggregatePushdown(ClientContext &context, unique_ptr<LogicalOperator> &op) {
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
	if (!CubitPlanGet(context, get, gpu, ranges)) {
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

static void CubitRewritePlan(ClientContext &context, unique_ptr<LogicalOperator> &op) {
	if (CubitTryAggregatePushdown(context, op)) {
		return;
	}
	if (op->type == LogicalOperatorType::LOGICAL_GET) {
		CubitRewriteGet(context, op->Cast<LogicalGet>());
	}
	for (auto &child : op->children) {
		CubitRewritePlan(context, child);
	}
}

// DML needs nothing here: the table's CUBIT index receives it through BoundIndex::Append / Delete (see CubitIndex).
// The scan that FEEDS a DELETE / UPDATE may itself be rewritten (it only produces row ids and column values).
static void CubitOptimize(OptimizerExtensionInput &input, unique_ptr<LogicalOperator> &plan) {
	if (getenv("CUBIT_DISABLE_REWRITE")) { // (tests: the same statement through the vanilla scan of the same table)
		return;
	}
	if (getenv("CUBIT_NO_AGG_PUSHDOWN")) { // (tests: compare the row-returning scan with the pushed-down aggregate)
		std::function<void(LogicalOperator &)> walk = [&](LogicalOperator &o) {
			if (o.type == LogicalOperatorType::LOGICAL_GET) {
				CubitRewriteGet(input.context, o.Cast<LogicalGet>());
			}
			for (auto &child : o.children) {
				walk(*child);
			}
		};
		walk(*plan);
		return;
	}
	CubitRewritePlan(input.context, plan);
}

static TableFunction CubitScanTableFunction() {
	TableFunction scan("cubit_scan", {LogicalType::VARCHAR, LogicalType::BIGINT, LogicalType::BIGINT}, CubitScanFunction,
	                   CubitScanBind, CubitScanInitGlobal, CubitScanInitLocal);
	scan.get_batch_index = CubitScanGetBatchIndex; // parallel workers, order kept by the window index
	scan.projection_pushdown = true; // column_ids tell the GPU which columns to probe
	scan.filter_prune = true;        // columns used only by the (absorbed) filter are not produced
	return scan;
}

// ---------------------------------------------------------------- registration
void RegisterCubitGpuFunctions(DatabaseInstance &db) {
	// structs cross the C-ABI by value: a glue compiled against another header version must not run
	if (cubit_gpu_abi_version() != CUBIT_GPU_ABI_VERSION) {
		throw InvalidInputException("cubit_gpu: libcubit_gpu.so has ABI %d, this extension was built for ABI %d",
		                            cubit_gpu_abi_version(), CUBIT_GPU_ABI_VERSION);
	}
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

	// the index type (index_type_set.cpp:24-30): lets a database file that holds a CUBIT index be attached — the
	// index is bound through create_instance, maintained by DML and checkpointed like any other
	if (!DBConfig::GetConfig(db).GetIndexTypes().FindByName(CUBIT_INDEX_TYPE)) {
		IndexType cubit_type;
		cubit_type.name = CUBIT_INDEX_TYPE;
		cubit_type.create_instance = CubitIndex::Create;
		DBConfig::GetConfig(db).GetIndexTypes().RegisterIndexType(cubit_type);
	}
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
