/*
 * cubit_gpu.h — C-ABI of the B200-native CUBIT bitmap-index scan path.
 *
 * This is the drop-in boundary: the only symbols the
 * reference's C++ operator code binds to.  Everything is `extern "C"`, plain
 * pointers and sizes; no C++ / torch types cross it.  The reference-side
 * caller is the table function that PhysicalTableScan::GetData dispatches to
 * (reference: src/execution/operator/scan/physical_table_scan.cpp:82-103 →
 * TableFunction::function, src/include/duckdb/function/table_function.hpp:194)
 * and the index object that would sit beside ART as a BoundIndex
 * (src/include/duckdb/execution/index/bound_index.hpp:67-126).  INTEGRATION.md
 * shows the binding a DuckDB maintainer adds.
 *
 * Conventions
 *   - every function returns CUBIT_OK (0) or a negative CUBIT_E* code; the
 *     message is available from cubit_gpu_last_error() (thread-local).  No
 *     exception ever crosses the ABI (reference errors are C++ exceptions that
 *     the table function re-throws: table_function-c.cpp:203-212).
 *   - bit order is DuckDB's ValidityMask order: row r is bit (r % 64) of
 *     64-bit word r / 64 (src/include/duckdb/common/types/validity_mask.hpp:163-168).
 *   - row IDs are row_t = int64_t (src/include/duckdb/common/typedefs.hpp:19),
 *     ascending and duplicate free, the contract of ART::Scan
 *     (src/execution/index/art/art.cpp:974-985).  A shard created with
 *     row_base = B reports global row IDs B + local position.
 *   - there is NO CPU fallback.  If no CUDA device is usable every entry point
 *     that needs one fails with CUBIT_ENODEVICE.
 */
#ifndef CUBIT_GPU_H
#define CUBIT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUBIT_GPU_ABI_VERSION 5

/* error codes */
#define CUBIT_OK 0
#define CUBIT_EINVAL (-1)    /* bad argument / out of range id            */
#define CUBIT_ENODEVICE (-2) /* no usable CUDA device                      */
#define CUBIT_ECUDA (-3)     /* CUDA runtime error (message has the text)  */
#define CUBIT_ENOMEM (-4)    /* host or device allocation failed           */
#define CUBIT_ESTATE (-5)    /* call not valid in the object's state       */

/* limits */
#define CUBIT_MAX_STREAMS 64 /* bitvectors read by one query (Σ group sizes) */
#define CUBIT_MAX_PROBE_COLS 8

typedef struct cubit_gpu_table cubit_gpu_table;   /* one table shard on one GPU, or a table sharded over several */
typedef struct cubit_gpu_result cubit_gpu_result; /* one query's result set     */
typedef struct cubit_gpu_fetch_ticket cubit_gpu_fetch_ticket; /* one asynchronous hand-off copy in flight */

/* (index, value) names one value bitvector B_v of one CUBIT index. */
typedef struct cubit_bv_ref {
	int32_t index_id;
	uint32_t value_id;
} cubit_bv_ref;

/* One OR group: OR over refs[0..n_refs).  A range predicate lo<=x<=hi on an
 * indexed column is the OR over the value bitvectors in [lo,hi]. */
typedef struct cubit_pred_group {
	uint32_t n_refs;
	const cubit_bv_ref *refs;
} cubit_pred_group;

/* what the query materialises */
#define CUBIT_Q_ROWIDS (1u << 0)    /* sorted int64 row IDs (device resident, fetchable)      */
#define CUBIT_Q_BITVECTOR (1u << 1) /* the merged query bitvector Q (device resident)          */
#define CUBIT_Q_VALUES (1u << 2)    /* probe cols[] at the selected rows (device, fetchable)   */
#define CUBIT_Q_TIMING (1u << 3)    /* record per-kernel CUDA-event times in cubit_result_info */
#define CUBIT_Q_UNFUSED (1u << 4)   /* force merge → decode → probe as three separate kernels  */
#define CUBIT_Q_ASYNC (1u << 5)     /* enqueue only; cubit_gpu_result_wait() completes it      */
#define CUBIT_Q_FUSE_PROBE (1u << 6) /* probe inside the scan kernel (one launch) instead of the
                                        bit-driven probe kernel that follows it by default      */

/* fused aggregate over the selected rows (SUM semantics of the reference:
 * int64 input, 128-bit accumulator — sum.cpp:172-178, sum_helpers.hpp:92-113) */
#define CUBIT_AGG_NONE 0
#define CUBIT_AGG_SUM 1      /* SUM(col a)            */
#define CUBIT_AGG_SUM_PROD 2 /* SUM(col a * col b), int64 product (arithmetic.cpp:766-795) */
#define CUBIT_AGG_SUM_F64 3  /* SUM(col a) over a DOUBLE column; order of additions is not fixed, so the result
                               is reproducible to ~1e-15 relative and specified to 1e-12 (north_star) */

/* Q = AND_j ( OR_{i in groups[j]} ( B_i XOR D_i ) ) */
typedef struct cubit_query {
	uint32_t n_groups;
	const cubit_pred_group *groups;
	uint32_t flags;       /* CUBIT_Q_*                                         */
	uint32_t n_cols;      /* projected columns to probe (CUBIT_Q_VALUES)        */
	const int32_t *cols;  /* column ids, as given to cubit_gpu_upload_column    */
	int32_t agg_kind;     /* CUBIT_AGG_*                                        */
	int32_t agg_col_a;
	int32_t agg_col_b;
} cubit_query;

typedef struct cubit_result_info {
	uint64_t count;         /* rows selected = popcount(Q)                        */
	uint64_t sum_lo;        /* 128-bit two's complement aggregate, low limb      */
	int64_t sum_hi;         /*                                        high limb  */
	uint64_t capacity;      /* rows the row-ID / value buffers were sized for    */
	uint32_t n_streams;     /* k: value bitvectors read                          */
	uint32_t n_launches;    /* kernels launched for this query                   */
	uint64_t delta_entries; /* pending-delta words XOR-ed at query time          */
	uint64_t algo_bytes_scan;  /* k*ceil(N/64)*8 + delta + 8*count   */
	uint64_t algo_bytes_probe; /* count*Σ col widths (+ 8*count when a gather kernel re-reads the row IDs); the
	                              re-read of the merged bitvector by the bit-driven / dense probe is not counted */
	float ms_scan;          /* merge+decode kernel(s), CUDA events (CUBIT_Q_TIMING) */
	float ms_probe;         /* probe kernel                                      */
	float ms_total;         /* first launch → last launch of the query           */
	uint32_t fused;         /* 1 if merge+decode(+probe) ran as one kernel       */
	const int64_t *d_rowids;   /* device pointers (valid until free_result)     */
	const uint64_t *d_bitvector;
	const void *d_values[CUBIT_MAX_PROBE_COLS];
	double sum_f64;         /* CUBIT_AGG_SUM_F64 */
	uint64_t agg_rows;      /* non-NULL inputs of the aggregate (= count when the aggregate columns hold no
	                           NULLs); 0 means the SQL SUM is NULL                                          */
	const uint32_t *d_validity[CUBIT_MAX_PROBE_COLS]; /* device validity bits of the projected values (bit j of
	                           32-bit word j/32 = result row j), NULL = that column holds no NULLs          */
	uint32_t probe_path;    /* CUBIT_PROBE_*: which kernel probed the columns                              */
	uint32_t scan_path;     /* CUBIT_SCAN_*: how merge + decode ran                                         */
} cubit_result_info;

/* cubit_result_info.scan_path */
#define CUBIT_SCAN_RING 0     /* the single-pass ring kernel (fused merge + decode)                          */
#define CUBIT_SCAN_TWO_PASS 1 /* short queries on large tables: streaming merge + count, then a decode pass  */
#define CUBIT_SCAN_NONE 2     /* no scan ran: the probe read the one value bitvector itself                  */
#define CUBIT_SCAN_LOOKBACK 3 /* short queries with row positions on large tables: one pass, decoupled look-back */

/* cubit_result_info.probe_path */
#define CUBIT_PROBE_NONE 0   /* no column was probed                                                        */
#define CUBIT_PROBE_FUSED 1  /* inside the scan kernel (CUBIT_Q_FUSE_PROBE)                                  */
#define CUBIT_PROBE_BITS 2   /* bit-driven gather probe over the merged bitvector                            */
#define CUBIT_PROBE_GATHER 3 /* gather over the row-ID list (sparse selections, NULL-bearing / 4-byte columns) */
#define CUBIT_PROBE_DENSE 4  /* dense selections over bit-packed columns: pack blocks streamed through shared memory */

/* ---- library ---------------------------------------------------------- */
int cubit_gpu_abi_version(void);
const char *cubit_gpu_last_error(void);
int cubit_gpu_device_count(int *count);

/* ---- table shard -------------------------------------------------------
 * n_rows    rows held by this shard (row-range sharding)
 * row_base  global row ID of local row 0 (a non-negative multiple of 64; shards of one table use multiples of seg_bits)
 * seg_bits  CUBIT segment size in rows: 32768, 65536 or 131072.  One segment
 *           is the unit of the merge kernel and of the pending-delta lists.  */
int cubit_gpu_create(int device, uint64_t n_rows, int64_t row_base, uint32_t seg_bits, cubit_gpu_table **out);
/* One table over several GPUs of the box (the reference's parallel unit is the row-group range handed
 * out by RowGroupCollection::NextParallelScan, src/storage/table/row_group_collection.cpp:174-224, bounded by
 * TableFunction MaxThreads, src/include/duckdb/function/table_function.hpp:45-67).  Rows are cut into contiguous
 * ranges of whole segments, shard i on devices[i] (a device may be named more than once; fewer shards are made when
 * the table has fewer segments than devices).  The returned handle is used with EVERY other entry point: uploads,
 * synthetic columns, index builds and deltas are cut or routed by row range; cubit_gpu_query runs on all shards at
 * once; COUNT / SUM are added exactly; cubit_gpu_fetch walks the shards in row order, so row IDs stay globally
 * sorted.  Per-shard operations (WAH upload, on-disk segment upload, index images, set_stream) report
 * CUBIT_ESTATE on the parent — address the shard's own data instead.  row_base must be a multiple of seg_bits. */
int cubit_gpu_create_sharded(const int *devices, uint32_t n_devices, uint64_t n_rows, int64_t row_base,
                             uint32_t seg_bits, cubit_gpu_table **out);
int cubit_gpu_shard_count(const cubit_gpu_table *t, uint32_t *n_shards);
int cubit_gpu_shard_info(const cubit_gpu_table *t, uint32_t shard, int *device, uint64_t *first_row, uint64_t *n_rows);
int cubit_gpu_destroy(cubit_gpu_table *t);
int cubit_gpu_row_count(const cubit_gpu_table *t, uint64_t *n_rows);
/* run this table's kernels on a caller-owned cudaStream_t (NULL = own stream) */
int cubit_gpu_set_stream(cubit_gpu_table *t, void *cuda_stream);
int cubit_gpu_words_per_bitvector(const cubit_gpu_table *t, uint64_t *n_words);
/* kernels launched by this table since creation (bench.py: gpu_launches) */
int cubit_gpu_launch_count(const cubit_gpu_table *t, uint64_t *n);

/* ---- CUBIT index: `cardinality` value bitvectors over the shard's rows -- */
int cubit_gpu_index_create(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id);
/* The same index kept COMPRESSED in HBM: one roaring-style container per (value, segment) —
 * empty, full, a sorted array of ≤ 512 16-bit row positions, or the verbatim segment — which the scan kernel
 * expands in shared memory, so sparse bitvectors cost neither HBM capacity nor bandwidth (a day-level l_shipdate
 * index, 2,526 bitvectors, is 190 GB verbatim at SF100).  Every entry point works on it unchanged: uploads
 * (verbatim or WAH: expanded on the GPU, stored as containers), GPU build from a column, pending deltas (XOR-ed
 * after the expansion), merge-back, append, images.  Needs seg_bits <= 65536. */
int cubit_gpu_index_create_compressed(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id);
typedef struct cubit_index_info {
	uint32_t cardinality;
	uint32_t compressed;     /* 1: containers, 0: verbatim bitvectors                     */
	uint64_t resident_bytes; /* HBM the bitvectors occupy (pool + directory / verbatim)    */
	uint64_t verbatim_bytes; /* cardinality * ceil(n_rows / seg_bits) * seg_bits / 8       */
	uint64_t delta_entries;  /* pending-delta entries (16 bytes each)                      */
	uint64_t auto_merges;    /* threshold-driven merge-backs so far                        */
} cubit_index_info;
int cubit_gpu_index_info(cubit_gpu_table *t, int32_t index_id, cubit_index_info *info);
/* words: ceil(n_rows/64) host words, bits >= n_rows must be 0 */
int cubit_gpu_upload_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const uint64_t *words,
                               uint64_t n_words);
int cubit_gpu_download_bitvector(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *words,
                                 uint64_t n_words);
/* Upload a value bitvector in the WAH-compressed form of FastBit's ibis::bitvector — what the upstream CUBIT
 * library keeps per value (32-bit words; literal: MSB 0 + 31 bits, first bit most significant; fill: MSB 1,
 * bit 30 = fill bit, low 30 bits = number of 31-bit groups; `active` = the trailing < 31 bits).  The compressed
 * words cross PCIe as they are and are expanded on the GPU.  The vector may describe fewer than n_rows bits
 * (the rest is 0, as ibis::bitvector::adjustSize would pad) but not more.  Malformed input (zero-length fill,
 * active_nbits > 30, too long) is rejected on the host. */
typedef struct cubit_wah_bitvector {
	const uint32_t *words;
	uint64_t n_words;
	uint32_t active_val;   /* trailing bits, first bit most significant of the low active_nbits bits */
	uint32_t active_nbits; /* 0..30 */
} cubit_wah_bitvector;
int cubit_gpu_upload_bitvector_wah(cubit_gpu_table *t, int32_t index_id, uint32_t value_id,
                                   const cubit_wah_bitvector *bv);
/* build every bitvector of the index on the GPU from a resident integer
 * column: row r sets bit r of B_(col[r]-base_value); values outside
 * [base_value, base_value+cardinality) are not indexed (NULL keys are not
 * indexed in the reference either: plan_create_index.cpp:60-78). */
int cubit_gpu_index_build(cubit_gpu_table *t, int32_t index_id, int32_t col_id, int64_t base_value);
/* popcount of B_v as stored (pending deltas not applied) */
int cubit_gpu_bitvector_count(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, uint64_t *count);

/* Pending update/delete deltas (the index side of DML: BoundIndex::Append / Delete / Insert,
 * src/include/duckdb/execution/index/bound_index.hpp:71-97).  D_v is the set of LOCAL row positions whose bit is
 * flipped at query time (B_v XOR D_v); a row listed an even number of times cancels.  UPDATE v→w of row r
 * contributes r to D_v and D_w, DELETE of a row whose value is v contributes r to D_v.
 *   cubit_gpu_add_delta        adds n flipped rows to D_v — INCREMENTAL: earlier pending rows stay
 *   cubit_gpu_add_delta_pairs  adds n (value, row) pairs in one call (what one UPDATE / DELETE statement produces)
 *   cubit_gpu_set_delta        replaces D_v
 * Ingestion runs ON THE DEVICE, in stream order behind the scans already enqueued (counting sort by
 * (value, segment): histogram, scan, move, scatter — delta_kernels.cu); the host only copies the pairs across and
 * never waits.  rows / value_ids may be pageable and are free to reuse on return.
 * Merge-back (B_v ^= D_v, lists cleared) happens on demand (cubit_gpu_merge_deltas) and AUTOMATICALLY after an
 * add once the pending entries of a touched value outweigh `fraction` of its bitvector
 * (cubit_gpu_set_merge_threshold, default 0.25; 0 disables).  Needs cardinality * segments <= 2^30. */
int cubit_gpu_add_delta(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const int64_t *rows, uint64_t n);
int cubit_gpu_add_delta_pairs(cubit_gpu_table *t, int32_t index_id, const uint32_t *value_ids, const int64_t *rows,
                              uint64_t n);
int cubit_gpu_set_delta(cubit_gpu_table *t, int32_t index_id, uint32_t value_id, const int64_t *rows, uint64_t n);
int cubit_gpu_merge_deltas(cubit_gpu_table *t, int32_t index_id);
int cubit_gpu_set_merge_threshold(cubit_gpu_table *t, int32_t index_id, double fraction);

/* Persistence of an index: the analog of BoundIndex::GetStorageInfo → IndexStorageInfo
 * (src/include/duckdb/execution/index/bound_index.hpp:117-118), written at checkpoint and read back when the
 * table is attached.  cubit_gpu_index_serialize produces a self-describing byte image of ONE index: header,
 * every value bitvector either verbatim or WAH-compressed (the form cubit_gpu_upload_bitvector_wah expands on
 * the GPU — whichever is smaller), the pending deltas as flipped-row lists (they stay pending after a reload),
 * and a checksum.  *image is malloc'ed by the library: release it with cubit_gpu_free_image.
 * cubit_gpu_index_deserialize validates the image (magic, sizes, row count of THIS table, checksum; CUBIT_EINVAL
 * before anything is uploaded) and recreates the index: *index_id receives the new id. */
int cubit_gpu_index_serialize(cubit_gpu_table *t, int32_t index_id, void **image, uint64_t *bytes);
int cubit_gpu_index_deserialize(cubit_gpu_table *t, const void *image, uint64_t bytes, int32_t *index_id);
void cubit_gpu_free_image(void *image);

/* ---- columns (decoded, fixed width 4 or 8 bytes, HBM resident) ---------- */
int cubit_gpu_upload_column(cubit_gpu_table *t, int32_t col_id, const void *data, uint32_t elem_bytes, uint64_t n);
int cubit_gpu_download_column(cubit_gpu_table *t, int32_t col_id, void *data, uint32_t elem_bytes, uint64_t n);
/* NULLs: the column's validity mask in the reference's layout (ValidityMask: 64-bit words, bit r%64 of word r/64
 * = 1 when row r is valid, src/include/duckdb/common/types/validity_mask.hpp:50,163-168 — also what a
 * validity_uncompressed segment stores, src/storage/compression/validity_uncompressed.cpp:381).  n_words must be
 * ceil(n_rows / 64); words = NULL drops the mask.
 * Upload it after the column data (cubit_gpu_upload_column resets a column to all-valid).  Probes report the
 * validity of every projected value (cubit_gpu_fetch_validity, StandardColumnData::FetchRow =
 * validity.FetchRow + data, standard_column_data.cpp:169-178); SUM skips NULL inputs, SUM(a*b) skips rows where
 * either factor is NULL, `count` stays COUNT(*) and cubit_result_info.agg_rows counts the non-NULL inputs.
 * Appended rows are valid. */
int cubit_gpu_upload_column_validity(cubit_gpu_table *t, int32_t col_id, const uint64_t *words, uint64_t n_words);
/* Synthetic columns generated on the device (bench / parity-test support;
 * the generators are restated in oracle/cubit_oracle.c):
 *   kind 0: int64 payload, value = row_base + r
 *   kind 1: int32 value column of SURVEY §8d config 2: z = splitmix64(seed + row_base + r);
 *           z < threshold ? hot_lo + (z>>7) % hot_n : the (z>>7) % (card-hot_n)-th value outside the hot range
 *   kind 2: int32 uniform in [hot_lo, hot_lo+card):      hot_lo + (z>>7) % card
 *   kind 3: int64 uniform in [hot_lo, hot_lo+threshold): hot_lo + (z>>7) % threshold */
int cubit_gpu_synth_column(cubit_gpu_table *t, int32_t col_id, int32_t kind, uint64_t seed, uint64_t threshold,
                           uint32_t card, uint32_t hot_lo, uint32_t hot_n);
int cubit_gpu_drop_column(cubit_gpu_table *t, int32_t col_id);
/* Store an 8-byte column FOR-bit-packed in HBM (lossless): blocks of 1024 rows, per block min + bit width of
 * (max - min) — the resident analog of the reference's BitPacking column storage.  Selections of >= 1/28 of the
 * rows over columns packed to <= 32 bits are probed by streaming the pack blocks through shared memory (the dense
 * probe: width/8 bytes per row instead of 128 bytes of DRAM per gathered value); sparser ones gather from whichever
 * form is cheaper.  keep_raw = 1 keeps both forms (the planner picks per query); keep_raw = 2 does the same but
 * leaves the column raw (*packed_bytes = 0) when a block needs more than 32 bits, i.e. when the dense probe could
 * not use the packed form; keep_raw = 0 frees the raw array (index builds and appends need it: build first).
 * *packed_bytes = resident size of the packed form. */
int cubit_gpu_pack_column(cubit_gpu_table *t, int32_t col_id, int keep_raw, uint64_t *packed_bytes);

/* Append n_new rows at the end of the shard (INSERT: the new rows take the next row ids — rowids are dense
 * table positions, src/storage/table/row_group.cpp:511-514; index side BoundIndex::Append,
 * src/include/duckdb/execution/index/bound_index.hpp:71-75).  Every resident column must be supplied (raw
 * form: a column that is resident ONLY bit-packed is rejected with CUBIT_ESTATE; one that keeps both forms loses
 * its packed form, which no longer covers the table — cubit_gpu_pack_column rebuilds it).  Indexes built with cubit_gpu_index_build are
 * extended from their source column on the GPU (only the new rows are scanned); bitvectors that were uploaded
 * get zero bits for the new rows.  Pending deltas stay pending.  cubit_gpu_words_per_bitvector changes. */
typedef struct cubit_append_column {
	int32_t col_id;
	uint32_t elem_bytes;
	const void *data; /* n_new elements */
} cubit_append_column;
int cubit_gpu_append_rows(cubit_gpu_table *t, uint64_t n_new, const cubit_append_column *cols, uint32_t n_cols);

/* Upload a column as the reference's ON-DISK column segments and decode them on the GPU:
 * the compressed bytes cross PCIe as stored, one CTA decodes one 2048-value metadata group.  What each segment
 * is — compression, first row, row count, bytes at (block_id, block_offset) — is what pragma_storage_info /
 * ColumnSegment report (src/storage/table/column_segment.cpp, DataPointer in src/include/duckdb/storage/
 * data_pointer.hpp); the layout decoded is src/storage/compression/bitpacking.cpp:22-75,524-544,660-860.
 * Segments must be sorted by row_start and tile [0, n_rows) exactly.  Malformed segments are rejected on the
 * host (CUBIT_EINVAL) before anything is launched.  The result is the same resident column
 * cubit_gpu_upload_column produces. */
#define CUBIT_SEG_UNCOMPRESSED 0 /* plain array of count elements (fixed_size_uncompressed.cpp)             */
#define CUBIT_SEG_BITPACKING 1   /* BitPacking segment: u64 metadata-end offset, group data, metadata words */
#define CUBIT_SEG_CONSTANT 2     /* Constant compression: data points to the one value of the segment       */
#define CUBIT_SEG_RLE 3          /* RLE segment (src/storage/compression/rle.cpp:190-205): u64 offset of the run
                                    lengths, run values, padding, u16 run lengths.  The format does not store its
                                    size: `bytes` is an upper bound (e.g. up to the end of the block); the runs
                                    needed to cover `count` rows are validated against it                    */
typedef struct cubit_column_segment {
	uint32_t kind;      /* CUBIT_SEG_*                                          */
	uint32_t reserved;
	uint64_t row_start; /* first row of the segment, local to this table shard  */
	uint64_t count;     /* rows in the segment                                  */
	const void *data;   /* the segment's bytes as stored in its block           */
	uint64_t bytes;
} cubit_column_segment;
typedef struct cubit_decode_info {
	uint64_t h2d_bytes;  /* compressed bytes copied host → device                */
	uint64_t n_groups;   /* metadata groups decoded                               */
	uint64_t mode_groups[6]; /* groups per BitpackingMode (bitpacking.hpp:15)     */
	uint64_t rle_runs;   /* runs decoded from RLE segments                        */
	uint32_t n_launches;
	float ms_decode;     /* decode kernel, CUDA events                            */
} cubit_decode_info;
int cubit_gpu_upload_column_segments(cubit_gpu_table *t, int32_t col_id, uint32_t elem_bytes,
                                     const cubit_column_segment *segs, uint32_t n_segs, cubit_decode_info *info);

/* ---- query --------------------------------------------------------------
 * Synchronous unless CUBIT_Q_ASYNC: on return info fields are final.
 * Thread-safe per table: the table's lock is held only while a query is planned and its kernels are enqueued,
 * never while a caller waits for the GPU or copies rows out, so several host threads keep queries and DataChunk
 * hand-offs in flight on one table (kernels of one shard run in order on its kernel stream; hand-off copies run
 * on separate copy streams).                                                */
int cubit_gpu_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out);
int cubit_gpu_result_wait(cubit_gpu_result *r);
int cubit_gpu_result_get(cubit_gpu_result *r, cubit_result_info *info);
/* copy result rows [offset, offset+n) to host: row IDs and/or the which-th
 * projected column (pass NULL to skip one).  This is what fills a DataChunk
 * (≤ 2048 rows per GetData call: table_scan.cpp:258-268). */
int cubit_gpu_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                    void *const *host_cols);
/* The same copy without the wait: window i+1 crosses PCIe while the caller consumes window i (ordered parallel
 * sources keep several in flight, the pattern of src/function/table/table_scan.cpp:179-189).  Buffers should come
 * from cubit_gpu_alloc_host (a pageable destination makes the copy synchronous) and must stay untouched until
 * cubit_gpu_fetch_wait, which completes and frees the ticket.  May be called from several threads at once. */
int cubit_gpu_fetch_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                          void *const *host_cols, cubit_gpu_fetch_ticket **ticket);
int cubit_gpu_fetch_wait(cubit_gpu_fetch_ticket *ticket);
/* ---- narrow-wire hand-off (include/cubit_gpu_wire.h has the format and the inline unpacker) -----------------
 * Result rows [offset, offset+n) as ONE wire: per DataChunk (2048 rows) and stream an int64 base + deltas of the
 * narrowest width that holds the chunk's range, written by the GPU straight into `host_wire` (zero-copy stores, so
 * only the narrowed bytes cross PCIe: 2-6 bytes per row and stream on sorted row IDs and FOR-friendly columns
 * instead of 8).  with_rowids != 0 puts the row IDs in stream 0; then the first n_cols projected columns follow.
 * host_wire must be page-locked (cubit_gpu_alloc_host / cudaHostAlloc / cudaHostRegister — CUBIT_EINVAL otherwise:
 * the device writes it directly) and hold cubit_wire_bytes(n, streams) bytes; one wire holds at most 32768 frames
 * (chunks x streams: a window, not a whole result — CUBIT_EINVAL beyond); it must stay untouched until
 * cubit_gpu_fetch_wait(ticket).  Thread-safe like cubit_gpu_fetch_async.  On a sharded result the shard that holds the
 * whole window writes the wire; a window that straddles two shards has no single device to write it from:
 * CUBIT_ESTATE — fetch that window with cubit_gpu_fetch_async, or use cubit_gpu_drain (which cuts its windows at the
 * shard boundaries). */
int cubit_gpu_fetch_wire_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int with_rowids, uint32_t n_cols,
                               void *host_wire, uint64_t host_wire_bytes, cubit_gpu_fetch_ticket **ticket);
/* exported copies of the inline helpers of cubit_gpu_wire.h (for bindings that cannot include C) */
uint64_t cubit_gpu_wire_bytes(uint64_t n_rows, uint32_t n_streams);
uint64_t cubit_gpu_wire_payload_bytes(const void *host_wire);
int cubit_gpu_wire_unpack(const void *host_wire, uint32_t stream, uint64_t chunk, void *out, uint32_t out_elem);

/* The whole parallel ordered hand-off in one call (the source side of PhysicalTableScan::GetData with
 * MaxThreads() > 1 and get_batch_index, table_function.hpp:45-67, table_scan.cpp:179-189): n_threads workers claim
 * windows of window_rows result rows (a multiple of 2048; 0 = default), keep two narrow-wire fetches in flight each,
 * widen one DataChunk at a time into worker-local vectors and hand it to fn together with its batch index (= the
 * window index: chunks of one window arrive in order on one worker; sort by (batch_index, row_offset) to restore
 * the global order).  cols[c] points to n values of the c-th projected column (4- or 8-byte elements as the query
 * returns them), validity[c] is NULL or the ValidityMask words of the chunk; rowids is NULL unless with_rowids.
 * A non-zero return of fn stops the drain (CUBIT_ESTATE); fn runs on the library's worker threads and must not throw
 * or unwind through them.  fn == NULL is the built-in checksum consumer: every
 * chunk is widened and summed (wrapping 64-bit sums of the row IDs and of every column's bit patterns widened to
 * 64 bits) — a consumer that reads every delivered value, for tests and bench.py.  Works on sharded results. */
typedef int (*cubit_chunk_fn)(void *ctx, uint32_t worker, uint64_t batch_index, uint64_t row_offset, uint32_t n,
                              const int64_t *rowids, const void *const *cols, const uint64_t *const *validity);
typedef struct cubit_drain_stats {
	uint64_t rows;       /* rows delivered                                        */
	uint64_t chunks;     /* DataChunks delivered                                  */
	uint64_t windows;
	uint64_t wire_bytes; /* bytes that crossed PCIe (directories + written slots) */
	uint64_t wide_bytes; /* what the wide hand-off would have moved               */
	uint64_t sum_rowids; /* checksum consumer: wrapping sums                      */
	uint64_t sum_cols[CUBIT_MAX_PROBE_COLS];
	uint32_t workers;
	uint32_t reserved;
} cubit_drain_stats;
int cubit_gpu_drain(cubit_gpu_result *r, int with_rowids, uint32_t n_cols, uint32_t n_threads, uint64_t window_rows,
                    cubit_chunk_fn fn, void *ctx, cubit_drain_stats *stats);

/* Multi-process reduce (one rank per GPU): ADD this result's (count, 128-bit sum) to device_dst[0..5)
 * as five int64 limbs — count, sum bits [0,32), [32,64), [64,96), signed [96,128) — on the device, in stream
 * order behind the query, so the caller can hand device_dst to one ncclAllReduce(sum) without the aggregates
 * visiting the host.  The limbs of up to 2^31 results can be accumulated before a carry could be lost. */
int cubit_gpu_result_add_limbs(cubit_gpu_result *r, int64_t *device_dst);
/* Validity mask of projected column `col` (index into the query's cols[]) for result rows [offset, offset+n):
 * bit j of host_words = row offset+j, ceil(n/64) words, bits past n zero — the mask the DataChunk vector gets
 * (FlatVector::Validity).  *all_valid (optional) = 1 when no NULL falls into the range, so the caller can skip
 * installing a mask; host_words may be NULL when only that flag is wanted. */
int cubit_gpu_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n, uint64_t *host_words,
                             int *all_valid);
int cubit_gpu_fetch_bitvector(cubit_gpu_result *r, uint64_t *host_words, uint64_t n_words);
/* Page-locked host memory for the buffers cubit_gpu_fetch fills (the staging window behind GetData): a copy
 * into pageable memory runs at a fraction of the PCIe rate.  Plain malloc'ed buffers keep working. */
int cubit_gpu_alloc_host(uint64_t bytes, void **ptr);
int cubit_gpu_free_host(void *ptr);
int cubit_gpu_free_result(cubit_gpu_result *r);

/* Probe a resident column at caller-supplied sorted row IDs (the
 * DataTable::Fetch analog, src/storage/data_table.cpp:373-377): gathers
 * col[row_ids[i] - row_base] into host_out and optionally sums it. */
int cubit_gpu_probe(cubit_gpu_table *t, int32_t col_id, const int64_t *host_rowids, uint64_t n, void *host_out,
                    uint64_t *sum_lo, int64_t *sum_hi);

#ifdef __cplusplus
}
#endif
#endif /* CUBIT_GPU_H */
