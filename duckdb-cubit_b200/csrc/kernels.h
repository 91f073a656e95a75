// kernels.h — internal launch interface between the C-ABI host code
// (cubit_gpu.cu) and the sm_100a kernels.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cubit {

constexpr int kMaxStreams = 64;       // == CUBIT_MAX_STREAMS
constexpr int kMaxFusedCols = 2;      // value columns the fused scan kernel can probe (int64 only)
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kScanThreads = kConsumerThreads + 64; // + one producer warp + one prefix (look-back) warp
constexpr int kScanRingBytes = 72 * 1024;           // bulk-copy ring per CTA (stages of one segment each); 2 CTAs / SM

// One pending-delta entry of one (bitvector, segment): XOR `mask` into word `word` of the staged segment.
// Entries of one (bitvector, segment) are contiguous in the index's delta CSR but unordered, and a word may
// appear more than once (the kernels XOR atomically), so ingestion never sorts or deduplicates.
// `pad` carries the entry's CSR key (value * n_seg + segment) for the device-side rebuilds (delta_kernels.cu).
struct DeltaEnt {
	uint32_t word;
	uint32_t pad;
	uint64_t mask;
};

// one device-side ingestion of pending deltas (delta_kernels.cu): old CSR + n_new (value, row) pairs → new CSR
struct DeltaIngest {
	const long long *rows;   // [n_new] local row positions (device)
	const uint32_t *values;  // [n_new] value ids (device), or nullptr: every pair belongs to `one_value`
	uint32_t one_value;
	uint64_t n_new;
	uint32_t n_seg;
	uint32_t seg_shift;      // log2(seg_bits)
	uint64_t n_keys;         // card * n_seg
	const uint32_t *old_off; // [n_keys + 1] or nullptr (no pending deltas yet)
	const DeltaEnt *old_ent;
	uint64_t n_old;
	uint32_t drop_value;     // entries of this value are NOT carried over (cubit_gpu_set_delta), 0xffffffff = none
	uint32_t *cnt;           // [n_keys] scratch, zeroed
	uint32_t *block_sum;     // [ceil(n_keys / 4096)] scratch
	uint32_t *new_off;       // [n_keys + 1] out
	DeltaEnt *new_ent;       // [surviving old + n_new] out
};

// ---- roaring-style containers of a compressed index (container_kernels.cu) -------------------------
// directory entry of one (value, segment): type (bits 1:0) | set bits of the segment (bits 18:2, 0..65536) | pool
// offset in 16-byte units (bits 63:19).  EMPTY = all-zero word, so a zeroed directory is an empty index.
enum ContainerType : uint32_t { CT_EMPTY = 0, CT_FULL = 1, CT_ARRAY = 2, CT_BITMAP = 3 };
constexpr int kArrayMax = 512; // ARRAY containers hold ≤ 512 sorted 16-bit positions (1 KiB staged per ring stage)
__host__ __device__ inline uint32_t ct_type(unsigned long long d) {
	return (uint32_t)(d & 3ull);
}
__host__ __device__ inline uint32_t ct_count(unsigned long long d) {
	return (uint32_t)((d >> 2) & 0x1ffffull);
}
__host__ __device__ inline unsigned long long ct_offset(unsigned long long d) {
	return (d >> 19) << 4; // bytes
}
__host__ __device__ inline unsigned long long ct_make(uint32_t type, uint32_t count, unsigned long long byte_off) {
	return (unsigned long long)type | ((unsigned long long)count << 2) | ((byte_off >> 4) << 19);
}

// ---- column storage -----------------------------------------------------------------------------
// An 8-byte column lives in HBM either raw (int64 per row) or FOR-bit-packed: blocks of kPackBlock rows,
// each with a 16-byte header {base = min of the block, width = bits of (max - min), word_off = first
// 64-bit word of the block's payload}.  Value i of a block sits at bit i*width of that payload (exactly
// 16*width words per block).  Lossless; the probe decodes in registers, so a dense probe reads width/8
// bytes per row instead of 8 (the on-disk BitPacking idea of the reference, src/storage/compression/
// bitpacking.cpp:22-75, restated with a GPU-friendly fixed block size — not its on-disk layout).
constexpr int kPackBlock = 1024;
struct PackHdr {
	long long base;
	uint32_t word_off;
	uint32_t width;
};
struct ColRef {
	const long long *raw;              // raw int64 column, or nullptr when packed
	const unsigned long long *words;   // packed payload (one spare word at the end)
	const PackHdr *hdr;                // one header per kPackBlock rows
};

// ---- decode of the reference's on-disk column segments (column_decode.cu) -------------------------
// BitpackingMode, src/include/duckdb/storage/compression/bitpacking.hpp:15
enum BpMode : uint32_t { BP_INVALID = 0, BP_AUTO = 1, BP_CONSTANT = 2, BP_CONSTANT_DELTA = 3, BP_DELTA_FOR = 4, BP_FOR = 5 };
// one 2048-value metadata group of a BitPacking segment (or a whole Constant segment), validated on the host
struct BpGroup {
	uint64_t data_off; // byte offset in the staged blob of the group's header (4-byte aligned)
	uint64_t row0;     // first local row the group decodes to
	uint32_t n;        // values to produce (≤ 2048 except BP_CONSTANT)
	uint32_t mode;     // BpMode
};

// one tile of an RLE column segment (src/storage/compression/rle.cpp:190-205: [u64 offset of the run lengths]
// [T values[n_runs]] pad [u16 run lengths[n_runs]]): ≤ kRleTileRuns consecutive runs producing ≤ ~128 K rows,
// cut and validated on the host
constexpr int kRleTileRuns = 1024;
struct RleTile {
	uint64_t val_off; // byte offset in the staged blob of the tile's first run value (element aligned)
	uint64_t cnt_off; // byte offset of the tile's first run length (uint16)
	uint64_t row0;    // first local row the tile decodes to
	uint32_t n_runs;
	uint32_t n_rows;  // rows the tile produces (the segment's last run may be cut by the row count)
};

// Device-side result header (one per query).
struct ResultHeader {
	unsigned long long count;
	unsigned long long sum_lo;
	long long sum_hi;
	unsigned int overflow; // int64 product overflow seen (CUBIT_AGG_SUM_PROD)
	unsigned int pad;
	double sum_f64;        // CUBIT_AGG_SUM_F64
	unsigned long long agg_rows; // non-NULL inputs of the aggregate (written only when a validity mask is involved)
};

struct BlockPartial {
	unsigned long long count;
	unsigned long long sum_lo;
	long long sum_hi;
	unsigned long long pad;
};

// Per-query control block, zeroed (cudaMemsetAsync) before every scan launch.
//   [0]            ticket counter (u32) | blocks-done counter (u32)
//   [1 .. n_seg]   decoupled look-back status word of every segment
struct ScanArgs {
	const uint64_t *bv[kMaxStreams];    // value bitvector B_i (padded to whole segments)
	const uint32_t *doff[kMaxStreams];  // delta CSR offsets of D_i (the index's CSR at key value*n_seg), or nullptr
	const DeltaEnt *dent[kMaxStreams];  // delta entries of D_i's index (offsets are absolute)
	const unsigned long long *cdir[kMaxStreams]; // container directory of B_i (compressed index; bv[i] = its pool), or nullptr
	uint64_t group_end;                 // bit i: stream i closes its OR group (then Q &= group)
	uint32_t k;                         // streams
	uint32_t n_seg;                     // segments (tiles)
	uint32_t ticket_depth;              // segment tickets a CTA keeps in flight (set by launch_scan)
	int64_t row_base;                   // global row ID of local row 0
	unsigned long long *ctrl;           // control block (see above)
	uint64_t *q_out;                    // merged bitvector out, or nullptr
	unsigned long long *tile_excl;      // per-segment exclusive prefix out (scan) / in (bit-driven probe), or nullptr
	unsigned long long *span_excl;      // scan only: exclusive prefix of every consumer warp's span ([n_seg * kConsumerWarps],
	                                    // written at emission) for the dense probe (probe_dense_kernel.cu), or nullptr
	long long *ids_out;                 // sorted row IDs out, or nullptr
	// fused probe: the distinct int64 columns read at every selected row
	int n_load;                           // 0..kMaxFusedCols
	ColRef lcol[kMaxFusedCols];           // the columns (local row indexed; raw or bit-packed)
	long long *lout[kMaxFusedCols];       // gathered values out (same positions as ids_out), or nullptr
	int agg_kind;                         // CUBIT_AGG_*: SUM(lcol[agg_ia]) / SUM(lcol[agg_ia]*lcol[agg_ib])
	int agg_ia;
	int agg_ib;
	BlockPartial *partials;             // [gridDim.x]
	ResultHeader *hdr;
	int skip_count;                     // 1: do not add this launch's popcounts to hdr->count (decode pass of UNFUSED)
	int count_rows;                     // bit-driven probe only: 1 = it also counts the set bits into hdr->count (no scan ran)
	unsigned int debug;                 // timing experiments only (CUBIT_SCAN_DEBUG): results are WRONG when != 0
};

// Dense probe over bit-packed columns (probe_dense_kernel.cu): streams the pack blocks of every span that has a
// selected row through per-warp shared-memory stages and decodes there.  A span = one scan-kernel consumer warp's
// share of a segment = seg_words / 8 words of Q = seg_bits / 8192 pack blocks.
struct DenseProbeArgs {
	const uint64_t *q;                   // merged bitvector (whole segments)
	const unsigned long long *span_excl; // [n_span] output position of every span's first selected row (positions only)
	uint32_t n_span;                     // n_seg * kConsumerWarps
	uint64_t n_blk;                      // pack blocks of the columns (ceil(n_rows / kPackBlock))
	int n_load;                          // 1..kMaxFusedCols distinct columns
	ColRef lcol[kMaxFusedCols];          // bit-packed (widths <= 32) or raw
	long long *lout[kMaxFusedCols];      // gathered values out (same positions as the row IDs), or nullptr
	int agg_kind, agg_ia, agg_ib;        // as in ScanArgs
	ResultHeader *hdr;                   // sums are ADDED
	uint32_t stage_bytes[kMaxFusedCols]; // bytes of one shared-memory stage of every column (0: raw) — dense_probe_plan
	uint32_t warp_bytes;                 // shared memory per warp
	uint32_t n_stages;                   // ring stages per column (2..4)
};

// Short queries as two streaming passes (small_scan_kernels.cu).  Unit = 128 words = 8192 rows; every warp of the
// grid owns a contiguous chunk of units.
struct SmallScanArgs {
	const uint64_t *bv[8];          // value bitvectors (k <= 8 here; the planner uses it for k <= 4)
	uint64_t group_end;             // bit i: stream i closes its OR group
	uint32_t k;
	uint32_t n_units, units_per_chunk, n_chunks; // small_scan_plan
	uint64_t *q_out;                // pass A: merged bitvector out, or nullptr (k = 1: Q is bv[0] itself)
	const uint64_t *q_in;           // pass C: the bitvector to decode
	unsigned long long *chunk_tot;  // [n_chunks + 1]: pass A totals → pass B exclusive prefixes (+ total)
	unsigned long long *span_excl;  // pass C out: output position of every unit's first row, or nullptr
	unsigned long long *tile_excl;  // pass C out: the same per 8 units (65536-row tile), or nullptr
	ResultHeader *hdr;
	int count_here;                 // 1: pass A adds COUNT to hdr (no pass B / C follows)
};
cudaError_t launch_small_merge_count(const SmallScanArgs &s, cudaStream_t stream);
cudaError_t launch_small_prefix(const SmallScanArgs &s, cudaStream_t stream);
cudaError_t launch_small_decode(const SmallScanArgs &s, const ScanArgs &a, cudaStream_t stream);
void small_scan_plan(SmallScanArgs &s, uint32_t n_units, int sm_count);
// the same merge + decode in ONE pass with a decoupled look-back (lookback_scan_kernel.cu): a.ctrl = 1 + n_tiles zeroed
// words, a.ids_out / a.row_base as for launch_small_decode; writes hdr->count, span_excl / tile_excl, q_out
uint32_t lookback_scan_tiles(uint32_t n_units);
cudaError_t launch_lookback_scan(const SmallScanArgs &s, const ScanArgs &a, int sm_count, cudaStream_t stream);

struct ProbeArgs {
	const long long *ids;               // sorted global row IDs
	const unsigned long long *count_ptr; // device count (hdr->count) or nullptr
	unsigned long long n;               // used when count_ptr == nullptr
	int64_t row_base;
	int n_cols;
	const void *col[8];                 // raw column (4- or 8-byte), or nullptr when packed[c] is used
	ColRef packed[8];                   // 8-byte columns stored bit-packed
	void *out[8];
	uint32_t elem_bytes[8];
	int agg_kind;
	ColRef agg_a;
	ColRef agg_b;
	// validity masks (DuckDB ValidityMask layout, bit = 1: valid) of the aggregate inputs, or nullptr = no NULLs;
	// rows whose input is NULL are skipped (SUM ignores NULLs; a product with a NULL factor is NULL)
	const unsigned long long *agg_valid_a;
	const unsigned long long *agg_valid_b;
	BlockPartial *partials;
	unsigned int *done;                 // blocks-done counter (zeroed)
	ResultHeader *hdr;                  // count is left untouched; sums written
};

// ---- launchers (all asynchronous on `stream`; return cudaGetLastError()) ----
// seg_words ∈ {512, 1024, 2048} (= 256 consumer threads × 2/4/8 words).  has_delta: any doff[i] != nullptr.
// compressed: any cdir[i] != nullptr (seg_words ≤ 1024, no fused probe).
cudaError_t launch_scan(const ScanArgs &args, uint32_t seg_words, bool has_delta, bool compressed, int sm_count,
                        cudaStream_t stream, int *grid_out);
int scan_max_grid(uint32_t seg_words, int sm_count);
// Bit-driven probe: re-decodes the merged bitvector args.q_out (input here) with the
// per-segment prefixes args.tile_excl and gathers / sums the fused-probe columns at full
// occupancy.  Sums are ADDED to args.hdr (which the scan kernel has already finalised).
cudaError_t launch_probe_bits(const ScanArgs &args, uint32_t seg_words, bool positions, int sm_count,
                              cudaStream_t stream);

// fills stage_bytes / warp_bytes from the columns' widest blocks; false: not eligible (no packed column, or a width > 32)
bool dense_probe_plan(DenseProbeArgs &args, const uint32_t *max_width);
cudaError_t launch_probe_dense(const DenseProbeArgs &args, uint32_t seg_words, bool positions, int sm_count,
                               cudaStream_t stream);

cudaError_t launch_probe(const ProbeArgs &args, int sm_count, cudaStream_t stream);
// validity of a probed column at the selected rows (ValidityFetchRow analog, validity_uncompressed.cpp:381):
// bit j of out32 = valid[ids[j] - row_base] for j < *count_ptr (out32 zeroed beforehand, ceil(cap/32) words)
cudaError_t launch_validity_gather(const long long *ids, const unsigned long long *count_ptr, int64_t row_base,
                                   const unsigned long long *valid, uint32_t *out32, int sm_count,
                                   cudaStream_t stream);
int probe_grid(int sm_count);

// index build: B_(col[r]-base) |= bit r for rows [row_begin, n_rows); bits of rows < row_begin are kept (append)
// valid: the column's validity mask or nullptr — NULL keys are not indexed (plan_create_index.cpp:60-78)
cudaError_t launch_index_build(const void *col, uint32_t elem_bytes, const unsigned long long *valid, uint64_t row_begin,
                               uint64_t n_rows, int64_t base_value,
                               uint32_t cardinality, uint64_t *bitvectors, uint64_t words_per_bv, int sm_count,
                               cudaStream_t stream, int *n_launches);
cudaError_t launch_popcount(const uint64_t *words, uint64_t n_words, unsigned long long *out, int sm_count,
                            cudaStream_t stream);
cudaError_t launch_popcount_many(const uint64_t *bitvectors, uint64_t words_per_bv, uint32_t n_bv,
                                 unsigned long long *out, cudaStream_t stream);
// pending-delta maintenance (delta_kernels.cu)
cudaError_t launch_delta_ingest(const DeltaIngest &a, int sm_count, cudaStream_t stream, int *n_launches);
// B ^= D for entries [e0, e1) of the CSR; bits = B_(value_base)'s first word
cudaError_t launch_delta_apply(const DeltaEnt *ent, uint64_t e0, uint64_t e1, uint32_t n_seg, uint32_t seg_words,
                               uint64_t *bits, uint64_t words_per_bv, uint32_t value_base, int sm_count,
                               cudaStream_t stream);
// out[v] = off[v * n_seg] for v in [0, card] — the per-value entry ranges of the CSR
cudaError_t launch_delta_value_offsets(const uint32_t *off, uint32_t n_seg, uint32_t card, uint32_t *out, cudaStream_t stream);
cudaError_t launch_delta_restride(const uint32_t *old_off, uint32_t *new_off, DeltaEnt *ent, uint64_t n_ent, uint32_t card,
                                  uint32_t old_n_seg, uint32_t new_n_seg, int sm_count, cudaStream_t stream);
// compressed indexes (container_kernels.cu)
// src: [nv][words_per_bv] verbatim → containers appended to the pool at *cursor (cursor[0], bytes; cursor[1]
// accumulates the bytes of the containers that are replaced), directory rows dir[v * dir_stride + seg]
cudaError_t launch_container_compress(const uint64_t *src, uint64_t words_per_bv, uint32_t nv, uint32_t n_seg,
                                      uint32_t seg_words, unsigned long long *dir, uint64_t dir_stride, uint8_t *pool,
                                      unsigned long long *cursor, cudaStream_t stream);
// one value: containers → verbatim words (segments [n_seg, n_seg_alloc) are zeroed)
cudaError_t launch_container_expand(const unsigned long long *dir, const uint8_t *pool, uint64_t n_seg_alloc, uint32_t n_seg,
                                    uint32_t seg_words, uint64_t *dst, cudaStream_t stream);
// out[v] = popcount of B_v (the directory stores every container's count)
cudaError_t launch_compressed_counts(const unsigned long long *dir, uint64_t dir_stride, uint32_t n_seg, uint32_t card,
                                     const uint8_t *pool, uint32_t seg_words, unsigned long long *out, cudaStream_t stream);
// (count, 128-bit sum) of a result header → five 32-bit limbs ADDED to dst[0..5) as int64 (multi-process reduce)
cudaError_t launch_add_limbs(const ResultHeader *hdr, long long *dst, cudaStream_t stream);
// FOR-bit-packing of an int64 column (see ColRef): pass 1 per-block min/width, pass 2 pack
cudaError_t launch_pack_widths(const long long *col, uint64_t n_rows, long long *base_out, uint32_t *width_out,
                               cudaStream_t stream);
cudaError_t launch_pack_blocks(const long long *col, uint64_t n_rows, const PackHdr *hdr, unsigned long long *words,
                               cudaStream_t stream);
// one CTA per BpGroup: packed words staged in shared memory, FOR / DELTA_FOR (block-wide running sum) decode
cudaError_t launch_bp_decode(const uint8_t *blob, const BpGroup *groups, uint32_t n_groups, void *out,
                             uint32_t elem_bytes, cudaStream_t stream);
// one CTA per RleTile: block scan of the run lengths in shared memory, then every output row finds its run by
// binary search (coalesced stores)
cudaError_t launch_rle_decode(const uint8_t *blob, const RleTile *tiles, uint32_t n_tiles, void *out, uint32_t elem_bytes,
                              cudaStream_t stream);
// WAH (FastBit ibis::bitvector) → verbatim bitvector; `out` must be zeroed; block_group0[b] = number of 31-bit
// groups before WAH word b * kWahBlockWords (host-computed while validating)
constexpr int kWahBlockWords = 1024;
cudaError_t launch_wah_expand(const uint32_t *wah, uint64_t n_wah, const unsigned long long *block_group0,
                              unsigned long long total_groups, uint32_t active_val, uint32_t active_nbits,
                              unsigned long long *out, cudaStream_t stream);
cudaError_t launch_synth_column(void *col, int kind, uint64_t n_rows, int64_t row_base, uint64_t seed,
                                uint64_t threshold, uint32_t card, uint32_t hot_lo, uint32_t hot_n, int sm_count,
                                cudaStream_t stream);

} // namespace cubit
