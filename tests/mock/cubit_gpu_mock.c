/* cubit_gpu_mock.c — TEST INFRASTRUCTURE ONLY.
 * A CPU stand-in for the subset of include/cubit_gpu.h that integration/duckdb_cubit_extension.cpp
 * calls, implemented with the ORACLE (oracle/cubit_oracle.c).  It exists so that the DuckDB-side glue can
 * be exercised end-to-end through real reference SQL in the GPU-less build container
 * (tests/test_duckdb_integration.py).  It is never built into, linked with or loaded by the product. */
#include "cubit_gpu.h"
#include "cubit_gpu_wire.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void oracle_merge(const uint64_t *const *streams, const uint64_t *const *deltas, const int32_t *group_of, int k,
                  uint64_t n_words, uint64_t *q);
uint64_t oracle_decode(const uint64_t *q, uint64_t n_words, int64_t row_base, int64_t *out);
uint64_t oracle_popcount(const uint64_t *q, uint64_t n_words);
void oracle_probe(const int64_t *ids, uint64_t n, int64_t row_base, const void *col, uint32_t elem_bytes, void *out);
void oracle_sum_i64(const int64_t *vals, uint64_t n, uint64_t *sum_lo, int64_t *sum_hi);
void oracle_build_index(const void *col, uint32_t elem_bytes, uint64_t n_rows, int64_t base_value, uint32_t card,
                        uint64_t *bitvectors, uint64_t n_words);

uint64_t oracle_probe_validity(const int64_t *ids, uint64_t n, int64_t row_base, const uint64_t *valid,
                               uint64_t *out_words);
uint64_t oracle_sum_nulls(const int64_t *ids, uint64_t n, int64_t row_base, const int64_t *a, const uint64_t *valid_a,
                          const int64_t *b, const uint64_t *valid_b, uint64_t *sum_lo, int64_t *sum_hi, int *ovf);

#define MAX_COLS 64
struct cubit_gpu_table {
	uint64_t n_rows, n_words;
	int64_t row_base;
	int64_t *cols[MAX_COLS];
	uint64_t *valid[MAX_COLS];
#define MAX_INDEXES 8
	uint64_t *bits[MAX_INDEXES];
	uint64_t *dbits[MAX_INDEXES]; /* pending deltas, dense (the mock has no reason to be sparse), NULL = none */
	uint32_t card[MAX_INDEXES];
	int32_t src_col[MAX_INDEXES]; /* column the index was built from (-1: uploaded / source dropped) */
	int64_t src_base[MAX_INDEXES];
	int n_indexes;
};
struct cubit_gpu_fetch_ticket {
	int dummy;
};
struct cubit_gpu_result {
	struct cubit_gpu_table *t;
	uint64_t count;
	int64_t *ids;
	int64_t *vals[CUBIT_MAX_PROBE_COLS];
	uint64_t *vmask[CUBIT_MAX_PROBE_COLS]; /* validity over result positions, NULL = all valid */
	uint64_t agg_rows;
	uint32_t n_cols;
	uint64_t sum_lo;
	int64_t sum_hi;
};
static char g_err[256] = "";

int cubit_gpu_abi_version(void) { return CUBIT_GPU_ABI_VERSION; }
const char *cubit_gpu_last_error(void) { return g_err; }
int cubit_gpu_create(int device, uint64_t n_rows, int64_t row_base, uint32_t seg_bits, cubit_gpu_table **out) {
	(void)device; (void)seg_bits;
	struct cubit_gpu_table *t = calloc(1, sizeof(*t));
	t->n_rows = n_rows; t->n_words = (n_rows + 63) / 64; t->row_base = row_base;
	*out = t;
	return CUBIT_OK;
}
int cubit_gpu_destroy(cubit_gpu_table *t) {
	if (!t) return CUBIT_OK;
	for (int i = 0; i < MAX_COLS; i++) { free(t->cols[i]); free(t->valid[i]); }
	for (int i = 0; i < t->n_indexes; i++) { free(t->bits[i]); free(t->dbits[i]); }
	free(t);
	return CUBIT_OK;
}
int cubit_gpu_upload_column(cubit_gpu_table *t, int32_t col_id, const void *data, uint32_t elem_bytes, uint64_t n) {
	if (elem_bytes != 8 || col_id < 0 || col_id >= MAX_COLS || n != t->n_rows) { snprintf(g_err, sizeof g_err, "mock: bad column"); return CUBIT_EINVAL; }
	free(t->cols[col_id]);
	t->cols[col_id] = malloc(n * 8 + 8);
	memcpy(t->cols[col_id], data, n * 8);
	return CUBIT_OK;
}
int cubit_gpu_download_column(cubit_gpu_table *t, int32_t col_id, void *data, uint32_t elem_bytes, uint64_t n) {
	if (elem_bytes != 8 || col_id < 0 || col_id >= MAX_COLS || !t->cols[col_id] || n != t->n_rows) { snprintf(g_err, sizeof g_err, "mock: bad column"); return CUBIT_EINVAL; }
	memcpy(data, t->cols[col_id], n * 8);
	return CUBIT_OK;
}
int cubit_gpu_drop_column(cubit_gpu_table *t, int32_t col_id) {
	if (col_id < 0 || col_id >= MAX_COLS || !t->cols[col_id]) { snprintf(g_err, sizeof g_err, "mock: bad column"); return CUBIT_EINVAL; }
	free(t->cols[col_id]); t->cols[col_id] = NULL; free(t->valid[col_id]); t->valid[col_id] = NULL;
	for (int ix = 0; ix < t->n_indexes; ix++) if (t->src_col[ix] == col_id) t->src_col[ix] = -1;
	return CUBIT_OK;
}
int cubit_gpu_upload_column_validity(cubit_gpu_table *t, int32_t col_id, const uint64_t *words, uint64_t n_words) {
	if (col_id < 0 || col_id >= MAX_COLS || !t->cols[col_id] || (words && n_words != t->n_words)) { snprintf(g_err, sizeof g_err, "mock: bad validity"); return CUBIT_EINVAL; }
	free(t->valid[col_id]); t->valid[col_id] = NULL;
	if (words) { t->valid[col_id] = malloc(n_words * 8); memcpy(t->valid[col_id], words, n_words * 8); }
	return CUBIT_OK;
}
int oracle_bitpacking_decode(const uint8_t *seg, uint64_t seg_bytes, uint32_t elem_bytes, uint64_t count, void *out,
                             uint64_t *mode_hist);
int oracle_rle_decode(const uint8_t *seg, uint64_t seg_bytes, uint32_t elem_bytes, uint64_t count, void *out,
                      uint64_t *n_runs_out);
int cubit_gpu_upload_column_segments(cubit_gpu_table *t, int32_t col_id, uint32_t elem_bytes,
                                     const cubit_column_segment *segs, uint32_t n_segs, cubit_decode_info *info) {
	if (elem_bytes != 8 || col_id < 0 || col_id >= MAX_COLS) { snprintf(g_err, sizeof g_err, "mock: bad column"); return CUBIT_EINVAL; }
	int64_t *col = malloc(t->n_rows * 8);
	uint64_t next = 0, hist[6] = {0};
	for (uint32_t i = 0; i < n_segs; i++) {
		const cubit_column_segment *s = &segs[i];
		if (s->row_start != next || next + s->count > t->n_rows) { free(col); snprintf(g_err, sizeof g_err, "mock: segments do not tile"); return CUBIT_EINVAL; }
		if (s->kind == CUBIT_SEG_UNCOMPRESSED) memcpy(col + next, s->data, s->count * 8);
		else if (s->kind == CUBIT_SEG_BITPACKING) {
			if (oracle_bitpacking_decode(s->data, s->bytes, 8, s->count, col + next, hist) != 0) { free(col); snprintf(g_err, sizeof g_err, "mock: malformed segment"); return CUBIT_EINVAL; }
		} else if (s->kind == CUBIT_SEG_RLE) {
			if (oracle_rle_decode(s->data, s->bytes, 8, s->count, col + next, NULL) != 0) { free(col); snprintf(g_err, sizeof g_err, "mock: malformed RLE segment"); return CUBIT_EINVAL; }
		} else { free(col); snprintf(g_err, sizeof g_err, "mock: kind"); return CUBIT_EINVAL; }
		next += s->count;
	}
	if (next != t->n_rows) { free(col); snprintf(g_err, sizeof g_err, "mock: rows"); return CUBIT_EINVAL; }
	free(t->cols[col_id]);
	t->cols[col_id] = col;
	if (info) { memset(info, 0, sizeof(*info)); info->n_groups = hist[2] + hist[3] + hist[4] + hist[5]; }
	return CUBIT_OK;
}
int cubit_gpu_index_create(cubit_gpu_table *t, uint32_t cardinality, int32_t *index_id) {
	if (t->n_indexes == MAX_INDEXES) { snprintf(g_err, sizeof g_err, "mock: too many indexes"); return CUBIT_EINVAL; }
	t->card[t->n_indexes] = cardinality; t->bits[t->n_indexes] = calloc((size_t)cardinality * t->n_words, 8);
	t->src_col[t->n_indexes] = -1;
	*index_id = t->n_indexes++;
	return CUBIT_OK;
}
int cubit_gpu_index_build(cubit_gpu_table *t, int32_t index_id, int32_t col_id, int64_t base_value) {
	if (index_id < 0 || index_id >= t->n_indexes) { snprintf(g_err, sizeof g_err, "mock: bad index"); return CUBIT_EINVAL; }
	memset(t->bits[index_id], 0, (size_t)t->card[index_id] * t->n_words * 8);
	oracle_build_index(t->cols[col_id], 8, t->n_rows, base_value, t->card[index_id], t->bits[index_id], t->n_words);
	t->src_col[index_id] = col_id; t->src_base[index_id] = base_value;
	return CUBIT_OK;
}
int cubit_gpu_add_delta_pairs(cubit_gpu_table *t, int32_t index_id, const uint32_t *value_ids, const int64_t *rows, uint64_t n) {
	if (index_id < 0 || index_id >= t->n_indexes) { snprintf(g_err, sizeof g_err, "mock: bad index"); return CUBIT_EINVAL; }
	if (!t->dbits[index_id]) t->dbits[index_id] = calloc((size_t)t->card[index_id] * t->n_words, 8);
	for (uint64_t i = 0; i < n; i++) {
		if (value_ids[i] >= t->card[index_id] || rows[i] < 0 || (uint64_t)rows[i] >= t->n_rows) { snprintf(g_err, sizeof g_err, "mock: bad delta pair"); return CUBIT_EINVAL; }
		t->dbits[index_id][(size_t)value_ids[i] * t->n_words + (uint64_t)rows[i] / 64] ^= 1ull << ((uint64_t)rows[i] % 64);
	}
	return CUBIT_OK;
}
/* INSERT: columns grow, bitvectors are re-strided (the mock keeps them at exactly n_words), indexes built from a
 * resident column index the new rows */
int cubit_gpu_append_rows(cubit_gpu_table *t, uint64_t n_new, const cubit_append_column *cols, uint32_t n_cols) {
	const uint64_t new_n = t->n_rows + n_new, new_w = (new_n + 63) / 64;
	for (uint32_t i = 0; i < n_cols; i++) {
		const int32_t c = cols[i].col_id;
		if (c < 0 || c >= MAX_COLS || !t->cols[c] || cols[i].elem_bytes != 8) { snprintf(g_err, sizeof g_err, "mock: bad append column"); return CUBIT_EINVAL; }
		if (t->valid[c]) { snprintf(g_err, sizeof g_err, "mock: append to a column with NULLs"); return CUBIT_ESTATE; }
		t->cols[c] = realloc(t->cols[c], new_n * 8 + 8);
		memcpy(t->cols[c] + t->n_rows, cols[i].data, n_new * 8);
	}
	for (int ix = 0; ix < t->n_indexes; ix++) {
		for (int which = 0; which < 2; which++) {
			uint64_t **pp = which ? &t->dbits[ix] : &t->bits[ix];
			if (!*pp) continue;
			uint64_t *nb = calloc((size_t)t->card[ix] * new_w, 8);
			for (uint32_t v = 0; v < t->card[ix]; v++) memcpy(nb + (size_t)v * new_w, *pp + (size_t)v * t->n_words, t->n_words * 8);
			free(*pp); *pp = nb;
		}
	}
	const uint64_t old_n = t->n_rows;
	t->n_rows = new_n; t->n_words = new_w;
	for (int ix = 0; ix < t->n_indexes; ix++) { /* only the NEW rows are indexed, like the library (row_begin = old_n) */
		if (t->src_col[ix] >= 0 && t->cols[t->src_col[ix]]) {
			for (uint64_t r = old_n; r < new_n; r++) {
				const int64_t v = t->cols[t->src_col[ix]][r] - t->src_base[ix];
				if (v >= 0 && v < (int64_t)t->card[ix]) t->bits[ix][(size_t)v * new_w + r / 64] |= 1ull << (r % 64);
			}
		}
	}
	return CUBIT_OK;
}
int cubit_gpu_shard_count(const cubit_gpu_table *t, uint32_t *n) { (void)t; *n = 1; return CUBIT_OK; }
int cubit_gpu_device_count(int *count) { *count = 1; return CUBIT_OK; }
int cubit_gpu_pack_column(cubit_gpu_table *t, int32_t col_id, int keep_raw, uint64_t *packed_bytes) {
	(void)t; (void)col_id; (void)keep_raw; /* storage form only: the mock answers from its decoded arrays */
	if (packed_bytes) *packed_bytes = 0;
	return CUBIT_OK;
}
int cubit_gpu_create_sharded(const int *devices, uint32_t n_devices, uint64_t n_rows, int64_t row_base, uint32_t seg_bits,
                             cubit_gpu_table **out) {
	(void)devices; (void)n_devices; /* the mock keeps one shard whatever the device list says */
	return cubit_gpu_create(0, n_rows, row_base, seg_bits, out);
}
/* index image: {n_rows, card, has_delta} + bits (+ delta bits) — the mock's own format */
int cubit_gpu_index_serialize(cubit_gpu_table *t, int32_t index_id, void **image, uint64_t *bytes) {
	if (index_id < 0 || index_id >= t->n_indexes) { snprintf(g_err, sizeof g_err, "mock: bad index"); return CUBIT_EINVAL; }
	const uint64_t nb = (uint64_t)t->card[index_id] * t->n_words * 8, hd = t->dbits[index_id] ? 1 : 0;
	uint64_t *img = malloc(40 + nb * (1 + hd));
	img[0] = t->n_rows; img[1] = t->card[index_id]; img[2] = hd;
	img[3] = (uint64_t)(int64_t)t->src_col[index_id]; img[4] = (uint64_t)t->src_base[index_id];
	memcpy(img + 5, t->bits[index_id], nb);
	if (hd) memcpy((char *)(img + 5) + nb, t->dbits[index_id], nb);
	*image = img; *bytes = 40 + nb * (1 + hd);
	return CUBIT_OK;
}
int cubit_gpu_index_deserialize(cubit_gpu_table *t, const void *image, uint64_t bytes, int32_t *index_id) {
	const uint64_t *img = image;
	if (bytes < 40 || img[0] != t->n_rows) { snprintf(g_err, sizeof g_err, "mock: image describes another table"); return CUBIT_EINVAL; }
	const uint64_t nb = img[1] * t->n_words * 8;
	if (bytes != 40 + nb * (1 + img[2])) { snprintf(g_err, sizeof g_err, "mock: bad image"); return CUBIT_EINVAL; }
	int rc = cubit_gpu_index_create(t, (uint32_t)img[1], index_id);
	if (rc) return rc;
	t->src_col[*index_id] = (int32_t)(int64_t)img[3]; t->src_base[*index_id] = (int64_t)img[4];
	memcpy(t->bits[*index_id], img + 5, nb);
	if (img[2]) { t->dbits[*index_id] = malloc(nb); memcpy(t->dbits[*index_id], (const char *)(img + 5) + nb, nb); }
	return CUBIT_OK;
}
void cubit_gpu_free_image(void *image) { free(image); }
int cubit_gpu_free_result(cubit_gpu_result *r);
int cubit_gpu_query(cubit_gpu_table *t, const cubit_query *q, cubit_gpu_result **out) {
	const uint64_t *streams[CUBIT_MAX_STREAMS], *deltas[CUBIT_MAX_STREAMS]; int32_t group_of[CUBIT_MAX_STREAMS]; int k = 0;
	for (uint32_t g = 0; g < q->n_groups; g++)
		for (uint32_t i = 0; i < q->groups[g].n_refs; i++) {
			const int32_t ix = q->groups[g].refs[i].index_id;
			if (ix < 0 || ix >= t->n_indexes || q->groups[g].refs[i].value_id >= t->card[ix] || k >= CUBIT_MAX_STREAMS) { snprintf(g_err, sizeof g_err, "mock: bad bitvector ref"); return CUBIT_EINVAL; }
			deltas[k] = t->dbits[ix] ? t->dbits[ix] + (size_t)q->groups[g].refs[i].value_id * t->n_words : NULL;
			streams[k] = t->bits[ix] + (size_t)q->groups[g].refs[i].value_id * t->n_words; group_of[k++] = (int32_t)g;
		}
	uint64_t *qb = malloc(t->n_words * 8);
	oracle_merge(streams, deltas, group_of, k, t->n_words, qb);
	struct cubit_gpu_result *r = calloc(1, sizeof(*r));
	r->t = t; r->count = oracle_popcount(qb, t->n_words);
	r->ids = malloc((r->count + 1) * 8);
	oracle_decode(qb, t->n_words, t->row_base, r->ids);
	free(qb);
	if (q->flags & CUBIT_Q_VALUES) {
		r->n_cols = q->n_cols;
		for (uint32_t c = 0; c < q->n_cols; c++) {
			r->vals[c] = malloc((r->count + 1) * 8);
			oracle_probe(r->ids, r->count, t->row_base, t->cols[q->cols[c]], 8, r->vals[c]);
			if (t->valid[q->cols[c]]) {
				r->vmask[c] = calloc((r->count + 63) / 64 + 1, 8);
				oracle_probe_validity(r->ids, r->count, t->row_base, t->valid[q->cols[c]], r->vmask[c]);
			}
		}
	}
	if (q->agg_kind == CUBIT_AGG_SUM || q->agg_kind == CUBIT_AGG_SUM_PROD) {
		int ovf = 0;
		const int prod = q->agg_kind == CUBIT_AGG_SUM_PROD;
		r->agg_rows = oracle_sum_nulls(r->ids, r->count, t->row_base, t->cols[q->agg_col_a], t->valid[q->agg_col_a],
		                               prod ? t->cols[q->agg_col_b] : NULL, prod ? t->valid[q->agg_col_b] : NULL,
		                               &r->sum_lo, &r->sum_hi, &ovf);
		if (ovf) { snprintf(g_err, sizeof g_err, "Overflow in multiplication of INT64 in SUM(a*b)"); cubit_gpu_free_result(r); return CUBIT_EINVAL; }
	}
	*out = r;
	return CUBIT_OK;
}
int cubit_gpu_result_get(cubit_gpu_result *r, cubit_result_info *info) {
	memset(info, 0, sizeof(*info));
	info->count = r->count; info->sum_lo = r->sum_lo; info->sum_hi = r->sum_hi; info->agg_rows = r->agg_rows;
	for (uint32_t c = 0; c < r->n_cols; c++) info->d_validity[c] = (const uint32_t *)r->vmask[c];
	return CUBIT_OK;
}
int cubit_gpu_fetch(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                    void *const *host_cols) {
	if (offset + n > r->count || n_cols > r->n_cols) { snprintf(g_err, sizeof g_err, "mock: bad fetch"); return CUBIT_EINVAL; }
	if (host_rowids) memcpy(host_rowids, r->ids + offset, n * 8);
	for (uint32_t c = 0; c < n_cols; c++) memcpy(host_cols[c], r->vals[c] + offset, n * 8);
	return CUBIT_OK;
}
int cubit_gpu_fetch_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int64_t *host_rowids, uint32_t n_cols,
                          void *const *host_cols, cubit_gpu_fetch_ticket **ticket) {
	*ticket = calloc(1, sizeof(**ticket));
	return cubit_gpu_fetch(r, offset, n, host_rowids, n_cols, host_cols);
}
int cubit_gpu_fetch_wait(cubit_gpu_fetch_ticket *ticket) { free(ticket); return CUBIT_OK; }
/* narrow wire (include/cubit_gpu_wire.h) written on the CPU: per frame the narrowest width, frames back to back in
 * chunk-major order, like the device writes them (the mock never uses bitmap frames; delta frames are always valid) */
int cubit_gpu_fetch_wire_async(cubit_gpu_result *r, uint64_t offset, uint64_t n, int with_rowids, uint32_t n_cols,
                               void *host_wire, uint64_t host_wire_bytes, cubit_gpu_fetch_ticket **ticket) {
	const uint32_t streams = (with_rowids ? 1u : 0u) + n_cols;
	if (offset + n > r->count || n_cols > r->n_cols || !streams || host_wire_bytes < cubit_wire_bytes(n, streams)) {
		snprintf(g_err, sizeof g_err, "mock: bad wire fetch"); return CUBIT_EINVAL;
	}
	cubit_wire_header *h = (cubit_wire_header *)host_wire;
	memset(h, 0, sizeof *h);
	h->magic = CUBIT_WIRE_MAGIC; h->n_streams = streams; h->n_rows = n;
	h->n_chunks = (n + CUBIT_WIRE_CHUNK - 1) / CUBIT_WIRE_CHUNK;
	h->data_offset = cubit_wire_data_offset(h->n_chunks, streams);
	cubit_wire_dir *dir = (cubit_wire_dir *)((char *)host_wire + sizeof(cubit_wire_header));
	uint64_t at = 0;
	for (uint64_t c = 0; c < h->n_chunks; c++) {
		for (uint32_t s = 0; s < streams; s++) {
			h->elem[s] = 8;
			const int64_t *v = ((with_rowids && s == 0) ? r->ids : r->vals[s - (with_rowids ? 1 : 0)]) + offset;
			const uint64_t r0 = c * CUBIT_WIRE_CHUNK;
			const uint32_t cn = (uint32_t)(n - r0 < CUBIT_WIRE_CHUNK ? n - r0 : CUBIT_WIRE_CHUNK);
			int64_t lo = v[r0], hi = v[r0];
			for (uint32_t i = 1; i < cn; i++) { if (v[r0 + i] < lo) lo = v[r0 + i]; if (v[r0 + i] > hi) hi = v[r0 + i]; }
			const uint64_t range = (uint64_t)hi - (uint64_t)lo;
			const uint32_t w = range == 0 ? 0 : range < 256 ? 1 : range < 65536 ? 2 : range < (1ull << 32) ? 4 : 8;
			cubit_wire_dir *d = dir + c * streams + s;
			d->base = lo; d->offset = at; d->width = w; d->n = cn;
			unsigned char *dst = (unsigned char *)host_wire + h->data_offset + at;
			for (uint32_t i = 0; i < cn; i++) {
				const uint64_t delta = (uint64_t)v[r0 + i] - (uint64_t)lo;
				memcpy(dst + (size_t)i * w, &delta, w); /* little endian */
			}
			at += ((uint64_t)cn * w + 15) & ~15ull;
		}
	}
	*ticket = calloc(1, sizeof(**ticket));
	return CUBIT_OK;
}
uint64_t cubit_gpu_wire_bytes(uint64_t n_rows, uint32_t n_streams) { return cubit_wire_bytes(n_rows, n_streams); }
uint64_t cubit_gpu_wire_payload_bytes(const void *w) { return cubit_wire_payload_bytes(w); }
int cubit_gpu_wire_unpack(const void *w, uint32_t stream, uint64_t chunk, void *out, uint32_t out_elem) {
	return cubit_wire_unpack_chunk(w, stream, chunk, out, out_elem);
}
int cubit_gpu_fetch_validity(cubit_gpu_result *r, uint32_t col, uint64_t offset, uint64_t n, uint64_t *host_words,
                             int *all_valid) {
	if (offset + n > r->count || col >= r->n_cols) { snprintf(g_err, sizeof g_err, "mock: bad fetch"); return CUBIT_EINVAL; }
	int all = 1;
	for (uint64_t w = 0; w < (n + 63) / 64; w++) host_words[w] = 0;
	for (uint64_t j = 0; j < n; j++) {
		uint64_t p = offset + j;
		int bit = r->vmask[col] ? (int)((r->vmask[col][p / 64] >> (p % 64)) & 1) : 1;
		if (bit) host_words[j / 64] |= 1ull << (j % 64); else all = 0;
	}
	if (all_valid) *all_valid = all;
	return CUBIT_OK;
}
int cubit_gpu_alloc_host(uint64_t bytes, void **ptr) { *ptr = malloc(bytes ? bytes : 1); return *ptr ? CUBIT_OK : CUBIT_ENOMEM; }
int cubit_gpu_free_host(void *ptr) { free(ptr); return CUBIT_OK; }
int cubit_gpu_free_result(cubit_gpu_result *r) {
	if (!r) return CUBIT_OK;
	free(r->ids);
	for (int c = 0; c < CUBIT_MAX_PROBE_COLS; c++) { free(r->vals[c]); free(r->vmask[c]); }
	free(r);
	return CUBIT_OK;
}
