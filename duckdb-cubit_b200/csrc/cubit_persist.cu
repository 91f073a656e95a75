#include "table.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

using namespace cubit;

// -------------------------------------------------------------- persistence
// Image layout (little endian, every section 8-byte aligned):
//   ImageHeader | per value: ImageEntry, payload words, delta rows | u64 FNV-1a checksum of everything before
namespace {
struct ImageHeader {
	char magic[8]; // "CUBITIX2"
	uint64_t n_rows;
	uint32_t card;
	int32_t src_col; // column the index was built from, -1 = uploaded bitvectors
	int64_t src_base;
	uint32_t flags;  // kImageCompressed: the index keeps roaring-style containers in HBM (recreated that way)
	uint32_t pad;
};
constexpr uint32_t kImageCompressed = 1u;
struct ImageEntry {
	uint32_t encoding; // 0 = verbatim 64-bit words, 1 = WAH 32-bit words
	uint32_t active_val;
	uint32_t active_nbits;
	uint32_t pad;
	uint64_t n_words;      // 64-bit words (verbatim) / 32-bit words (WAH)
	uint64_t n_delta_rows; // pending flipped rows that follow the payload
};
const char kImageMagic[8] = {'C', 'U', 'B', 'I', 'T', 'I', 'X', '2'};

uint64_t fnv1a(const uint8_t *p, uint64_t n) {
	uint64_t h = 1469598103934665603ull;
	for (uint64_t i = 0; i < n; i++) {
		h = (h ^ p[i]) * 1099511628211ull;
	}
	return h;
}

// WAH-compress a verbatim bitvector (row r = bit r%64 of word r/64) — host side of the persistence path.
// 31-bit groups are cut from a 64-bit window; a literal keeps the group's FIRST row in its most significant
// bit, hence the bit reversal.  Stops (returns false) as soon as the output would not be smaller than `limit`
// 32-bit words, so incompressible bitvectors cost one partial pass.
bool wah_compress(const uint64_t *words, uint64_t n_rows, uint64_t limit, std::vector<uint32_t> &out, uint32_t &active_val,
                  uint32_t &active_nbits) {
	out.clear();
	const uint64_t n_groups = n_rows / 31;
	auto bits_at = [&](uint64_t row, uint32_t n) -> uint32_t { // n ≤ 31 rows starting at `row`, LSB = first row
		const uint64_t w = row >> 6, sh = row & 63;
		uint64_t v = words[w] >> sh;
		if (sh + n > 64) {
			v |= words[w + 1] << (64 - sh);
		}
		return (uint32_t)(v & ((1ull << n) - 1ull));
	};
	auto reverse = [](uint32_t v, uint32_t n) -> uint32_t { // first row → most significant of n bits
		v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
		v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
		v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
		v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
		v = (v >> 16) | (v << 16);
		return v >> (32 - n);
	};
	for (uint64_t g = 0; g < n_groups; g++) {
		const uint32_t raw = bits_at(g * 31, 31);
		if (raw == 0 || raw == 0x7fffffffu) { // (FastBit's append rule: a lone fill group is a literal, fills start at 2)
			const uint32_t fill = 0x80000000u | (raw ? 0x40000000u : 0u);
			if (!out.empty() && out.back() == raw) {
				out.back() = fill | 2u;
				continue;
			}
			if (!out.empty() && (out.back() & 0xc0000000u) == fill && (out.back() & 0x3fffffffu) < 0x3fffffffu) {
				out.back()++;
				continue;
			}
			out.push_back(raw);
		} else {
			out.push_back(reverse(raw, 31));
		}
		if (out.size() >= limit) {
			return false;
		}
	}
	active_nbits = (uint32_t)(n_rows % 31);
	active_val = active_nbits ? reverse(bits_at(n_groups * 31, active_nbits), active_nbits) : 0u;
	return true;
}
} // namespace

extern "C" void cubit_gpu_free_image(void *image) {
	free(image);
}

extern "C" int cubit_gpu_index_serialize(cubit_gpu_table *t, int32_t index_id, void **image, uint64_t *bytes) {
	ABI_BEGIN
	if (!t || !image || !bytes) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	*image = nullptr;
	*bytes = 0;
	if (t->sharded()) {
		return fail(CUBIT_ESTATE, "index images are per shard: serialize every shard of a sharded table");
	}
	TableLock lk(t);
	if (use_device(t)) {
		return CUBIT_ECUDA;
	}
	Index *ix = get_index(t, index_id);
	if (!ix) {
		return fail(CUBIT_EINVAL, "bad index %d", index_id);
	}
	CU_TRY(cudaStreamSynchronize(t->stream));
	std::vector<uint8_t> img;
	auto append = [&](const void *p, size_t n) {
		const uint8_t *b = static_cast<const uint8_t *>(p);
		img.insert(img.end(), b, b + n);
		while (img.size() & 7) {
			img.push_back(0);
		}
	};
	ImageHeader h;
	memset(&h, 0, sizeof(h));
	memcpy(h.magic, kImageMagic, 8);
	h.n_rows = t->n_rows;
	h.card = ix->card;
	h.src_col = ix->src_col;
	h.src_base = ix->src_base;
	h.flags = ix->compressed ? kImageCompressed : 0u;
	append(&h, sizeof(h));
	std::vector<uint64_t> words(t->n_words + 1, 0); // + 1: the group extractor may touch one word past the end
	std::vector<uint32_t> wah;
	std::vector<int64_t> rows;
	uint64_t *d_tmp = nullptr; // compressed indexes: containers → verbatim words on the device, one value at a time
	if (ix->compressed) {
		CU_TRY(cudaMalloc(&d_tmp, t->words_per_bv * 8));
	}
	int rc = CUBIT_OK;
	for (uint32_t v = 0; v < ix->card && rc == CUBIT_OK; v++) {
		const uint64_t *src = ix->compressed ? d_tmp : bv_ptr(t, ix, v);
		if (ix->compressed) {
			rc = expand_value_locked(t, ix, v, d_tmp);
			if (rc) {
				break;
			}
		}
		cudaError_t e = cudaMemcpyAsync(words.data(), src, t->n_words * 8, cudaMemcpyDeviceToHost, t->stream);
		if (e == cudaSuccess) {
			e = cudaStreamSynchronize(t->stream);
		}
		if (e != cudaSuccess) {
			rc = fail(CUBIT_ECUDA, "serialize: %s", cudaGetErrorString(e));
			break;
		}
		words[t->n_words] = 0;
		// pending deltas of this value: the net flipped rows (they stay pending after a reload)
		rc = delta_rows_locked(t, ix, v, rows);
		if (rc) {
			break;
		}
		ImageEntry en;
		memset(&en, 0, sizeof(en));
		en.n_delta_rows = rows.size();
		if (wah_compress(words.data(), t->n_rows, t->n_words * 2, wah, en.active_val, en.active_nbits)) {
			en.encoding = 1;
			en.n_words = wah.size();
			append(&en, sizeof(en));
			append(wah.data(), wah.size() * 4);
		} else {
			en.encoding = 0;
			en.active_val = en.active_nbits = 0;
			en.n_words = t->n_words;
			append(&en, sizeof(en));
			append(words.data(), t->n_words * 8);
		}
		append(rows.data(), rows.size() * 8);
	}
	if (d_tmp) {
		cudaFree(d_tmp);
	}
	if (rc) {
		return rc;
	}
	const uint64_t sum = fnv1a(img.data(), img.size());
	append(&sum, 8);
	void *outp = malloc(img.size());
	if (!outp) {
		return fail(CUBIT_ENOMEM, "host allocation of %zu bytes failed", img.size());
	}
	memcpy(outp, img.data(), img.size());
	*image = outp;
	*bytes = img.size();
	return CUBIT_OK;
	ABI_END
}

extern "C" int cubit_gpu_index_deserialize(cubit_gpu_table *t, const void *image, uint64_t bytes, int32_t *index_id) {
	ABI_BEGIN
	if (!t || !image || !index_id) {
		return fail(CUBIT_EINVAL, "NULL argument");
	}
	if (t->sharded()) {
		return fail(CUBIT_ESTATE, "index images are per shard: deserialize into every shard of a sharded table");
	}
	const uint8_t *p = static_cast<const uint8_t *>(image);
	if (bytes < sizeof(ImageHeader) + 8 || (bytes & 7)) {
		return fail(CUBIT_EINVAL, "index image: bad size %llu", (unsigned long long)bytes);
	}
	ImageHeader h;
	memcpy(&h, p, sizeof(h));
	if (memcmp(h.magic, kImageMagic, 8) != 0) {
		return fail(CUBIT_EINVAL, "index image: bad magic");
	}
	uint64_t sum;
	memcpy(&sum, p + bytes - 8, 8);
	if (sum != fnv1a(p, bytes - 8)) {
		return fail(CUBIT_EINVAL, "index image: checksum mismatch");
	}
	if (h.n_rows != t->n_rows) {
		return fail(CUBIT_EINVAL, "index image describes %llu rows, table has %llu", (unsigned long long)h.n_rows,
		            (unsigned long long)t->n_rows);
	}
	if (h.card == 0) {
		return fail(CUBIT_EINVAL, "index image: cardinality 0");
	}
	// pass 1: structure
	struct Sec {
		ImageEntry e;
		uint64_t payload, rows;
	};
	std::vector<Sec> secs;
	uint64_t at = sizeof(ImageHeader);
	const uint64_t end = bytes - 8;
	for (uint32_t v = 0; v < h.card; v++) {
		if (at + sizeof(ImageEntry) > end) {
			return fail(CUBIT_EINVAL, "index image: truncated at value %u", v);
		}
		Sec s;
		memcpy(&s.e, p + at, sizeof(ImageEntry));
		at += sizeof(ImageEntry);
		if (s.e.encoding > 1 || (s.e.encoding == 0 && s.e.n_words != t->n_words) || s.e.n_words > end ||
		    s.e.n_delta_rows > end) {
			return fail(CUBIT_EINVAL, "index image: bad entry for value %u", v);
		}
		const uint64_t pbytes = ((s.e.encoding ? s.e.n_words * 4 : s.e.n_words * 8) + 7) & ~7ull;
		if (at + pbytes + s.e.n_delta_rows * 8 > end) {
			return fail(CUBIT_EINVAL, "index image: truncated payload of value %u", v);
		}
		s.payload = at;
		s.rows = at + pbytes;
		at = s.rows + s.e.n_delta_rows * 8;
		secs.push_back(s);
	}
	if (at != end) {
		return fail(CUBIT_EINVAL, "index image: %llu trailing bytes", (unsigned long long)(end - at));
	}
	// pass 2: rebuild through the public entry points (each validates its input again)
	int32_t id = -1;
	int rc = (h.flags & kImageCompressed) ? cubit_gpu_index_create_compressed(t, h.card, &id)
	                                      : cubit_gpu_index_create(t, h.card, &id);
	for (uint32_t v = 0; v < h.card && rc == CUBIT_OK; v++) {
		const Sec &s = secs[v];
		if (s.e.encoding == 1) {
			cubit_wah_bitvector bv;
			bv.words = reinterpret_cast<const uint32_t *>(p + s.payload);
			bv.n_words = s.e.n_words;
			bv.active_val = s.e.active_val;
			bv.active_nbits = s.e.active_nbits;
			rc = cubit_gpu_upload_bitvector_wah(t, id, v, &bv);
		} else {
			rc = cubit_gpu_upload_bitvector(t, id, v, reinterpret_cast<const uint64_t *>(p + s.payload), s.e.n_words);
		}
		if (rc == CUBIT_OK && s.e.n_delta_rows) {
			rc = cubit_gpu_set_delta(t, id, v, reinterpret_cast<const int64_t *>(p + s.rows), s.e.n_delta_rows);
		}
	}
	if (rc != CUBIT_OK) {
		// drop the half-built index (it is the last one created: nothing else can hold its id yet)
		const std::string why = last_error_cstr();
		TableLock lk(t);
		if (id >= 0 && (size_t)id + 1 == t->indexes.size()) {
			Index *ix = t->indexes.back();
			cudaStreamSynchronize(t->stream);
			free_delta(ix->delta);
			free_compressed(ix->cs);
			if (ix->d_bits) {
				cudaFree(ix->d_bits);
			}
			delete ix;
			t->indexes.pop_back();
		}
		return fail(rc, "%s", why.c_str());
	}
	{
		TableLock lk(t);
		Index *ix = get_index(t, id);
		if (ix) { // remember where the bitvectors came from, so appends keep extending the index on the GPU
			ix->src_col = h.src_col;
			ix->src_base = h.src_base;
		}
	}
	*index_id = id;
	return CUBIT_OK;
	ABI_END
}
