// ref_fastpfor_shim.cpp — extern "C" doors into the REFERENCE's own bit packer, for oracle/_ref only.
// Compiled together with /root/reference/third_party/fastpforlib/bitpacking.cpp (where it lies; never copied)
// by oracle.build_ref() / `make -C oracle ref`.  TEST INFRASTRUCTURE: validates oracle/bitpacking_oracle.c.
#include "bitpackinghelpers.h"

extern "C" {
__attribute__((visibility("default"))) void ref_fastunpack64(const uint32_t *in, uint64_t *out, uint32_t bit) {
	duckdb_fastpforlib::fastunpack(in, out, bit);
}
__attribute__((visibility("default"))) void ref_fastpack64(const uint64_t *in, uint32_t *out, uint32_t bit) {
	duckdb_fastpforlib::fastpack(in, out, bit);
}
__attribute__((visibility("default"))) void ref_fastunpack32(const uint32_t *in, uint32_t *out, uint32_t bit) {
	duckdb_fastpforlib::fastunpack(in, out, bit);
}
__attribute__((visibility("default"))) void ref_fastpack32(const uint32_t *in, uint32_t *out, uint32_t bit) {
	duckdb_fastpforlib::fastpack(in, out, bit);
}
}
