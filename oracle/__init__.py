"""CPU oracle of the CUBIT bitmap-index scan path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package.  The product path (duckdb-cubit_b200/) never does.

Two independent restatements live here:
  * cubit_oracle.c  (C, via ctypes)  — also the multi-threaded CPU baseline
  * np_*            (numpy)          — a second implementation used to cross-check the C one
Parity status and reference citations: see the header of cubit_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcubit_oracle.so")
_lib = None

_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


_SOURCES = ["cubit_oracle.c", "bitpacking_oracle.c", "wah_oracle.c", "rle_oracle.c"]
_REF_DIR = os.path.join(_HERE, "_ref")
_REF_FASTPFOR = os.path.join(_REF_DIR, "libfastpfor_ref.so")
_REFERENCE_FASTPFOR_SRC = "/root/reference/third_party/fastpforlib"


def build(force=False):
    srcs = [os.path.join(_HERE, s) for s in _SOURCES]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        flags = ["-O3", "-fPIC", "-Wall", "-Wextra", "-fvisibility=hidden"]
        # -march=native is only valid for the machine that compiles; the library may have been
        # built in another container, so rebuild here if it was not produced on this host.
        subprocess.check_call(["gcc"] + flags + ["-march=native", "-shared", "-o", _LIB_PATH] + srcs + ["-lpthread"])
    return _LIB_PATH


def build_ref(force=False):
    """oracle/_ref/libfastpfor_ref.so: the REFERENCE's own 32-value pack/unpack (third_party/fastpforlib,
    the arithmetic under BitpackingPrimitives::PackGroup/UnPackGroup), compiled from the sources where they
    lie under /root/reference plus a 20-line extern "C" shim of ours.  Only possible in the build container;
    the built .so travels to the GPU box.  Returns the path, or None when neither sources nor a built .so exist."""
    shim = os.path.join(_HERE, "ref_fastpfor_shim.cpp")
    if os.path.isdir(_REFERENCE_FASTPFOR_SRC):
        ref_src = os.path.join(_REFERENCE_FASTPFOR_SRC, "bitpacking.cpp")
        if force or not os.path.exists(_REF_FASTPFOR) or os.path.getmtime(_REF_FASTPFOR) < os.path.getmtime(shim):
            os.makedirs(_REF_DIR, exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++11", "-I", _REFERENCE_FASTPFOR_SRC,
                                   "-o", _REF_FASTPFOR, shim, ref_src])
    return _REF_FASTPFOR if os.path.exists(_REF_FASTPFOR) else None


_ref_fastpfor = None


def ref_fastpfor():
    """ctypes handle of oracle/_ref/libfastpfor_ref.so (None if it cannot be built or found)"""
    global _ref_fastpfor
    if _ref_fastpfor is None:
        path = build_ref()
        if path is None:
            return None
        L = C.CDLL(path)
        u32p = C.POINTER(C.c_uint32)
        L.ref_fastunpack64.argtypes = [u32p, _u64p, C.c_uint32]
        L.ref_fastpack64.argtypes = [_u64p, u32p, C.c_uint32]
        L.ref_fastunpack32.argtypes = [u32p, u32p, C.c_uint32]
        L.ref_fastpack32.argtypes = [u32p, u32p, C.c_uint32]
        for f in (L.ref_fastunpack64, L.ref_fastpack64, L.ref_fastunpack32, L.ref_fastpack32):
            f.restype = None
        _ref_fastpfor = L
    return _ref_fastpfor


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.oracle_merge.argtypes = [C.POINTER(_u64p), C.POINTER(_u64p), _i32p, C.c_int, C.c_uint64, _u64p]
        L.oracle_merge.restype = None
        L.oracle_merge_segmented.argtypes = [C.POINTER(_u64p), C.POINTER(_u64p), _i32p, C.c_int, C.c_uint64,
                                             C.c_uint64, _u64p]
        L.oracle_merge_segmented.restype = None
        L.oracle_decode.argtypes = [_u64p, C.c_uint64, C.c_int64, _i64p]
        L.oracle_decode.restype = C.c_uint64
        L.oracle_popcount.argtypes = [_u64p, C.c_uint64]
        L.oracle_popcount.restype = C.c_uint64
        L.oracle_probe.argtypes = [_i64p, C.c_uint64, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_probe.restype = None
        L.oracle_sum_i64.argtypes = [_i64p, C.c_uint64, _u64p, _i64p]
        L.oracle_sum_i64.restype = None
        L.oracle_sum_prod_i64.argtypes = [_i64p, _i64p, C.c_uint64, _u64p, _i64p]
        L.oracle_sum_prod_i64.restype = C.c_int
        L.oracle_sum_f64.argtypes = [C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_sum_f64.restype = None
        L.oracle_probe_validity.argtypes = [_i64p, C.c_uint64, C.c_int64, _u64p, _u64p]
        L.oracle_probe_validity.restype = C.c_uint64
        L.oracle_sum_nulls.argtypes = [_i64p, C.c_uint64, C.c_int64, _i64p, _u64p, _i64p, _u64p, _u64p, _i64p,
                                       C.POINTER(C.c_int)]
        L.oracle_sum_nulls.restype = C.c_uint64
        L.oracle_build_index.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64, C.c_uint32, _u64p,
                                         C.c_uint64]
        L.oracle_build_index.restype = None
        L.oracle_delta_from_rows.argtypes = [_i64p, C.c_uint64, _u64p]
        L.oracle_delta_from_rows.restype = None
        L.oracle_synth_column.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_synth_column.restype = None
        L.oracle_synth_bitvectors.argtypes = [_u64p, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint32,
                                              C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.oracle_synth_bitvectors.restype = None
        L.oracle_scan_mt.argtypes = [C.POINTER(_u64p), C.POINTER(_u64p), _i32p, C.c_int, C.c_uint64, C.c_int64,
                                     _u64p, _i64p, _i64p, _i64p, _u64p, _i64p, C.c_int]
        L.oracle_scan_mt.restype = C.c_uint64
        u8p = C.POINTER(C.c_uint8)
        L.oracle_bitpacking_decode.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, _u64p]
        L.oracle_bitpacking_decode.restype = C.c_int
        L.oracle_bitpacking_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, u8p, C.c_uint64]
        L.oracle_bitpacking_encode.restype = C.c_int64
        u32p = C.POINTER(C.c_uint32)
        L.oracle_wah_encode.argtypes = [_u64p, C.c_uint64, u32p, C.c_uint64, u32p, u32p]
        L.oracle_wah_encode.restype = C.c_int64
        L.oracle_wah_bits.argtypes = [u32p, C.c_uint64, C.c_uint32]
        L.oracle_wah_bits.restype = C.c_int64
        L.oracle_wah_decode.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, C.c_uint64]
        L.oracle_wah_decode.restype = C.c_int64
        L.oracle_rle_decode.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, _u64p]
        L.oracle_rle_decode.restype = C.c_int
        L.oracle_rle_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, u8p, C.c_uint64]
        L.oracle_rle_encode.restype = C.c_int64
        L.oracle_bp_unpack_group64.argtypes = [u8p, _u64p, C.c_uint32]
        L.oracle_bp_unpack_group64.restype = None
        L.oracle_bp_unpack_group32.argtypes = [u8p, C.POINTER(C.c_uint32), C.c_uint32]
        L.oracle_bp_unpack_group32.restype = None
    return _lib


def _p(a, typ):
    return a.ctypes.data_as(typ)


def _stream_ptrs(arrs):
    """arrs: list of 1-D uint64 arrays (or None) -> (ctypes array of pointers, keepalive)"""
    ptrs = (_u64p * len(arrs))()
    for i, a in enumerate(arrs):
        ptrs[i] = _p(a, _u64p) if a is not None else _u64p()
    return ptrs


def int128(lo, hi):
    """(uint64 lo, int64 hi) two's complement limbs -> python int"""
    return (int(hi) << 64) + int(lo)


# ------------------------------------------------------------------ C oracle
def merge(groups, deltas=None, seg_words=None):
    """groups: list of lists of uint64 arrays (OR inside a group, AND across).
    deltas: same shape, dense uint64 delta bitvectors or None entries."""
    streams, dl, gof = [], [], []
    for g, grp in enumerate(groups):
        for i, b in enumerate(grp):
            b = np.ascontiguousarray(b, dtype=np.uint64)
            streams.append(b)
            d = None
            if deltas is not None and deltas[g][i] is not None:
                d = np.ascontiguousarray(deltas[g][i], dtype=np.uint64)
            dl.append(d)
            gof.append(g)
    n_words = len(streams[0])
    q = np.empty(n_words, dtype=np.uint64)
    gof = np.asarray(gof, dtype=np.int32)
    sp = _stream_ptrs(streams)
    dp = _stream_ptrs(dl)
    dpp = dp if any(d is not None for d in dl) else C.POINTER(_u64p)()
    if seg_words:
        lib().oracle_merge_segmented(sp, dpp, _p(gof, _i32p), len(streams), n_words, seg_words, _p(q, _u64p))
    else:
        lib().oracle_merge(sp, dpp, _p(gof, _i32p), len(streams), n_words, _p(q, _u64p))
    return q


def decode(q, row_base=0):
    q = np.ascontiguousarray(q, dtype=np.uint64)
    n = lib().oracle_popcount(_p(q, _u64p), len(q))
    out = np.empty(n, dtype=np.int64)
    got = lib().oracle_decode(_p(q, _u64p), len(q), row_base, _p(out, _i64p))
    assert got == n
    return out


def popcount(q):
    q = np.ascontiguousarray(q, dtype=np.uint64)
    return int(lib().oracle_popcount(_p(q, _u64p), len(q)))


def probe(ids, col, row_base=0):
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    col = np.ascontiguousarray(col)
    out = np.empty(len(ids), dtype=col.dtype)
    lib().oracle_probe(_p(ids, _i64p), len(ids), row_base, col.ctypes.data, col.dtype.itemsize, out.ctypes.data)
    return out


def probe_validity(ids, valid, row_base=0):
    """validity bits of a probed column at the row ids → (uint64 mask words over result positions, n valid)"""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.zeros((len(ids) + 63) // 64, dtype=np.uint64)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint64)
    nv = lib().oracle_probe_validity(_p(ids, _i64p), len(ids), row_base, None if v is None else _p(v, _u64p),
                                     _p(out, _u64p) if len(out) else None)
    return out, int(nv)


def sum_nulls(ids, a, valid_a=None, b=None, valid_b=None, row_base=0):
    """SUM(a) / SUM(a*b) over the rows `ids`, skipping NULL inputs → (int128 sum, rows aggregated, overflow)"""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=np.int64)
    va = None if valid_a is None else np.ascontiguousarray(valid_a, dtype=np.uint64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.int64)
    vb = None if valid_b is None else np.ascontiguousarray(valid_b, dtype=np.uint64)
    lo, hi, ovf = C.c_uint64(0), C.c_int64(0), C.c_int(0)
    rows = lib().oracle_sum_nulls(_p(ids, _i64p), len(ids), row_base, _p(a, _i64p),
                                  None if va is None else _p(va, _u64p), None if bb is None else _p(bb, _i64p),
                                  None if vb is None else _p(vb, _u64p), C.byref(lo), C.byref(hi), C.byref(ovf))
    return int128(lo.value, hi.value), int(rows), bool(ovf.value)


def sum_i64(vals):
    vals = np.ascontiguousarray(vals, dtype=np.int64)
    lo, hi = C.c_uint64(0), C.c_int64(0)
    lib().oracle_sum_i64(_p(vals, _i64p), len(vals), C.byref(lo), C.byref(hi))
    return int128(lo.value, hi.value)


def sum_prod_i64(a, b):
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    lo, hi = C.c_uint64(0), C.c_int64(0)
    ovf = lib().oracle_sum_prod_i64(_p(a, _i64p), _p(b, _i64p), len(a), C.byref(lo), C.byref(hi))
    return int128(lo.value, hi.value), bool(ovf)


def sum_f64(vals):
    """→ (scan-order double sum, compensated sum)"""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    a, b = C.c_double(0), C.c_double(0)
    lib().oracle_sum_f64(vals.ctypes.data_as(C.POINTER(C.c_double)), len(vals), C.byref(a), C.byref(b))
    return a.value, b.value


def build_index(col, base_value, card):
    col = np.ascontiguousarray(col)
    assert col.dtype.itemsize in (4, 8)
    n = len(col)
    n_words = (n + 63) // 64
    bv = np.zeros((card, n_words), dtype=np.uint64)
    lib().oracle_build_index(col.ctypes.data, col.dtype.itemsize, n, base_value, card, _p(bv, _u64p), n_words)
    return bv


def delta_from_rows(rows, n_rows):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    d = np.zeros((n_rows + 63) // 64, dtype=np.uint64)
    lib().oracle_delta_from_rows(_p(rows, _i64p), len(rows), _p(d, _u64p))
    return d


def synth_column(kind, n_rows, row_base=0, seed=0, threshold=0, card=100, hot_lo=10, hot_n=10):
    col = np.empty(n_rows, dtype=np.int64 if kind in (0, 3) else np.int32)
    lib().oracle_synth_column(col.ctypes.data, kind, n_rows, row_base, seed, threshold, card, hot_lo, hot_n)
    return col


def synth_bitvectors(n_rows, row_base, seed, threshold, card, hot_lo, hot_n, v0, nv, n_threads=1):
    n_words = (n_rows + 63) // 64
    out = np.empty((nv, n_words), dtype=np.uint64)
    lib().oracle_synth_bitvectors(_p(out, _u64p), n_rows, row_base, seed, threshold, card, hot_lo, hot_n, v0, nv,
                                  n_threads)
    return out


def scan_mt(groups, payload=None, row_base=0, n_threads=1, want_ids=True, deltas=None, bufs=None):
    """The whole path, multi-threaded: merge -> decode -> probe payload -> SUM.
    Returns (count, ids, vals, sum).  `bufs` = (q, ids, vals) preallocated arrays to reuse."""
    streams, dl, gof = [], [], []
    for g, grp in enumerate(groups):
        for i, b in enumerate(grp):
            streams.append(b)
            dl.append(deltas[g][i] if deltas is not None else None)
            gof.append(g)
    n_words = len(streams[0])
    gof = np.asarray(gof, dtype=np.int32)
    if bufs is None:
        cap = n_words * 64
        q = np.empty(n_words, dtype=np.uint64)
        ids = np.empty(cap, dtype=np.int64) if want_ids else None
        vals = np.empty(cap, dtype=np.int64) if payload is not None and want_ids else None
    else:
        q, ids, vals = bufs
    sp = _stream_ptrs(streams)
    dp = _stream_ptrs(dl)
    dpp = dp if any(d is not None for d in dl) else C.POINTER(_u64p)()
    lo, hi = C.c_uint64(0), C.c_int64(0)
    n = lib().oracle_scan_mt(sp, dpp, _p(gof, _i32p), len(streams), n_words, row_base, _p(q, _u64p),
                             _p(ids, _i64p) if ids is not None else _i64p(),
                             _p(payload, _i64p) if payload is not None else _i64p(),
                             _p(vals, _i64p) if vals is not None else _i64p(), C.byref(lo), C.byref(hi), n_threads)
    return int(n), ids, vals, int128(lo.value, hi.value)


# --------------------------------------------------------------- numpy oracle
def np_merge(groups, deltas=None):
    acc = None
    for g, grp in enumerate(groups):
        o = np.zeros_like(np.asarray(grp[0], dtype=np.uint64))
        for i, b in enumerate(grp):
            v = np.asarray(b, dtype=np.uint64)
            if deltas is not None and deltas[g][i] is not None:
                v = v ^ np.asarray(deltas[g][i], dtype=np.uint64)
            o = o | v
        acc = o if acc is None else (acc & o)
    return acc


def np_decode(q, row_base=0):
    bits = np.unpackbits(np.ascontiguousarray(q, dtype="<u8").view(np.uint8), bitorder="little")
    return np.flatnonzero(bits).astype(np.int64) + row_base


def np_build_index(col, base_value, card):
    col = np.asarray(col).astype(np.int64)
    n = len(col)
    n_words = (n + 63) // 64
    out = np.zeros((card, n_words), dtype=np.uint64)
    for v in range(card):
        m = np.zeros(n_words * 64, dtype=np.uint8)
        m[:n] = col == (base_value + v)
        out[v] = np.packbits(m, bitorder="little").view("<u8")
    return out


def np_sum(vals):
    return int(sum(int(x) for x in np.asarray(vals, dtype=np.int64).tolist()))


# ---------------------------------------------------------------- DuckDB BitPacking column segments
BP_MODES = ["invalid", "auto", "constant", "constant_delta", "delta_for", "for"]


def bitpacking_decode(seg, elem_bytes, count):
    """decode one BitPacking segment (bytes as stored in the block) → (values, {mode: metadata groups})
    restates BitpackingScanPartial, src/storage/compression/bitpacking.cpp:776-860"""
    seg = np.ascontiguousarray(seg, dtype=np.uint8)
    out = np.empty(count, dtype=np.int64 if elem_bytes == 8 else np.int32)
    hist = np.zeros(6, dtype=np.uint64)
    rc = lib().oracle_bitpacking_decode(_p(seg, C.POINTER(C.c_uint8)), seg.size, elem_bytes, count,
                                        out.ctypes.data_as(C.c_void_p), _p(hist, _u64p))
    if rc != 0:
        raise ValueError("malformed BitPacking segment (oracle code %d)" % rc)
    return out, {BP_MODES[i]: int(hist[i]) for i in range(6) if hist[i]}


def decode_column_segments(blob, seg_meta, elem_bytes, n_rows):
    """a whole column from its segment directory rows (kind, row_start, count, offset, bytes, elem):
    kind 0 = Uncompressed (fixed_size_uncompressed.cpp), 1 = BitPacking, 2 = Constant (one value)"""
    out = np.empty(n_rows, dtype=np.int64 if elem_bytes == 8 else np.int32)
    modes = {}
    for kind, start, count, off, nbytes, elem in seg_meta:
        assert elem == elem_bytes
        raw = blob[off:off + nbytes]
        if kind == 0:
            out[start:start + count] = np.frombuffer(raw.tobytes(), dtype=out.dtype)[:count]
        elif kind == 2:
            out[start:start + count] = np.frombuffer(raw.tobytes(), dtype=out.dtype)[0]
        else:
            vals, h = bitpacking_decode(raw, elem_bytes, count)
            out[start:start + count] = vals
            for k, v in h.items():
                modes[k] = modes.get(k, 0) + v
    return out, modes


def bitpacking_encode(values, mode="auto"):
    """one BitPacking segment (uint8 array) holding all of `values` (int64 or int32), written the way the
    reference's BitpackingCompressState does (bitpacking.cpp:229-289,392-446,524-544)"""
    values = np.ascontiguousarray(values)
    assert values.dtype in (np.int64, np.int32)
    n_grp = (len(values) + 2047) // 2048
    out = np.empty(8 + n_grp * (24 + 2048 * 8 + 4) + 8, dtype=np.uint8)
    n = lib().oracle_bitpacking_encode(values.ctypes.data_as(C.c_void_p), len(values), values.dtype.itemsize,
                                       BP_MODES.index(mode), _p(out, C.POINTER(C.c_uint8)), out.size)
    if n < 0:
        raise ValueError("cannot encode (oracle code %d)" % n)
    return out[:n].copy()


def encode_column_segments(values, rows_per_segment=30720, mode="auto"):
    """a whole column as BitPacking segments → [(kind=1, row_start, count, bytes)]"""
    segs = []
    for s in range(0, len(values), rows_per_segment):
        part = values[s:s + rows_per_segment]
        segs.append((1, s, len(part), bitpacking_encode(part, mode)))
    return segs


# ---------------------------------------------------------------- WAH (FastBit ibis::bitvector) bitvectors
def wah_encode(words, n_bits):
    """verbatim bitvector (uint64 words, DuckDB bit order) → (wah uint32 words, active_val, active_nbits)"""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    out = np.empty(n_bits // 31 + 1, dtype=np.uint32)
    av, an = C.c_uint32(0), C.c_uint32(0)
    n = lib().oracle_wah_encode(_p(words, _u64p), n_bits, _p(out, C.POINTER(C.c_uint32)), out.size, C.byref(av),
                                C.byref(an))
    assert n >= 0
    return out[:n].copy(), av.value, an.value


def wah_decode(wah, active_val, active_nbits, n_words):
    """→ (verbatim uint64 words, bits described); raises on malformed input"""
    wah = np.ascontiguousarray(wah, dtype=np.uint32)
    out = np.empty(n_words, dtype=np.uint64)
    n = lib().oracle_wah_decode(_p(wah, C.POINTER(C.c_uint32)), wah.size, active_val, active_nbits, _p(out, _u64p),
                                n_words)
    if n < 0:
        raise ValueError("malformed WAH bitvector")
    return out, int(n)


def rle_decode(seg, elem_bytes, count):
    """one RLE column segment (rle.cpp layout) → (values, n_runs); raises ValueError when malformed"""
    seg = np.ascontiguousarray(seg, dtype=np.uint8)
    out = np.empty(count, dtype=np.int64 if elem_bytes == 8 else np.int32)
    runs = C.c_uint64(0)
    rc = lib().oracle_rle_decode(seg.ctypes.data_as(C.POINTER(C.c_uint8)), len(seg), elem_bytes, count, out.ctypes.data,
                                 C.byref(runs))
    if rc != 0:
        raise ValueError("malformed RLE segment")
    return out, int(runs.value)


def rle_encode(values):
    """values (int32 / int64) → one RLE segment as the reference writes it (uint8 array)"""
    values = np.ascontiguousarray(values)
    assert values.dtype.itemsize in (4, 8)
    cap = 16 + len(values) * (values.dtype.itemsize + 2) + 8
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().oracle_rle_encode(values.ctypes.data, len(values), values.dtype.itemsize,
                                out.ctypes.data_as(C.POINTER(C.c_uint8)), cap)
    assert n > 0
    return out[:n].copy()
