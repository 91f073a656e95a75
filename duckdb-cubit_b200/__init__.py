"""duckdb-cubit_b200 — B200-native CUBIT bitmap-index scan (merge → decode → probe).

This package is a thin ctypes binding over the C-ABI in include/cubit_gpu.h
(libcubit_gpu.so: hand-written sm_100a kernels + C++ host code).  It exists for the
tests and bench.py; the product is the shared library, and DuckDB binds to it from C++
(INTEGRATION.md).  There is no CPU fallback: loading fails loudly if the library has
not been built, and every compute call fails if no B200 is present.

The directory name has a hyphen, so import it with
    importlib.import_module("duckdb-cubit_b200")      (tests/conftest.py does this)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUBIT_GPU_LIB") or os.path.join(_HERE, "libcubit_gpu.so")  # override: kernel experiments
HOST_LIB_PATH = os.path.join(_HERE, "libcubit_host.so")

# ---- mirror of include/cubit_gpu.h -------------------------------------------------
ABI_VERSION = 5
OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5
MAX_STREAMS = 64
MAX_PROBE_COLS = 8
Q_ROWIDS, Q_BITVECTOR, Q_VALUES, Q_TIMING, Q_UNFUSED, Q_ASYNC, Q_FUSE_PROBE = 1, 2, 4, 8, 16, 32, 64
PROBE_NONE, PROBE_FUSED, PROBE_BITS, PROBE_GATHER, PROBE_DENSE = 0, 1, 2, 3, 4
SCAN_RING, SCAN_TWO_PASS, SCAN_NONE, SCAN_LOOKBACK = 0, 1, 2, 3
AGG_NONE, AGG_SUM, AGG_SUM_PROD, AGG_SUM_F64 = 0, 1, 2, 3

# every symbol include/cubit_gpu.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "cubit_gpu_abi_version", "cubit_gpu_last_error", "cubit_gpu_device_count", "cubit_gpu_create",
    "cubit_gpu_destroy", "cubit_gpu_set_stream", "cubit_gpu_words_per_bitvector", "cubit_gpu_launch_count",
    "cubit_gpu_index_create", "cubit_gpu_upload_bitvector", "cubit_gpu_download_bitvector", "cubit_gpu_index_build",
    "cubit_gpu_bitvector_count", "cubit_gpu_set_delta", "cubit_gpu_merge_deltas", "cubit_gpu_upload_column",
    "cubit_gpu_download_column", "cubit_gpu_synth_column", "cubit_gpu_drop_column", "cubit_gpu_pack_column",
    "cubit_gpu_upload_column_segments", "cubit_gpu_append_rows", "cubit_gpu_upload_bitvector_wah",
    "cubit_gpu_query",
    "cubit_gpu_result_wait", "cubit_gpu_result_get", "cubit_gpu_fetch", "cubit_gpu_fetch_bitvector",
    "cubit_gpu_free_result", "cubit_gpu_probe", "cubit_gpu_upload_column_validity", "cubit_gpu_fetch_validity",
    "cubit_gpu_index_serialize", "cubit_gpu_index_deserialize", "cubit_gpu_free_image",
    "cubit_gpu_alloc_host", "cubit_gpu_free_host",
    "cubit_gpu_create_sharded", "cubit_gpu_shard_count", "cubit_gpu_shard_info", "cubit_gpu_row_count",
    "cubit_gpu_index_create_compressed", "cubit_gpu_index_info", "cubit_gpu_add_delta", "cubit_gpu_add_delta_pairs",
    "cubit_gpu_set_merge_threshold", "cubit_gpu_fetch_async", "cubit_gpu_fetch_wait", "cubit_gpu_result_add_limbs",
    "cubit_gpu_fetch_wire_async", "cubit_gpu_wire_bytes", "cubit_gpu_wire_payload_bytes", "cubit_gpu_wire_unpack",
    "cubit_gpu_drain",
]


class BvRef(C.Structure):
    _fields_ = [("index_id", C.c_int32), ("value_id", C.c_uint32)]


class PredGroup(C.Structure):
    _fields_ = [("n_refs", C.c_uint32), ("refs", C.POINTER(BvRef))]


class Query(C.Structure):
    _fields_ = [("n_groups", C.c_uint32), ("groups", C.POINTER(PredGroup)), ("flags", C.c_uint32),
                ("n_cols", C.c_uint32), ("cols", C.POINTER(C.c_int32)), ("agg_kind", C.c_int32),
                ("agg_col_a", C.c_int32), ("agg_col_b", C.c_int32)]


WIRE_CHUNK = 2048
CHUNK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_int64),
                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))


class DrainStats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("chunks", C.c_uint64), ("windows", C.c_uint64), ("wire_bytes", C.c_uint64),
                ("wide_bytes", C.c_uint64), ("sum_rowids", C.c_uint64), ("sum_cols", C.c_uint64 * MAX_PROBE_COLS),
                ("workers", C.c_uint32), ("reserved", C.c_uint32)]


class ResultInfo(C.Structure):
    _fields_ = [("count", C.c_uint64), ("sum_lo", C.c_uint64), ("sum_hi", C.c_int64), ("capacity", C.c_uint64),
                ("n_streams", C.c_uint32), ("n_launches", C.c_uint32), ("delta_entries", C.c_uint64),
                ("algo_bytes_scan", C.c_uint64), ("algo_bytes_probe", C.c_uint64), ("ms_scan", C.c_float),
                ("ms_probe", C.c_float), ("ms_total", C.c_float), ("fused", C.c_uint32),
                ("d_rowids", C.c_void_p), ("d_bitvector", C.c_void_p), ("d_values", C.c_void_p * MAX_PROBE_COLS),
                ("sum_f64", C.c_double), ("agg_rows", C.c_uint64), ("d_validity", C.c_void_p * MAX_PROBE_COLS),
                ("probe_path", C.c_uint32), ("scan_path", C.c_uint32)]


class ColumnSegment(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("reserved", C.c_uint32), ("row_start", C.c_uint64), ("count", C.c_uint64),
                ("data", C.c_void_p), ("bytes", C.c_uint64)]


class WahBitvector(C.Structure):
    _fields_ = [("words", C.c_void_p), ("n_words", C.c_uint64), ("active_val", C.c_uint32), ("active_nbits", C.c_uint32)]


class AppendColumn(C.Structure):
    _fields_ = [("col_id", C.c_int32), ("elem_bytes", C.c_uint32), ("data", C.c_void_p)]


class IndexInfo(C.Structure):
    _fields_ = [("cardinality", C.c_uint32), ("compressed", C.c_uint32), ("resident_bytes", C.c_uint64),
                ("verbatim_bytes", C.c_uint64), ("delta_entries", C.c_uint64), ("auto_merges", C.c_uint64)]


class DecodeInfo(C.Structure):
    _fields_ = [("h2d_bytes", C.c_uint64), ("n_groups", C.c_uint64), ("mode_groups", C.c_uint64 * 6),
                ("rle_runs", C.c_uint64), ("n_launches", C.c_uint32), ("ms_decode", C.c_float)]


SEG_UNCOMPRESSED, SEG_BITPACKING, SEG_CONSTANT = 0, 1, 2


class CubitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cubit_gpu error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """dlopen libcubit_gpu.so.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built — run `python duckdb-cubit_b200/build.py` "
                          "(there is no CPU fallback for the CUBIT GPU path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, i64, u32, i32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_int32
    P = C.POINTER
    sig = {
        "cubit_gpu_abi_version": ([], C.c_int),
        "cubit_gpu_last_error": ([], C.c_char_p),
        "cubit_gpu_device_count": ([P(C.c_int)], C.c_int),
        "cubit_gpu_create": ([C.c_int, u64, i64, u32, P(vp)], C.c_int),
        "cubit_gpu_destroy": ([vp], C.c_int),
        "cubit_gpu_set_stream": ([vp, vp], C.c_int),
        "cubit_gpu_words_per_bitvector": ([vp, P(u64)], C.c_int),
        "cubit_gpu_launch_count": ([vp, P(u64)], C.c_int),
        "cubit_gpu_index_create": ([vp, u32, P(i32)], C.c_int),
        "cubit_gpu_upload_bitvector": ([vp, i32, u32, vp, u64], C.c_int),
        "cubit_gpu_download_bitvector": ([vp, i32, u32, vp, u64], C.c_int),
        "cubit_gpu_index_build": ([vp, i32, i32, i64], C.c_int),
        "cubit_gpu_bitvector_count": ([vp, i32, u32, P(u64)], C.c_int),
        "cubit_gpu_set_delta": ([vp, i32, u32, vp, u64], C.c_int),
        "cubit_gpu_merge_deltas": ([vp, i32], C.c_int),
        "cubit_gpu_upload_column": ([vp, i32, vp, u32, u64], C.c_int),
        "cubit_gpu_download_column": ([vp, i32, vp, u32, u64], C.c_int),
        "cubit_gpu_synth_column": ([vp, i32, i32, u64, u64, u32, u32, u32], C.c_int),
        "cubit_gpu_drop_column": ([vp, i32], C.c_int),
        "cubit_gpu_pack_column": ([vp, i32, C.c_int, P(u64)], C.c_int),
        "cubit_gpu_upload_bitvector_wah": ([vp, i32, u32, P(WahBitvector)], C.c_int),
        "cubit_gpu_append_rows": ([vp, u64, P(AppendColumn), u32], C.c_int),
        "cubit_gpu_upload_column_segments": ([vp, i32, u32, P(ColumnSegment), u32, P(DecodeInfo)], C.c_int),
        "cubit_gpu_query": ([vp, P(Query), P(vp)], C.c_int),
        "cubit_gpu_result_wait": ([vp], C.c_int),
        "cubit_gpu_result_get": ([vp, P(ResultInfo)], C.c_int),
        "cubit_gpu_fetch": ([vp, u64, u64, vp, u32, P(vp)], C.c_int),
        "cubit_gpu_fetch_bitvector": ([vp, vp, u64], C.c_int),
        "cubit_gpu_free_result": ([vp], C.c_int),
        "cubit_gpu_probe": ([vp, i32, vp, u64, vp, P(u64), P(i64)], C.c_int),
        "cubit_gpu_upload_column_validity": ([vp, i32, vp, u64], C.c_int),
        "cubit_gpu_fetch_validity": ([vp, u32, u64, u64, vp, P(C.c_int)], C.c_int),
        "cubit_gpu_index_serialize": ([vp, i32, P(vp), P(u64)], C.c_int),
        "cubit_gpu_index_deserialize": ([vp, vp, u64, P(i32)], C.c_int),
        "cubit_gpu_free_image": ([vp], None),
        "cubit_gpu_alloc_host": ([u64, P(vp)], C.c_int),
        "cubit_gpu_free_host": ([vp], C.c_int),
        "cubit_gpu_create_sharded": ([P(C.c_int), u32, u64, i64, u32, P(vp)], C.c_int),
        "cubit_gpu_shard_count": ([vp, P(u32)], C.c_int),
        "cubit_gpu_shard_info": ([vp, u32, P(C.c_int), P(u64), P(u64)], C.c_int),
        "cubit_gpu_row_count": ([vp, P(u64)], C.c_int),
        "cubit_gpu_index_create_compressed": ([vp, u32, P(i32)], C.c_int),
        "cubit_gpu_index_info": ([vp, i32, P(IndexInfo)], C.c_int),
        "cubit_gpu_add_delta": ([vp, i32, u32, vp, u64], C.c_int),
        "cubit_gpu_add_delta_pairs": ([vp, i32, vp, vp, u64], C.c_int),
        "cubit_gpu_set_merge_threshold": ([vp, i32, C.c_double], C.c_int),
        "cubit_gpu_fetch_async": ([vp, u64, u64, vp, u32, P(vp), P(vp)], C.c_int),
        "cubit_gpu_fetch_wait": ([vp], C.c_int),
        "cubit_gpu_result_add_limbs": ([vp, vp], C.c_int),
        "cubit_gpu_fetch_wire_async": ([vp, u64, u64, C.c_int, u32, vp, u64, P(vp)], C.c_int),
        "cubit_gpu_wire_bytes": ([u64, u32], u64),
        "cubit_gpu_wire_payload_bytes": ([vp], u64),
        "cubit_gpu_wire_unpack": ([vp, u32, u64, vp, u32], C.c_int),
        "cubit_gpu_drain": ([vp, C.c_int, u32, u32, u64, vp, vp, P(DrainStats)], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    if L.cubit_gpu_abi_version() != ABI_VERSION:
        raise ImportError("libcubit_gpu.so ABI %d != binding %d" % (L.cubit_gpu_abi_version(), ABI_VERSION))
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise CubitError(rc, load_library().cubit_gpu_last_error().decode())


class HostBuffer:
    """page-locked host memory from cubit_gpu_alloc_host as a uint8 array (`.array`); free() or context manager"""

    def __init__(self, nbytes):
        self._p = C.c_void_p()
        _check(load_library().cubit_gpu_alloc_host(nbytes, C.byref(self._p)))
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, nbytes)).from_address(self._p.value))

    def free(self):
        if self._p is not None:
            self.array = None
            load_library().cubit_gpu_free_host(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()


def wire_bytes(n_rows, n_streams):
    return int(load_library().cubit_gpu_wire_bytes(n_rows, n_streams))


def wire_unpack(wire, stream, chunk, dtype=np.int64):
    """widen one DataChunk of one stream of a narrow wire (host-only code) → array"""
    out = np.empty(WIRE_CHUNK, dtype=dtype)
    n = load_library().cubit_gpu_wire_unpack(wire.ctypes.data, stream, chunk, out.ctypes.data, out.dtype.itemsize)
    if n < 0:
        raise CubitError(EINVAL, load_library().cubit_gpu_last_error().decode())
    return out[:n]


def device_count():
    n = C.c_int(0)
    rc = load_library().cubit_gpu_device_count(C.byref(n))
    return n.value if rc == OK else 0


def int128(lo, hi):
    return (int(hi) << 64) + int(lo)


class QueryPlan:
    """A cubit_query kept alive on the host side (ctypes arrays) so it can be re-issued cheaply."""

    def __init__(self, groups, flags=Q_ROWIDS, cols=(), agg=AGG_NONE, agg_a=-1, agg_b=-1):
        self._ref_arrays = []
        self._groups = (PredGroup * len(groups))()
        for g, grp in enumerate(groups):
            arr = (BvRef * len(grp))()
            for i, (ix, v) in enumerate(grp):
                arr[i].index_id, arr[i].value_id = ix, v
            self._ref_arrays.append(arr)
            self._groups[g].n_refs = len(grp)
            self._groups[g].refs = arr
        self._cols = (C.c_int32 * max(1, len(cols)))(*cols)
        self.q = Query(len(groups), self._groups, flags, len(cols), self._cols, agg, agg_a, agg_b)


class Result:
    """One query's result set (cubit_gpu_result)."""

    def __init__(self, table, handle, col_dtypes):
        self._t, self._h, self._dtypes = table, handle, col_dtypes
        self._info = None

    def wait(self):
        _check(self._t._L.cubit_gpu_result_wait(self._h))
        return self

    @property
    def info(self):
        if self._info is None:
            inf = ResultInfo()
            _check(self._t._L.cubit_gpu_result_get(self._h, C.byref(inf)))
            self._info = inf
        return self._info

    @property
    def count(self):
        return int(self.info.count)

    @property
    def sum(self):
        return int128(self.info.sum_lo, self.info.sum_hi)

    @property
    def sum_f64(self):
        return float(self.info.sum_f64)

    def fetch(self, offset=0, n=None, rowids=True, out_ids=None, out_cols=None):
        """→ (row_ids or None, [column arrays]) for result rows [offset, offset+n)"""
        if n is None:
            n = self.count - offset
        ids = None
        if rowids:
            ids = out_ids[:n] if out_ids is not None else np.empty(n, dtype=np.int64)
        cols = []
        ptrs = (C.c_void_p * max(1, len(self._dtypes)))()
        for c, dt in enumerate(self._dtypes):
            a = out_cols[c][:n] if out_cols is not None else np.empty(n, dtype=dt)
            cols.append(a)
            ptrs[c] = a.ctypes.data
        _check(self._t._L.cubit_gpu_fetch(self._h, offset, n, ids.ctypes.data if ids is not None else None,
                                          len(self._dtypes), ptrs))
        return ids, cols

    def fetch_async(self, offset, n, out_ids=None, out_cols=()):
        """enqueue the copy of result rows [offset, offset+n) into (page-locked) arrays → ticket for fetch_wait"""
        ptrs = (C.c_void_p * max(1, len(out_cols)))()
        for c, a in enumerate(out_cols):
            ptrs[c] = a.ctypes.data
        tk = C.c_void_p()
        _check(self._t._L.cubit_gpu_fetch_async(self._h, offset, n, out_ids.ctypes.data if out_ids is not None else None,
                                                len(out_cols), ptrs, C.byref(tk)))
        return tk

    def fetch_wait(self, ticket):
        _check(self._t._L.cubit_gpu_fetch_wait(ticket))

    def fetch_wire_async(self, offset, n, wire, rowids=True, n_cols=None):
        """enqueue the narrow-wire hand-off of result rows [offset, offset+n) into the page-locked uint8 array `wire`
        (include/cubit_gpu_wire.h) → ticket for fetch_wait"""
        tk = C.c_void_p()
        nc = len(self._dtypes) if n_cols is None else n_cols
        _check(self._t._L.cubit_gpu_fetch_wire_async(self._h, offset, n, 1 if rowids else 0, nc, wire.ctypes.data,
                                                     wire.nbytes, C.byref(tk)))
        return tk

    def drain(self, rowids=True, n_cols=None, threads=1, window_rows=0, fn=None, validity=False):
        """the parallel ordered DataChunk hand-off over the narrow wire (cubit_gpu_drain).  fn(worker, batch_index,
        row_offset, ids or None, [column arrays]) is called per DataChunk (copies); with validity=True a sixth
        argument carries, per column, the chunk's ValidityMask words or None.  fn=None = checksum consumer.
        → DrainStats"""
        nc = len(self._dtypes) if n_cols is None else n_cols
        st = DrainStats()
        cb = None
        if fn is not None:
            dts = self._dtypes[:nc]

            def tramp(ctx, worker, batch, row_off, n, ids_p, cols_p, val_p):
                try:
                    ids = np.ctypeslib.as_array(ids_p, shape=(n,)).copy() if (rowids and n) else (
                        np.empty(0, np.int64) if rowids else None)
                    cols = []
                    for c, dt in enumerate(dts):
                        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(cols_p[c]) if n else b""
                        cols.append(np.frombuffer(buf, dtype=dt, count=n).copy())
                    if validity:
                        vw = [np.ctypeslib.as_array(C.cast(val_p[c], C.POINTER(C.c_uint64)),
                                                    shape=((n + 63) // 64,)).copy() if val_p[c] else None
                              for c in range(len(dts))]
                        return int(bool(fn(worker, batch, row_off, ids, cols, vw)))
                    return int(bool(fn(worker, batch, row_off, ids, cols)))
                except Exception:  # an exception must not unwind through the C frames
                    import traceback
                    traceback.print_exc()
                    return 1
            cb = CHUNK_FN(tramp)
        _check(self._t._L.cubit_gpu_drain(self._h, 1 if rowids else 0, nc, threads, window_rows,
                                          C.cast(cb, C.c_void_p) if cb is not None else None, None, C.byref(st)))
        return st

    def add_limbs(self, device_ptr):
        """ADD (count, 128-bit sum) as five int64 limbs to device memory, in stream order (multi-process reduce)"""
        _check(self._t._L.cubit_gpu_result_add_limbs(self._h, C.c_void_p(device_ptr)))

    def fetch_validity(self, col, offset=0, n=None):
        """→ (uint64 mask words with bit j = result row offset+j valid, all_valid) for projected column `col`"""
        if n is None:
            n = self.count - offset
        words = np.zeros(max(1, (n + 63) // 64), dtype=np.uint64)
        allv = C.c_int(0)
        _check(self._t._L.cubit_gpu_fetch_validity(self._h, col, offset, n, words.ctypes.data, C.byref(allv)))
        return words[:(n + 63) // 64], bool(allv.value)

    @property
    def agg_rows(self):
        return int(self.info.agg_rows)

    def bitvector(self):
        q = np.empty(self._t.n_words, dtype=np.uint64)
        _check(self._t._L.cubit_gpu_fetch_bitvector(self._h, q.ctypes.data, len(q)))
        return q

    def free(self):
        if self._h is not None:
            self._t._L.cubit_gpu_free_result(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CubitTable:
    """One table shard resident on one B200 (cubit_gpu_table)."""

    def __init__(self, n_rows, row_base=0, seg_bits=65536, device=0, devices=None):
        """devices=[d0, d1, ...]: one table sharded by row range over those GPUs (cubit_gpu_create_sharded)"""
        self._L = load_library()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(self._L.cubit_gpu_create_sharded(arr, len(devices), n_rows, row_base, seg_bits, C.byref(h)))
        else:
            _check(self._L.cubit_gpu_create(device, n_rows, row_base, seg_bits, C.byref(h)))
        self._h = h
        self.n_rows, self.row_base, self.seg_bits, self.device = n_rows, row_base, seg_bits, device
        self.n_words = (n_rows + 63) // 64
        self._col_dtype = {}

    def close(self):
        if self._h is not None:
            self._L.cubit_gpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self._L.cubit_gpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    @property
    def launch_count(self):
        n = C.c_uint64(0)
        _check(self._L.cubit_gpu_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- index
    def create_index(self, cardinality, compressed=False):
        """compressed=True: roaring-style containers in HBM (cubit_gpu_index_create_compressed)"""
        ix = C.c_int32(-1)
        fn = self._L.cubit_gpu_index_create_compressed if compressed else self._L.cubit_gpu_index_create
        _check(fn(self._h, cardinality, C.byref(ix)))
        return ix.value

    def index_info(self, index_id):
        info = IndexInfo()
        _check(self._L.cubit_gpu_index_info(self._h, index_id, C.byref(info)))
        return info

    @property
    def shard_count(self):
        n = C.c_uint32(0)
        _check(self._L.cubit_gpu_shard_count(self._h, C.byref(n)))
        return n.value

    def shard_info(self, shard):
        dev, r0, n = C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        _check(self._L.cubit_gpu_shard_info(self._h, shard, C.byref(dev), C.byref(r0), C.byref(n)))
        return dev.value, r0.value, n.value

    def upload_bitvector(self, index_id, value_id, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        _check(self._L.cubit_gpu_upload_bitvector(self._h, index_id, value_id, words.ctypes.data, len(words)))

    def upload_bitvector_wah(self, index_id, value_id, wah_words, active_val=0, active_nbits=0):
        """upload a WAH-compressed (FastBit ibis::bitvector) value bitvector; expanded on the GPU"""
        w = np.ascontiguousarray(wah_words, dtype=np.uint32)
        bv = WahBitvector(w.ctypes.data if len(w) else None, len(w), active_val, active_nbits)
        _check(self._L.cubit_gpu_upload_bitvector_wah(self._h, index_id, value_id, C.byref(bv)))

    def upload_index(self, bitvectors, compressed=False):
        """bitvectors: [card, n_words] uint64 → new index id"""
        ix = self.create_index(len(bitvectors), compressed)
        for v, w in enumerate(bitvectors):
            self.upload_bitvector(ix, v, w)
        return ix

    def download_bitvector(self, index_id, value_id):
        out = np.empty(self.n_words, dtype=np.uint64)
        _check(self._L.cubit_gpu_download_bitvector(self._h, index_id, value_id, out.ctypes.data, len(out)))
        return out

    def serialize_index(self, index_id):
        """→ bytes: the persistent image of one index (bitvectors verbatim / WAH, pending deltas, checksum)"""
        img, n = C.c_void_p(None), C.c_uint64(0)
        _check(self._L.cubit_gpu_index_serialize(self._h, index_id, C.byref(img), C.byref(n)))
        try:
            return C.string_at(img.value, n.value)
        finally:
            self._L.cubit_gpu_free_image(img)

    def deserialize_index(self, image):
        """recreate an index from serialize_index's image → new index id"""
        buf = np.frombuffer(image, dtype=np.uint8)
        ix = C.c_int32(-1)
        _check(self._L.cubit_gpu_index_deserialize(self._h, buf.ctypes.data, len(buf), C.byref(ix)))
        return ix.value

    def build_index(self, index_id, col_id, base_value=0):
        _check(self._L.cubit_gpu_index_build(self._h, index_id, col_id, base_value))

    def bitvector_count(self, index_id, value_id):
        n = C.c_uint64(0)
        _check(self._L.cubit_gpu_bitvector_count(self._h, index_id, value_id, C.byref(n)))
        return n.value

    def set_delta(self, index_id, value_id, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        _check(self._L.cubit_gpu_set_delta(self._h, index_id, value_id, rows.ctypes.data if len(rows) else None,
                                           len(rows)))

    def add_delta(self, index_id, value_id, rows):
        """incremental: the rows join the pending delta of (index, value); ingested on the device"""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        _check(self._L.cubit_gpu_add_delta(self._h, index_id, value_id, rows.ctypes.data if len(rows) else None,
                                           len(rows)))

    def add_delta_pairs(self, index_id, value_ids, rows):
        """incremental: (value, row) pairs — what one UPDATE / DELETE statement produces"""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        vals = np.ascontiguousarray(value_ids, dtype=np.uint32)
        assert len(rows) == len(vals)
        _check(self._L.cubit_gpu_add_delta_pairs(self._h, index_id, vals.ctypes.data if len(rows) else None,
                                                 rows.ctypes.data if len(rows) else None, len(rows)))

    def set_merge_threshold(self, index_id, fraction):
        _check(self._L.cubit_gpu_set_merge_threshold(self._h, index_id, float(fraction)))

    def merge_deltas(self, index_id):
        _check(self._L.cubit_gpu_merge_deltas(self._h, index_id))

    # ---- columns
    def upload_column(self, col_id, data):
        data = np.ascontiguousarray(data)
        if data.dtype.itemsize not in (4, 8):
            raise ValueError("columns are 4 or 8 bytes wide")
        _check(self._L.cubit_gpu_upload_column(self._h, col_id, data.ctypes.data, data.dtype.itemsize, len(data)))
        self._col_dtype[col_id] = data.dtype

    def upload_validity(self, col_id, words):
        """validity mask of a column (DuckDB ValidityMask words, bit = 1: valid); None drops it"""
        if words is None:
            _check(self._L.cubit_gpu_upload_column_validity(self._h, col_id, None, 0))
            return
        words = np.ascontiguousarray(words, dtype=np.uint64)
        _check(self._L.cubit_gpu_upload_column_validity(self._h, col_id, words.ctypes.data, len(words)))

    def upload_column_segments(self, col_id, elem_bytes, segments):
        """upload a column as the reference's on-disk segments [(kind, row_start, count, uint8 array)] and decode
        them on the GPU → DecodeInfo"""
        arr = (ColumnSegment * max(1, len(segments)))()
        keep = []
        for i, (kind, start, count, data) in enumerate(segments):
            data = np.ascontiguousarray(data).view(np.uint8)
            keep.append(data)
            arr[i] = ColumnSegment(kind, 0, start, count, data.ctypes.data, data.size)
        info = DecodeInfo()
        _check(self._L.cubit_gpu_upload_column_segments(self._h, col_id, elem_bytes, arr, len(segments), C.byref(info)))
        self._col_dtype[col_id] = np.dtype(np.int64 if elem_bytes == 8 else np.int32)
        return info

    def append_rows(self, columns):
        """INSERT: append rows at the end of the shard; columns = {col_id: array of the new rows' values}"""
        arrs = {c: np.ascontiguousarray(a) for c, a in columns.items()}
        n_new = len(next(iter(arrs.values()))) if arrs else 0
        arr = (AppendColumn * max(1, len(arrs)))()
        for i, (c, a) in enumerate(arrs.items()):
            if len(a) != n_new:
                raise ValueError("append columns differ in length")
            arr[i] = AppendColumn(c, a.dtype.itemsize, a.ctypes.data)
        _check(self._L.cubit_gpu_append_rows(self._h, n_new, arr, len(arrs)))
        self.n_rows += n_new
        n = C.c_uint64(0)
        _check(self._L.cubit_gpu_words_per_bitvector(self._h, C.byref(n)))
        self.n_words = n.value

    def download_column(self, col_id):
        dt = self._col_dtype[col_id]
        out = np.empty(self.n_rows, dtype=dt)
        _check(self._L.cubit_gpu_download_column(self._h, col_id, out.ctypes.data, dt.itemsize, len(out)))
        return out

    def synth_column(self, col_id, kind, seed=0, threshold=0, card=100, hot_lo=10, hot_n=10):
        _check(self._L.cubit_gpu_synth_column(self._h, col_id, kind, seed, threshold, card, hot_lo, hot_n))
        self._col_dtype[col_id] = np.dtype(np.int64 if kind in (0, 3) else np.int32)

    def pack_column(self, col_id, keep_raw=False):
        """store an int64 column FOR-bit-packed in HBM (lossless) → resident bytes"""
        b = C.c_uint64(0)
        _check(self._L.cubit_gpu_pack_column(self._h, col_id, int(keep_raw), C.byref(b)))  # 0 / 1 / 2 (see the header)
        return b.value

    def drop_column(self, col_id):
        _check(self._L.cubit_gpu_drop_column(self._h, col_id))
        self._col_dtype.pop(col_id, None)

    # ---- query
    def execute(self, plan):
        """issue a prepared QueryPlan → Result"""
        h = C.c_void_p()
        _check(self._L.cubit_gpu_query(self._h, C.byref(plan.q), C.byref(h)))
        ncols = plan.q.n_cols if (plan.q.flags & Q_VALUES) else 0
        return Result(self, h, [self._col_dtype[plan._cols[c]] for c in range(ncols)])

    def query(self, groups, flags=Q_ROWIDS, cols=(), agg=AGG_NONE, agg_a=-1, agg_b=-1):
        """groups: [[(index_id, value_id), ...], ...]  — OR inside a group, AND across groups"""
        return self.execute(QueryPlan(groups, flags, cols, agg, agg_a, agg_b))

    def probe(self, col_id, row_ids, want_sum=False):
        """gather col[row_ids] (the DataTable::Fetch analog) → (values, sum or None)"""
        row_ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        dt = self._col_dtype[col_id]
        out = np.empty(len(row_ids), dtype=dt)
        lo, hi = C.c_uint64(0), C.c_int64(0)
        _check(self._L.cubit_gpu_probe(self._h, col_id, row_ids.ctypes.data if len(row_ids) else None, len(row_ids),
                                       out.ctypes.data, C.byref(lo) if want_sum else None,
                                       C.byref(hi) if want_sum else None))
        return out, (int128(lo.value, hi.value) if want_sum else None)
