#!/usr/bin/env python
"""Generate tests/golden/bitpacking_segments.npz by running the REFERENCE itself.

The reference shell (DuckDB built from /root/reference, /tmp/duckdb_build/duckdb in the build container)
creates a table whose numeric columns exercise every BitPacking mode (src/storage/compression/
bitpacking.cpp: CONSTANT, CONSTANT_DELTA, DELTA_FOR, FOR; BIGINT, INTEGER and DATE physical types; forced
modes through PRAGMA force_bitpacking_mode), checkpoints it, and this script then lifts the COMPRESSED
column segments straight out of the database file:

    pragma_storage_info('t')  →  (column, row group, start, count, compression, block_id, block_offset)
    file offset of a block    =  3 * 4096 + block_id * 262144          (single_file_block_manager.hpp:35,
                                                                        single_file_block_manager.cpp:443-445)
    segment bytes             =  block data (after the 8-byte checksum, storage_info.hpp:40) at block_offset;
                                 a BitPacking segment starts with the u64 offset of the end of its metadata
                                 (bitpacking.cpp:524-544), which is its size

and stores them next to the values the reference returns for `SELECT col FROM t` (rowid order).  The CPU
oracle's decoder (oracle_bitpacking_decode) and the GPU decode kernel are both checked against these.

Usage:  python tests/golden/make_bitpacking_golden.py [--duckdb /path/to/duckdb]
"""
import argparse
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK_START, BLOCK_ALLOC, BLOCK_HEADER = 3 * 4096, 262144, 8
SMALL_ROWS = 50_003   # several segments for the wide columns, partial last metadata group
N_ROWS = 131_075  # one full row group (122,880) + a ragged second one whose last metadata group is partial

# column name → (SQL type, expression over i, forced bitpacking mode or None)
COLUMNS = [
    ("c_for", "BIGINT", "((i * 7919) % 100003) + 5000000000", None),
    ("c_seq", "BIGINT", "i * 3 - 1000", None),                                   # CONSTANT_DELTA
    ("c_const", "BIGINT", "CASE WHEN i < 122880 THEN 42 ELSE 40 END", None),
    ("c_sorted", "BIGINT", "i * 1000003 + ((i * 2654435761) % 977)", None),      # DELTA_FOR (small deltas)
    ("c_neg", "BIGINT", "((i * 48271) % 2147483647) - 1073741823 - (i % 7) * 4611686018427387", None),
    ("c_price", "BIGINT", "90000 + ((i * 1103515245 + 12345) % 10404951)", None),  # l_extendedprice-like cents
    ("c_disc", "BIGINT", "(i * 31 + i // 11) % 11", None),                        # l_discount-like
    ("c_date", "DATE", "DATE '1992-01-02' + CAST((i * 17) % 2526 AS INTEGER)", None),
    ("c_i32", "INTEGER", "((i * 69069) % 70001) - 35000", None),
    ("f_for", "BIGINT", "((i * 7919) % 100003) - 50000", "for"),
    ("f_delta", "BIGINT", "((i * 7919) % 100003) - 50000", "delta_for"),         # negative deltas, wraps
    ("f_delta32", "INTEGER", "((i * 69069) % 70001) - 35000", "delta_for"),
    ("f_for32", "INTEGER", "((i * 69069) % 70001) - 35000", "for"),
    ("f_const", "BIGINT", "-77 + (i // 20480) * 1000000007", "constant"),        # CONSTANT groups, several values
    ("u_plain", "INTEGER", "(i * 7) % 1000 - 500", "uncompressed"),              # Uncompressed segments
    ("f_wide", "BIGINT", "(i * 6364136223846793005 + 1442695040888963407) % 9223372036854775807", "for"),
]


def run(duck, db, stmt, csv=True):
    cmd = [duck, db] + (["-csv", "-noheader"] if csv else []) + ["-c", stmt]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError("duckdb failed: %s\n%s" % (stmt, r.stderr))
    return r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duckdb", default="/tmp/duckdb_build/duckdb")
    ap.add_argument("--out", default=os.path.join(HERE, "bitpacking_segments.npz"))
    args = ap.parse_args()
    out = {}
    seg_dir = []  # (column, kind, row_start, count, byte_offset into blob, bytes, elem_bytes)
    blob = bytearray()
    with tempfile.TemporaryDirectory() as tmp:
        for name, typ, expr, mode in COLUMNS:
            n_rows = N_ROWS if name in ("c_seq", "c_const", "c_disc", "c_sorted") else SMALL_ROWS  # keeps the fixture small
            db = os.path.join(tmp, name + ".db")
            if mode == "uncompressed":
                pre = "PRAGMA force_compression='uncompressed';"
            else:
                pre = ("PRAGMA force_compression='bitpacking'; PRAGMA force_bitpacking_mode='%s';" % mode) if mode else ""
            # i is a HUGEINT so the generator expressions cannot overflow before the final cast
            e = expr if typ == "DATE" else "CAST((%s) AS %s)" % (expr, typ)
            run(args.duckdb, db, pre + "CREATE TABLE t AS SELECT %s AS v FROM (SELECT CAST(range AS HUGEINT) AS i "
                "FROM range(0, %d)); CHECKPOINT;" % (e, n_rows))
            sel = "v - DATE '1970-01-01'" if typ == "DATE" else "v"
            vals = np.array([int(x) for x in run(args.duckdb, db, "SELECT %s FROM t ORDER BY rowid" % sel).split()],
                            dtype=np.int64)
            assert len(vals) == n_rows
            elem = 8 if typ == "BIGINT" else 4
            out["values_" + name] = vals.astype(np.int64 if elem == 8 else np.int32)
            info = run(args.duckdb, db,
                       "SELECT start, count, compression, block_id, block_offset, stats FROM pragma_storage_info('t') "
                       "WHERE segment_type <> 'VALIDITY' AND column_name = 'v' ORDER BY start")
            raw = open(db, "rb").read()
            for line in info.strip().splitlines():
                m = re.match(r'(\d+),(\d+),(\w+),(-?\d+),(-?\d+),"?(.*?)"?$', line)
                start, count, comp, block_id, block_off, stats = m.groups()
                start, count, block_id, block_off = int(start), int(count), int(block_id), int(block_off)
                if comp == "Constant":
                    mn = int(re.search(r"Min: (-?\d+)", stats).group(1)) if typ != "DATE" else None
                    if typ == "DATE":
                        mn = int(vals[start])  # the stats print a calendar date; the value is the same constant
                    data = np.array([mn], dtype="<i8" if elem == 8 else "<i4").tobytes()
                    kind = 2
                else:
                    base = BLOCK_START + block_id * BLOCK_ALLOC + BLOCK_HEADER + block_off
                    if comp == "BitPacking":
                        size = int(np.frombuffer(raw[base:base + 8], dtype="<u8")[0])
                        kind = 1
                    elif comp == "Uncompressed":
                        size = count * elem
                        kind = 0
                    else:
                        raise RuntimeError("unexpected compression %s for %s" % (comp, name))
                    data = raw[base:base + size]
                while len(blob) % 8:
                    blob.append(0)
                seg_dir.append((name, kind, start, count, len(blob), len(data), elem))
                blob += data
                print("%-10s %-12s rows [%7d, +%6d)  %7d bytes  (%.2f bits/value)" %
                      (name, comp, start, count, len(data), 8.0 * len(data) / count))
    out["blob"] = np.frombuffer(bytes(blob), dtype=np.uint8)
    out["seg_column"] = np.array([s[0] for s in seg_dir])
    out["seg_meta"] = np.array([s[1:] for s in seg_dir], dtype=np.int64)  # kind,row_start,count,offset,bytes,elem
    np.savez_compressed(args.out, **out)
    print("wrote %s (%d bytes, %d segments)" % (args.out, os.path.getsize(args.out), len(seg_dir)))


if __name__ == "__main__":
    main()
