#!/usr/bin/env python
"""Generate tests/golden/rle_segments.npz by running the REFERENCE itself (build container only).

Like make_bitpacking_golden.py: the reference shell creates tables under PRAGMA force_compression='rle',
checkpoints, and the COMPRESSED RLE column segments are lifted straight out of the database file
(pragma_storage_info → block_id / block_offset; file offset = 3*4096 + block_id*262144 + 8 + block_offset).
An RLE segment does not store its size: it is [u64 offset of the run lengths][values][pad][u16 run lengths]
(src/storage/compression/rle.cpp:190-205) and its run count follows from the row count, so the bytes kept are
offset + 2 * (runs needed to cover the segment's rows).  The values the reference returns for
`SELECT v FROM t ORDER BY rowid` are stored next to them.

Usage:  python tests/golden/make_rle_golden.py [--duckdb /path/to/duckdb]
"""
import argparse
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK_START, BLOCK_ALLOC, BLOCK_HEADER = 3 * 4096, 262144, 8

# column → (SQL type, expression over i, rows)
COLUMNS = [
    ("r_runs64", "BIGINT", "((i // 37) * 7919) % 1000 - 500", 131_075),          # runs of 37, two row groups
    ("r_runs32", "INTEGER", "((i // 5) * 69069) % 70001 - 35000", 50_003),       # short runs, 4-byte values (padding)
    ("r_date", "DATE", "DATE '1992-01-02' + CAST((i // 1000) % 2526 AS INTEGER)", 50_003),
    ("r_long", "BIGINT", "CASE WHEN i < 100000 THEN -42 WHEN i < 130000 THEN 4611686018427387904 ELSE 9 END", 131_075),  # runs > 65535 are split
    ("r_single", "BIGINT", "(i * 2654435761) % 1000003", 40_001),                # every run has length 1 (worst case)
    ("r_mixed", "INTEGER", "CASE WHEN i % 4096 < 4000 THEN 7 ELSE i END", 131_075),
]


def run(duck, db, stmt):
    r = subprocess.run([duck, db, "-csv", "-noheader", "-c", stmt], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("duckdb failed: %s\n%s" % (stmt, r.stderr))
    return r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duckdb", default="/tmp/duckdb_build/duckdb")
    ap.add_argument("--out", default=os.path.join(HERE, "rle_segments.npz"))
    args = ap.parse_args()
    out = {}
    seg_dir = []  # (column, row_start, count, byte offset into blob, bytes, elem_bytes)
    blob = bytearray()
    with tempfile.TemporaryDirectory() as tmp:
        for name, typ, expr, n_rows in COLUMNS:
            db = os.path.join(tmp, name + ".db")
            e = expr if typ == "DATE" else "CAST((%s) AS %s)" % (expr, typ)
            run(args.duckdb, db, "PRAGMA force_compression='rle'; CREATE TABLE t AS SELECT %s AS v FROM "
                "(SELECT CAST(range AS HUGEINT) AS i FROM range(0, %d)); CHECKPOINT;" % (e, n_rows))
            sel = "v - DATE '1970-01-01'" if typ == "DATE" else "v"
            vals = np.array([int(x) for x in run(args.duckdb, db, "SELECT %s FROM t ORDER BY rowid" % sel).split()],
                            dtype=np.int64)
            assert len(vals) == n_rows
            elem = 8 if typ == "BIGINT" else 4
            out["values_" + name] = vals.astype(np.int64 if elem == 8 else np.int32)
            info = run(args.duckdb, db, "SELECT start, count, compression, block_id, block_offset FROM "
                       "pragma_storage_info('t') WHERE segment_type <> 'VALIDITY' AND column_name = 'v' ORDER BY start")
            raw = open(db, "rb").read()
            for line in info.strip().splitlines():
                start, count, comp, block_id, block_off = line.split(",")
                start, count, block_id, block_off = int(start), int(count), int(block_id), int(block_off)
                assert comp == "RLE", (name, comp)
                base = BLOCK_START + block_id * BLOCK_ALLOC + BLOCK_HEADER + block_off
                off = int(np.frombuffer(raw[base:base + 8], dtype="<u8")[0])
                counts = np.frombuffer(raw[base + off:base + off + 2 * ((off - 8) // elem)], dtype="<u2")
                n_runs = int(np.searchsorted(np.cumsum(counts.astype(np.int64)), count)) + 1
                assert int(counts[:n_runs].sum()) == count, (name, start)
                size = off + 2 * n_runs
                data = raw[base:base + size]
                while len(blob) % 8:
                    blob.append(0)
                seg_dir.append((name, start, count, len(blob), len(data), elem))
                blob += data
                print("%-9s RLE rows [%7d, +%6d) %6d runs %7d bytes (max run %d)" %
                      (name, start, count, n_runs, size, int(counts[:n_runs].max())))
    out["blob"] = np.frombuffer(bytes(blob), dtype=np.uint8)
    out["seg_column"] = np.array([s[0] for s in seg_dir])
    out["seg_meta"] = np.array([s[1:] for s in seg_dir], dtype=np.int64)  # row_start,count,offset,bytes,elem
    np.savez_compressed(args.out, **out)
    print("wrote %s (%d bytes, %d segments)" % (args.out, os.path.getsize(args.out), len(seg_dir)))


if __name__ == "__main__":
    main()
