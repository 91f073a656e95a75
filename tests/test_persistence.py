"""Index persistence (BoundIndex::GetStorageInfo / IndexStorageInfo, bound_index.hpp:117-118):
serialize → image → deserialize into another table shard reproduces the bitvectors, the pending deltas and every
query answer; the WAH payloads inside the image are exactly what the oracle's WAH writer produces; damaged images
are rejected before anything is uploaded."""
import struct

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

HDR = struct.Struct("<8sQIiqII")    # magic, n_rows, card, src_col, src_base, flags (1 = compressed), pad
ENT = struct.Struct("<IIIIQQ")      # encoding, active_val, active_nbits, pad, n_words, n_delta_rows


def parse(image):
    magic, n_rows, card, src_col, src_base, _flags, _pad = HDR.unpack_from(image, 0)
    assert magic == b"CUBITIX2"
    at, out = HDR.size, []
    for _ in range(card):
        enc, aval, anb, _pad, nw, nd = ENT.unpack_from(image, at)
        at += ENT.size
        nbytes = nw * (4 if enc else 8)
        payload = np.frombuffer(image, dtype="<u4" if enc else "<u8", count=nw, offset=at)
        at += (nbytes + 7) & ~7
        rows = np.frombuffer(image, dtype="<i8", count=nd, offset=at)
        at += nd * 8
        out.append((enc, aval, anb, payload, rows))
    assert at + 8 == len(image)
    return n_rows, card, src_col, src_base, out


@pytest.mark.parametrize("n", [200_003, 31 * 64 * 5, 1])
def test_serialize_deserialize_round_trip(cubit, n):
    rng = np.random.default_rng(n)
    # values 0..3 common (dense, incompressible bitvectors), 4..11 rare or clustered (WAH wins)
    key = rng.integers(0, 4, n).astype(np.int64)
    rare = rng.random(n) < 0.002
    key[rare] = rng.integers(4, 10, int(rare.sum()))
    if n > 70_000:
        key[50_000:50_500] = 10          # one run of ones inside zeros
    payload = rng.integers(-10**9, 10**9, n).astype(np.int64)
    t = cubit.CubitTable(n, seg_bits=32768)
    t.upload_column(0, key)
    t.upload_column(1, payload)
    ix = t.create_index(12)
    t.build_index(ix, 0, 0)
    deltas = {}
    if n > 100:
        for v in (1, 5):
            rows = np.unique(rng.integers(0, n, 300)).astype(np.int64)
            t.set_delta(ix, v, rows)
            deltas[v] = rows
    image = t.serialize_index(ix)
    n_rows, card, src_col, src_base, entries = parse(image)
    assert (n_rows, card, src_col, src_base) == (n, 12, 0, 0)
    encodings = set()
    for v, (enc, aval, anb, words, rows) in enumerate(entries):
        bv = t.download_bitvector(ix, v)
        encodings.add(enc)
        wah, oval, onb = oracle.wah_encode(bv, n)
        assert enc == (1 if len(wah) * 4 < len(bv) * 8 else 0), v             # the smaller form is kept
        if enc:
            assert np.array_equal(words, wah) and (aval, anb) == (oval, onb), v   # product writer == oracle writer
            assert len(words) * 4 < len(bv) * 8
        else:
            assert np.array_equal(words, bv)
        assert np.array_equal(rows, deltas.get(v, np.zeros(0, dtype=np.int64)))
    if n > 70_000:
        assert encodings == {0, 1}

    t2 = cubit.CubitTable(n, seg_bits=65536)     # another shard object, another segment size
    t2.upload_column(0, key)
    t2.upload_column(1, payload)
    ix2 = t2.deserialize_index(image)
    for v in range(12):
        assert np.array_equal(t2.download_bitvector(ix2, v), t.download_bitvector(ix, v)), v
    for vals in ([1], [0, 1, 2, 3], [4, 5, 6, 10], [5, 11]):
        with t.query([[(ix, v) for v in vals]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=1) as a, \
                t2.query([[(ix2, v) for v in vals]], flags=cubit.Q_ROWIDS, agg=cubit.AGG_SUM, agg_a=1) as b:
            assert a.count == b.count and a.sum == b.sum
            assert np.array_equal(a.fetch()[0], b.fetch()[0])
    # the reloaded index still knows its source column: an append extends it on the GPU
    extra = np.array([5, 5, 2], dtype=np.int64)
    t2.append_rows({0: extra, 1: np.array([7, 8, 9], dtype=np.int64)})
    with t2.query([[(ix2, 5)]], flags=cubit.Q_ROWIDS) as r:
        ids = r.fetch()[0]
        assert n in ids and n + 1 in ids and n + 2 not in ids
    # damaged images are rejected
    bad = bytearray(image)
    bad[len(bad) // 2] ^= 0x10
    t3 = cubit.CubitTable(n)
    for img in (bytes(bad), image[:-8], image[:40], b"NOTANIDX" + image[8:]):
        with pytest.raises(cubit.CubitError):
            t3.deserialize_index(img)
    t4 = cubit.CubitTable(n + 64)
    with pytest.raises(cubit.CubitError):
        t4.deserialize_index(image)               # image of a table with another row count
    for x in (t, t2, t3, t4):
        x.close()
