"""WAH-compressed value bitvectors (FastBit ibis::bitvector form, what the upstream CUBIT library stores).

FastBit is not in /root/reference: the format is restated from its published description and the
one known-answer vector that description carries (Wu/Otoo/Shoshani, TODS 2006, Fig. 1) is checked here.
GPU: cubit_gpu_upload_bitvector_wah expands on the device; bit-exact against the oracle's decoder.
"""
import numpy as np
import pytest

import oracle


def bits_to_words(bits):
    return np.packbits(np.pad(bits, (0, -len(bits) % 64)), bitorder="little").view(np.uint64)


def make_bitmap(rng, n, kind):
    if kind == "sparse":
        b = rng.random(n) < 0.001
    elif kind == "dense":
        b = rng.random(n) < 0.5
    elif kind == "runs":        # long 0-fills and long 1-fills with literal words at the edges
        b = np.zeros(n, dtype=bool)
        pos = 0
        val = False
        while pos < n:
            ln = int(rng.integers(1, 200_000 if rng.random() < 0.3 else 50))
            b[pos:pos + ln] = val
            val = not val
            pos += ln
    elif kind == "ones":
        b = np.ones(n, dtype=bool)
    else:
        b = np.zeros(n, dtype=bool)
    return b


def test_published_example_vector():
    # 128 bits: 1, 20 zeros, 3 ones, 79 zeros, 25 ones  →  40000380 80000002 001FFFFF, active 0000000F (4 bits)
    bits = np.array([1] + [0] * 20 + [1] * 3 + [0] * 79 + [1] * 25, dtype=bool)
    wah, av, an = oracle.wah_encode(bits_to_words(bits), 128)
    assert [int(x) for x in wah] == [0x40000380, 0x80000002, 0x001FFFFF] and (av, an) == (0xF, 4)
    back, n = oracle.wah_decode(np.array([0x40000380, 0x80000002, 0x001FFFFF], dtype=np.uint32), 0xF, 4, 2)
    assert n == 128 and np.array_equal(back, bits_to_words(bits))


def golden_vectors():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wah_vectors.json")
    for v in json.load(open(path))["vectors"]:
        bits = np.concatenate([np.full(n, b, dtype=bool) for b, n in v["runs"]])
        yield (v["name"], bits, np.array([int(w, 16) for w in v["wah"]], dtype=np.uint32), int(v["active_val"], 16),
               v["active_nbits"])


def test_golden_vectors_from_the_independent_encoder():
    """tests/golden/wah_vectors.json (40 vectors of tests/golden/make_wah_golden.py, a pure-Python encoder that shares
    no code with the oracle): the oracle writes exactly those words and decodes them back to the same bits"""
    seen = 0
    for name, bits, wah, av, an in golden_vectors():
        w = bits_to_words(bits)
        got, gav, gan = oracle.wah_encode(w, len(bits))
        assert [int(x) for x in got] == [int(x) for x in wah] and (gav, gan) == (av, an), name
        back, n = oracle.wah_decode(wah, av, an, len(w))
        assert n == len(bits) and np.array_equal(back, w), name
        seen += 1
    assert seen == 40


def test_reader_accepts_fills_of_one_group():
    """a fill word spanning ONE group is never written (FastBit appends a literal) but is legal input"""
    bits = np.array([0] * 31 + [1] * 31 + [1, 0, 1], dtype=bool)
    back, n = oracle.wah_decode(np.array([0x80000001, 0xC0000001], dtype=np.uint32), 0b101, 3, 2)
    assert n == 65 and np.array_equal(back, bits_to_words(bits))
    wah, _, _ = oracle.wah_encode(bits_to_words(bits), 65)
    assert [int(x) for x in wah] == [0x00000000, 0x7FFFFFFF]


@pytest.mark.parametrize("kind", ["sparse", "dense", "runs", "ones", "zeros"])
def test_oracle_round_trip(kind):
    rng = np.random.default_rng(17)
    for n in (1, 30, 31, 32, 62, 63, 64, 1000, 65536 + 17, 1_000_003):
        b = make_bitmap(rng, n, kind)
        w = bits_to_words(b)
        wah, av, an = oracle.wah_encode(w, n)
        assert an == n % 31
        back, nb = oracle.wah_decode(wah, av, an, len(w))
        assert nb == n and np.array_equal(back, w), (kind, n)
        if kind in ("ones", "zeros") and n >= 31:
            assert len(wah) == 1                      # one fill word (a literal when it is a single group)
    with pytest.raises(ValueError):
        oracle.wah_decode(np.array([0x80000000], dtype=np.uint32), 0, 0, 4)       # zero-length fill
    with pytest.raises(ValueError):
        oracle.wah_decode(np.array([0x80000010], dtype=np.uint32), 0, 0, 4)       # longer than the output
    with pytest.raises(ValueError):
        oracle.wah_decode(np.zeros(0, dtype=np.uint32), 0, 31, 4)                 # active word too long


@pytest.mark.gpu
@pytest.mark.parametrize("seg_bits", [65536, 32768])
def test_gpu_expands_wah_bitvectors(cubit, seg_bits):
    rng = np.random.default_rng(23)
    n = 3_000_017
    with cubit.CubitTable(n, row_base=seg_bits, seg_bits=seg_bits) as t:
        ix = t.create_index(8)
        maps = []
        for v, kind in enumerate(["sparse", "dense", "runs", "ones", "zeros", "runs", "sparse", "dense"]):
            b = make_bitmap(rng, n if v != 6 else n - 100_003, kind)   # v=6: shorter than the table → zero padded
            wah, av, an = oracle.wah_encode(bits_to_words(b), len(b))
            t.upload_bitvector_wah(ix, v, wah, av, an)
            want, _ = oracle.wah_decode(wah, av, an, t.n_words)
            got = t.download_bitvector(ix, v)
            assert np.array_equal(got, want), (v, kind)
            full = np.zeros(n, dtype=bool)
            full[:len(b)] = b
            assert np.array_equal(got, bits_to_words(full))
            assert t.bitvector_count(ix, v) == int(full.sum())
            maps.append(want)
        # re-upload over a dense bitvector: the destination is cleared first
        wah, av, an = oracle.wah_encode(maps[0], n)
        t.upload_bitvector_wah(ix, 1, wah, av, an)
        assert np.array_equal(t.download_bitvector(ix, 1), maps[0])
        maps[1] = maps[0]
        # the scan over WAH-uploaded bitvectors: (B0 | B2 | B5) & (B3) & (B7 | B6)
        groups = [[0, 2, 5], [3], [7, 6]]
        want = oracle.decode(oracle.merge([[maps[v] for v in g] for g in groups]), seg_bits)
        with t.query([[(ix, v) for v in g] for g in groups], flags=cubit.Q_ROWIDS) as r:
            assert np.array_equal(r.fetch()[0], want)
        # only an active word (fewer than 31 bits described)
        t.upload_bitvector_wah(ix, 4, np.zeros(0, dtype=np.uint32), 0b101, 3)
        got = t.download_bitvector(ix, 4)
        assert got[0] == 0b101 and not got[1:].any()
        for bad in ((np.array([0x80000000], dtype=np.uint32), 0, 0),                     # zero-length fill
                    (np.array([0xC0000000 | (n // 31 + 1)], dtype=np.uint32), 0, 0),      # longer than the table
                    (np.zeros(0, dtype=np.uint32), 0, 31),                                # active too long
                    (np.zeros(0, dtype=np.uint32), 0b1000, 3)):                           # active value wider than nbits
            with pytest.raises(cubit.CubitError):
                t.upload_bitvector_wah(ix, 0, *bad)
        assert np.array_equal(t.download_bitvector(ix, 0), maps[0])   # untouched by the rejected uploads


@pytest.mark.gpu
def test_gpu_expands_the_golden_vectors(cubit):
    """every vector of tests/golden/wah_vectors.json, uploaded in its WAH form, expands on the GPU to the bits the
    independent encoder started from (shorter than the table → zero padded)"""
    n = 31 * 70001 + 64
    with cubit.CubitTable(n, seg_bits=32768) as t:
        ix = t.create_index(1)
        for name, bits, wah, av, an in golden_vectors():
            t.upload_bitvector_wah(ix, 0, wah, av, an)
            full = np.zeros(n, dtype=bool)
            full[:len(bits)] = bits
            assert np.array_equal(t.download_bitvector(ix, 0), bits_to_words(full)), name
            assert t.bitvector_count(ix, 0) == int(bits.sum()), name
