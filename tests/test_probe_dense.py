"""GPU: the dense probe (probe_dense_kernel.cu — pack blocks of bit-packed columns streamed through per-warp
shared-memory stages) against the oracle, through the C-ABI.  Bit-exact: row IDs, values, integer aggregates.

Covers every block width 0..32 with positive and negative FOR bases, widths above 32 (the planner must fall back to
the gather probe), a raw column riding along, two columns (SUM(a*b)), aggregate-only queries (no positions),
selections with empty pack blocks, empty spans and empty segments (copies skipped), all three segment sizes, and a
table whose last pack block and last segment are ragged.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def make_table(cubit, seg_bits, n=1_300_017, seed=5):
    rng = np.random.default_rng(seed)
    blk = np.arange(n) // 1024
    key = rng.integers(0, 11, n).astype(np.int32)
    # holes: whole pack blocks, whole spans and whole segments in which only value 11 occurs
    key[40 * 1024:43 * 1024] = 11
    key[8192 * 9:8192 * 11] = 11
    key[3 * seg_bits:5 * seg_bits] = 11
    key[-700:] = 11
    # a: every width 0..32 in turn, bases of both signs
    width = (blk % 33).astype(np.int64)
    lo = rng.integers(-2**45, 2**45, n // 1024 + 1)[blk]
    a = lo + (rng.random(n) * (1 << width)).astype(np.int64)
    a[blk % 33 == 32] = lo[blk % 33 == 32] + rng.integers(0, 2**32, int((blk % 33 == 32).sum()), dtype=np.int64)
    b = rng.integers(0, 1000, n).astype(np.int64)            # 10 bits
    b[: 5 * 1024] = 7                                         # constant blocks
    c = rng.integers(-2**38, 2**38, n).astype(np.int64)       # 40 bits: not eligible for the dense probe
    d = rng.integers(-2**60, 2**60, n).astype(np.int64)       # stays raw
    base = seg_bits * 2
    t = cubit.CubitTable(n, row_base=base, seg_bits=seg_bits)
    for cid, col in enumerate((a, b, c, d)):
        t.upload_column(cid, col)
    t.upload_column(9, key)
    ix = t.create_index(12)
    t.build_index(ix, 9, 0)
    for cid in (0, 1, 2):
        t.pack_column(cid)
    return t, ix, base, key, (a, b, c, d)


@pytest.mark.parametrize("seg_bits", [32768, 65536, 131072])
def test_dense_probe_matches_oracle(cubit, seg_bits):
    t, ix, base, key, (a, b, c, d) = make_table(cubit, seg_bits)
    bv = oracle.build_index(key, 0, 12)
    for vals in ([3], [0, 1, 2, 3, 4, 5, 6], list(range(11)), list(range(12)), [11]):
        want = oracle.decode(oracle.merge([[bv[v] for v in vals]]), base)
        wa, wb, wc, wd = (oracle.probe(want, col, base) for col in (a, b, c, d))
        groups = [[(ix, v) for v in vals]]
        dense = cubit.PROBE_DENSE
        # one packed column, values + SUM
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0) as r:
            ids, (ga,) = r.fetch()
            assert r.info.probe_path == dense
            assert np.array_equal(ids, want) and np.array_equal(ga, wa) and r.sum == oracle.sum_i64(wa)
        # values without row IDs asked for (positions are still the scan kernel's)
        with t.query(groups, flags=cubit.Q_VALUES, cols=[1]) as r:
            _, (gb,) = r.fetch(rowids=False)
            assert r.info.probe_path == dense and np.array_equal(gb, wb)
        # two packed columns, SUM(a*b) with the overflow check
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[1, 0], agg=cubit.AGG_SUM_PROD, agg_a=1,
                     agg_b=1) as r:
            ids, (gb, ga) = r.fetch()
            sp, ovf = oracle.sum_prod_i64(wb, wb)
            assert r.info.probe_path == dense and not ovf
            assert np.array_equal(ids, want) and np.array_equal(ga, wa) and np.array_equal(gb, wb) and r.sum == sp
        # aggregate only: no positions, no values written
        # (a single clean bitvector is probed in place by the bit-driven probe: no scan launch, no copy of Q)
        agg_path = dense if len(vals) > 1 else cubit.PROBE_BITS
        with t.query(groups, flags=0, agg=cubit.AGG_SUM, agg_a=0) as r:
            assert r.info.probe_path == agg_path and r.count == len(want) and r.sum == oracle.sum_i64(wa)
        with t.query(groups, flags=0, agg=cubit.AGG_SUM_PROD, agg_a=1, agg_b=1) as r:
            assert r.info.probe_path == agg_path and r.sum == oracle.sum_prod_i64(wb, wb)[0]
        # a raw column next to a packed one
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[3, 1], agg=cubit.AGG_SUM, agg_a=3) as r:
            ids, (gd, gb) = r.fetch()
            assert r.info.probe_path == dense
            assert np.array_equal(ids, want) and np.array_equal(gd, wd) and np.array_equal(gb, wb)
            assert r.sum == oracle.sum_i64(wd)
        # 40-bit blocks: the planner keeps the gather probe
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[2], agg=cubit.AGG_SUM, agg_a=2) as r:
            ids, (gc,) = r.fetch()
            assert r.info.probe_path in (cubit.PROBE_BITS, cubit.PROBE_GATHER)
            assert np.array_equal(ids, want) and np.array_equal(gc, wc) and r.sum == oracle.sum_i64(wc)
    t.close()


def test_dense_probe_overflow_and_sparse_fallback(cubit):
    """SUM(a*b) overflow is reported from the dense path too; below the density threshold the gather probes run"""
    t, ix, base, key, (a, b, c, d) = make_table(cubit, 65536, n=400_000, seed=9)
    big = np.full(400_000, 2**31 - 5, dtype=np.int64)
    big[::1000] += 3
    t.upload_column(5, big * 2**10)
    t.pack_column(5)
    with pytest.raises(cubit.CubitError, match="Overflow"):
        t.query([[(ix, v) for v in range(11)]], flags=0, agg=cubit.AGG_SUM_PROD, agg_a=5, agg_b=5)
    # a selection of < 1/28 of the rows: not the dense probe
    rare = np.zeros(400_000, dtype=np.int32)
    rare[::97] = 1
    t.upload_column(6, rare)
    ix2 = t.create_index(2)
    t.build_index(ix2, 6, 0)
    with t.query([[(ix2, 1)]], flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0], agg=cubit.AGG_SUM, agg_a=0) as r:
        ids, (ga,) = r.fetch()
        assert r.info.probe_path != cubit.PROBE_DENSE
        assert np.array_equal(ids, np.arange(0, 400_000, 97) + base) and np.array_equal(ga, a[::97])
    t.close()


def test_append_drops_the_packed_form_and_queries_stay_exact(cubit):
    """a column resident in both forms loses its packed form when rows are appended (it no longer covers the table);
    keep_raw = 2 packs only what the dense probe can use; a packed-only column refuses the append"""
    n, extra = 300_000, 70_001
    rng = np.random.default_rng(3)
    key = rng.integers(0, 4, n + extra).astype(np.int32)
    a = rng.integers(-5000, 5000, n + extra).astype(np.int64)          # 14 bits
    wide = rng.integers(-2**40, 2**40, n + extra).astype(np.int64)     # 41 bits: keep_raw = 2 leaves it raw
    t = cubit.CubitTable(n, seg_bits=32768)
    t.upload_column(0, a[:n])
    t.upload_column(1, wide[:n])
    t.upload_column(9, key[:n])
    ix = t.create_index(4)
    t.build_index(ix, 9, 0)
    assert t.pack_column(0, keep_raw=2) > 0
    assert t.pack_column(1, keep_raw=2) == 0               # left raw: a block needs more than 32 bits
    groups = [[(ix, 1), (ix, 2)]]
    with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1], agg=cubit.AGG_SUM, agg_a=0) as r:
        ids, (ga, gw) = r.fetch()
        sel = np.flatnonzero((key[:n] == 1) | (key[:n] == 2))
        assert r.info.probe_path == cubit.PROBE_DENSE            # column 0 packed, column 1 rides along raw
        assert np.array_equal(ids, sel) and np.array_equal(ga, a[sel]) and np.array_equal(gw, wide[sel])
    t.append_rows({0: a[n:], 1: wide[n:], 9: key[n:]})
    with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1], agg=cubit.AGG_SUM, agg_a=0) as r:
        ids, (ga, gw) = r.fetch()
        sel = np.flatnonzero((key == 1) | (key == 2))
        assert r.info.probe_path != cubit.PROBE_DENSE            # the packed form is gone
        assert np.array_equal(ids, sel) and np.array_equal(ga, a[sel]) and np.array_equal(gw, wide[sel])
        assert r.sum == int(a[sel].sum())
    t.pack_column(0, keep_raw=True)                              # ... and comes back on request
    with t.query(groups, flags=cubit.Q_VALUES, cols=[0]) as r:
        assert r.info.probe_path == cubit.PROBE_DENSE and np.array_equal(r.fetch(rowids=False)[1][0], a[sel])
    t.pack_column(1, keep_raw=False)                             # packed only: appends are refused
    with pytest.raises(cubit.CubitError):
        t.append_rows({0: a[:1], 1: wide[:1], 9: key[:1]})
    t.close()
