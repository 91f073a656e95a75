"""GPU: short queries (k <= 4 bitvectors, no pending deltas) as two streaming passes — merge + count, then decode
(small_scan_kernels.cu) — or, with row positions, as ONE pass with a decoupled look-back (lookback_scan_kernel.cu;
CUBIT_NO_LOOKBACK=1 keeps the passes), against the oracle, through the C-ABI.  The planner takes this path on large tables only;
CUBIT_TWO_PASS_MIN_ROWS=0 forces it here at sizes the oracle finishes in seconds.  Bit-exact: COUNT, row IDs, the
merged bitvector, probed values (gather, bit-driven and dense probes behind it), aggregates."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_two_pass():
    old = os.environ.get("CUBIT_TWO_PASS_MIN_ROWS")
    os.environ["CUBIT_TWO_PASS_MIN_ROWS"] = "0"
    yield
    if old is None:
        del os.environ["CUBIT_TWO_PASS_MIN_ROWS"]
    else:
        os.environ["CUBIT_TWO_PASS_MIN_ROWS"] = old


@pytest.mark.parametrize("lookback", [True, False])
@pytest.mark.parametrize("seg_bits,n", [(65536, 1_300_017), (32768, 900_001), (131072, 700_003), (65536, 5_000),
                                        (65536, 262_144 * 3), (65536, 262_144 * 3 + 1)])
def test_two_pass_scan_matches_oracle(cubit, force_two_pass, monkeypatch, seg_bits, n, lookback):
    if not lookback:
        monkeypatch.setenv("CUBIT_NO_LOOKBACK", "1")
    rng = np.random.default_rng(n)
    base = seg_bits * 3
    key = rng.integers(0, 12, n).astype(np.int32)
    key[n // 3: n // 3 + min(n // 4, 200_000)] = 11                    # a long stretch without selected rows
    rare = (rng.random(n) < 0.003).astype(np.int32)                    # a sparse second index
    pay = rng.integers(-2**40, 2**40, n).astype(np.int64)
    small = rng.integers(0, 5000, n).astype(np.int64)                  # packs to 13 bits
    bv = oracle.build_index(key, 0, 12)
    bvr = oracle.build_index(rare, 0, 2)
    t = cubit.CubitTable(n, row_base=base, seg_bits=seg_bits)
    t.upload_column(0, pay)
    t.upload_column(1, small)
    t.upload_column(8, key)
    t.upload_column(9, rare)
    ix = t.create_index(12)
    t.build_index(ix, 8, 0)
    ixr = t.create_index(2)
    t.build_index(ixr, 9, 0)
    t.pack_column(1, keep_raw=True)
    cases = [
        ([[(ix, 3)]], [[bv[3]]]),
        ([[(ixr, 1)]], [[bvr[1]]]),                                      # sparse: gather probe over the row IDs
        ([[(ix, 0), (ix, 1), (ix, 2), (ix, 3)]], [[bv[0], bv[1], bv[2], bv[3]]]),
        ([[(ix, 4), (ix, 5)], [(ix, 5), (ix, 6)]], [[bv[4], bv[5]], [bv[5], bv[6]]]),   # AND of two ORs
        ([[(ix, 7), (ix, 8), (ix, 9)], [(ixr, 0)]], [[bv[7], bv[8], bv[9]], [bvr[0]]]),
        ([[(ix, 11)]], [[bv[11]]]),
        ([[(ix, 0), (ix, 1), (ix, 2)]], [[bv[0], bv[1], bv[2]]]),          # k = 3, one OR group
        ([[(ix, 8), (ix, 9)], [(ixr, 0)]], [[bv[8], bv[9]], [bvr[0]]]),     # k = 3, AND of an OR and a single bitvector
        ([[(ix, 1)], [(ixr, 0)]], [[bv[1]], [bvr[0]]]),                  # AND of two single bitvectors (a = 1 AND b = 0)
        ([[(ix, 2)], [(ixr, 1)]], [[bv[2]], [bvr[1]]]),                  # the same, nearly empty
        ([[(ix, 3)], [(ix, 4)]], [[bv[3]], [bv[4]]]),                    # disjoint: empty result
    ]
    for groups, bvs in cases:
        q = oracle.merge(bvs)
        want = oracle.decode(q, base)
        wp, ws = oracle.probe(want, pay, base), oracle.probe(want, small, base)
        k = sum(len(g) for g in groups)
        # the planner's rule: k <= 3, with row positions (look-back kernel) and for count / bitvector / aggregate only
        with_pos = (cubit.SCAN_LOOKBACK if lookback else cubit.SCAN_TWO_PASS) if k <= 3 else cubit.SCAN_RING
        no_pos = cubit.SCAN_TWO_PASS if k <= 3 else cubit.SCAN_RING
        with t.query(groups, flags=cubit.Q_ROWIDS) as r:
            assert r.info.scan_path == with_pos and r.count == len(want)
            assert np.array_equal(r.fetch()[0], want)
        with t.query(groups, flags=0) as r:                                # COUNT only: pass A alone
            assert r.info.scan_path == no_pos and r.count == len(want) and r.info.n_launches == 1
        with t.query(groups, flags=cubit.Q_BITVECTOR | cubit.Q_ROWIDS) as r:
            assert np.array_equal(r.bitvector(), q) and np.array_equal(r.fetch()[0], want)
        with t.query(groups, flags=cubit.Q_ROWIDS | cubit.Q_VALUES, cols=[0, 1], agg=cubit.AGG_SUM_PROD, agg_a=1, agg_b=1) as r:
            ids, (gp, gs) = r.fetch()
            assert r.info.scan_path == with_pos
            assert np.array_equal(ids, want) and np.array_equal(gp, wp) and np.array_equal(gs, ws)
            assert r.sum == oracle.sum_prod_i64(ws, ws)[0]
        with t.query(groups, flags=cubit.Q_VALUES, cols=[1], agg=cubit.AGG_SUM, agg_a=1) as r:   # positions, no row IDs
            assert np.array_equal(r.fetch(rowids=False)[1][0], ws) and r.sum == oracle.sum_i64(ws)
        if k > 1:                                                          # (k = 1, aggregate only: probed in place, no scan)
            with t.query(groups, flags=0, agg=cubit.AGG_SUM, agg_a=0) as r:
                # (a sparse selection is probed by the gather over its row IDs: the scan then needs positions after all)
                assert r.info.scan_path in (no_pos, with_pos) and r.count == len(want) and r.sum == oracle.sum_i64(wp)
    # five bitvectors, or pending deltas: the ring kernel keeps the job
    with t.query([[(ix, v) for v in range(5)]], flags=cubit.Q_ROWIDS) as r:
        assert r.info.scan_path == cubit.SCAN_RING
        assert np.array_equal(r.fetch()[0], oracle.decode(oracle.merge([[bv[v] for v in range(5)]]), base))
    flips = rng.integers(0, n, 100)
    t.set_merge_threshold(ix, 0)                                          # keep them pending whatever the table size
    t.add_delta(ix, 3, flips)
    with t.query([[(ix, 3)]], flags=cubit.Q_ROWIDS) as r:
        assert r.info.scan_path == cubit.SCAN_RING and r.info.delta_entries > 0
        d = oracle.delta_from_rows(flips, n)
        assert np.array_equal(r.fetch()[0], oracle.decode(oracle.merge([[bv[3]]], [[d]]), base))
    t.close()


def test_small_tables_keep_the_ring_kernel(cubit):
    t = cubit.CubitTable(100_000)
    t.synth_column(1, 2, seed=5, card=4, hot_lo=0)
    ix = t.create_index(4)
    t.build_index(ix, 1, 0)
    with t.query([[(ix, 1)]], flags=cubit.Q_ROWIDS) as r:
        assert r.info.scan_path == cubit.SCAN_RING
    t.close()
