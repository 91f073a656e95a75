"""GPU box: the REAL reference DuckDB (baseline/_ref/libduckdb.so, built from /root/reference) driving the REAL
CUDA path through SQL — table functions, optimizer rewrite, DataChunk hand-off, NULL masks, storage route —
and every answer compared with the reference's own vanilla scan of the same table in the same process.
The binaries are produced in the build container by tools/build_ref_bundle.py (they need the reference's
headers) and travel in baseline/_ref/; skipped where that bundle is absent."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "duckdb_sql_gpu_test")),
                                 reason="baseline/_ref bundle not built (tools/build_ref_bundle.py, build container)")]


@pytest.mark.parametrize("devices", [None, "0,0,0", "all"])
def test_sql_through_reference_duckdb_on_the_gpu(tmp_path, devices):
    """devices = None: one GPU, one shard.  "0,0,0": the glue cuts every table into three row-range shards behind one
    handle (cubit_gpu_create_sharded; here all on device 0, which exercises the whole fan-out on a one-GPU box).
    "all": one shard per visible GPU (the same as None on a one-GPU box)."""
    env = dict(os.environ)
    env.pop("CUBIT_GPU_DEVICES", None)
    if devices:
        env["CUBIT_GPU_DEVICES"] = devices
    r = subprocess.run([os.path.join(REF, "duckdb_sql_gpu_test"), "--db", str(tmp_path / "route.db")], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for marker in ("aggregate push-down ok", "multi-index conjunctions ok", "binned indexes ok", "null semantics ok",
                   "dml through the index ok", "storage route ok", "index persistence ok", "duckdb_sql_test ok"):
        assert marker in r.stdout, r.stdout


def test_config1_on_reference_tpch_data_small_scale():
    """SF0.1 here (seconds); the SF1 run with timings is tools → profiles/r1_config1_duckdb.json"""
    r = subprocess.run([os.path.join(REF, "duckdb_config1"), "0.1", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["tpch_sf01"]["answers"]["q_eq_24"]
    cnt, total = out["answer_q1"].split("|")
    assert int(cnt) == g["count"] and int(total.replace(".", "")) == g["sum_price_cents"]
    # TPC-H Q6 as written, answered by the GPU through plain SQL, equals the reference's answer for SF0.1
    # (extension/tpch/dbgen/answers/sf0.1/q06.csv; also recorded in tests/golden/golden.json)
    assert out["tpch_q6_as_written"]["answer"] == "11803420.2534"
