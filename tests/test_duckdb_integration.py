"""CPU, build container only: the DuckDB-side glue (integration/duckdb_cubit_extension.cpp) compiled against
the REAL reference headers, linked with the reference's libduckdb.so and driven through SQL.
  * with the real libcubit_gpu.so (no GPU here): the C-ABI error must surface as a DuckDB exception
  * with the oracle-backed mock of the C-ABI (tests/mock, test infrastructure): cubit_scan / cubit_agg must
    return exactly what the vanilla scan returns for the same predicate.
Skipped where /root/reference or its build is absent (e.g. on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/src/include"
REF_LIB_DIR = os.environ.get("CUBIT_REF_LIBDIR", "/tmp/duckdb_build/src")
PKG = os.path.join(ROOT, "duckdb-cubit_b200")

pytestmark = pytest.mark.skipif(not (os.path.isdir(REF_INC) and os.path.exists(os.path.join(REF_LIB_DIR, "libduckdb.so"))),
                                reason="reference DuckDB headers/library not present")


def _build(tmp, name, abi_lib_dir, abi_lib):
    exe = os.path.join(tmp, name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", REF_INC, "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "duckdb_cubit_extension.cpp"),
                           os.path.join(ROOT, "tests", "cpp", "duckdb_sql_test.cpp"), "-o", exe,
                           "-L", REF_LIB_DIR, "-lduckdb", "-L", abi_lib_dir, "-l" + abi_lib,
                           "-Wl,-rpath," + REF_LIB_DIR, "-Wl,-rpath," + abi_lib_dir, "-lpthread", "-ldl"])
    return exe


def test_glue_with_real_library_reports_missing_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = _build(str(tmp_path), "sql_real", PKG, "cubit_gpu")
    r = subprocess.run([exe, "--expect-no-device"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr


def test_glue_sql_equals_vanilla_scan_with_oracle_mock(tmp_path):
    import oracle
    oracle.build()
    mock_dir = str(tmp_path)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "mock", "cubit_gpu_mock.c"), "-o",
                           os.path.join(mock_dir, "libcubit_gpu_mock.so"), "-L", os.path.join(ROOT, "oracle"),
                           "-lcubit_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    exe = _build(mock_dir, "sql_mock", mock_dir, "cubit_gpu_mock")
    r = subprocess.run([exe, "--db", os.path.join(mock_dir, "storage_route.db")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "duckdb_sql_test ok" in r.stdout
    # columns of a checkpointed, file-backed table reached the C-ABI as the reference's compressed segments
    assert "storage route ok" in r.stdout
    # UPDATE / DELETE / INSERT reached the C-ABI through the registered CUBIT index type (BoundIndex::Append / Delete)
    assert "dml through the index ok" in r.stdout
    # ... and the index came back from its checkpointed image after the database file was re-opened
    assert "index persistence ok" in r.stdout
