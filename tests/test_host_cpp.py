"""The C++ host layer (duckdb-cubit_b200/host): CubitIndex + the cubit_scan table function driven the way
PhysicalTableScan::GetData drives a table function.  The checks themselves live in tests/cpp/host_scan_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "duckdb-cubit_b200")


@pytest.fixture(scope="module")
def host_test_binary(tmp_path_factory):
    assert os.path.exists(os.path.join(PKG, "libcubit_host.so")), "run python duckdb-cubit_b200/build.py"
    exe = str(tmp_path_factory.mktemp("hostcpp") / "host_scan_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(PKG, "host"), os.path.join(ROOT, "tests", "cpp", "host_scan_test.cpp"), "-o",
                           exe, "-L", PKG, "-lcubit_host", "-lcubit_gpu", "-Wl,-rpath," + PKG, "-lpthread"])
    return exe


def test_host_layer_builds_and_has_no_cpu_fallback(host_test_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([host_test_binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr  # a C++ exception, not a silent CPU path


@pytest.mark.gpu
def test_host_scan_against_brute_force(host_test_binary):
    r = subprocess.run([host_test_binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "host_scan_test ok" in r.stdout
    print(r.stdout)
