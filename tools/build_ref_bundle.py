#!/usr/bin/env python
"""Build container only: put what the GPU box needs to run the REAL reference next to the GPU path into
baseline/_ref/ (git-ignored, travels with gpurun):
  libduckdb.so          the unmodified reference, as built by the driver from /root/reference (Release, tpch linked)
  duckdb_sql_gpu_test   integration/duckdb_cubit_extension.cpp + tests/cpp/duckdb_sql_test.cpp linked with the
                        reference's libduckdb.so and the REAL libcubit_gpu.so (on the GPU box: SQL through the
                        optimizer rewrite → CUDA kernels → DataChunks, compared with the vanilla scan)
  duckdb_config1        tests/cpp/duckdb_config1.cpp, config 1 on the reference's own TPC-H data
  duckdb_cfg2_baseline  tests/cpp/duckdb_cfg2_baseline.cpp: the reference's CPU path on a sample of the bench workload
Nothing here is product code and no reference SOURCE is copied.  Usage: python tools/build_ref_bundle.py"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/src/include"
REF_LIB = os.environ.get("CUBIT_REF_LIBDIR", "/tmp/duckdb_build/src")
OUT = os.path.join(ROOT, "baseline", "_ref")


def stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def main():
    lib = os.path.join(REF_LIB, "libduckdb.so")
    if not (os.path.isdir(REF_INC) and os.path.exists(lib)):
        print("reference headers / libduckdb.so not present: nothing built")
        return 1
    os.makedirs(OUT, exist_ok=True)
    if stale(os.path.join(OUT, "libduckdb.so"), [lib]):
        shutil.copy2(lib, os.path.join(OUT, "libduckdb.so"))
    glue = os.path.join(ROOT, "integration", "duckdb_cubit_extension.cpp")
    hdr = os.path.join(ROOT, "include", "cubit_gpu.h")
    for exe, src in (("duckdb_sql_gpu_test", "duckdb_sql_test.cpp"), ("duckdb_config1", "duckdb_config1.cpp")):
        src = os.path.join(ROOT, "tests", "cpp", src)
        dst = os.path.join(OUT, exe)
        if stale(dst, [glue, hdr, os.path.join(ROOT, "include", "cubit_gpu_wire.h"), src]):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", REF_INC, "-I", os.path.join(ROOT, "include"), glue, src,
                                   "-o", dst, "-L", OUT, "-lduckdb", "-L", os.path.join(ROOT, "duckdb-cubit_b200"),
                                   "-lcubit_gpu", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../duckdb-cubit_b200",
                                   "-lpthread", "-ldl"])
            print("built", dst)
    # the reference's dbgen as a slice generator (config 3 on real TPC-H data): libduckdb.so only
    src = os.path.join(ROOT, "tests", "cpp", "tpch_slices.cpp")
    dst = os.path.join(OUT, "tpch_slices")
    if stale(dst, [src]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", REF_INC, src, "-o", dst, "-L", OUT, "-lduckdb",
                               "-Wl,-rpath,$ORIGIN", "-lpthread", "-ldl"])
        print("built", dst)
    # the reference's CPU path on the bench workload (bench.py cpu_baseline / --impl reference): no glue, no GPU
    # library — libduckdb.so + the oracle's table generator
    src = os.path.join(ROOT, "tests", "cpp", "duckdb_cfg2_baseline.cpp")
    dst = os.path.join(OUT, "duckdb_cfg2_baseline")
    sys.path.insert(0, ROOT)
    import oracle
    oracle.build()
    if stale(dst, [src]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", REF_INC, src, "-o", dst, "-L", OUT, "-lduckdb",
                               "-L", os.path.join(ROOT, "oracle"), "-lcubit_oracle", "-Wl,-rpath,$ORIGIN",
                               "-Wl,-rpath,$ORIGIN/../../oracle", "-lpthread", "-ldl"])
        print("built", dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
