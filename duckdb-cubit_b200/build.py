"""Build script for the CUDA library (sm_100a only) and the C++ host layer.

`python duckdb-cubit_b200/build.py` compiles, in-tree:
  duckdb-cubit_b200/libcubit_gpu.so    kernels + C-ABI (include/cubit_gpu.h)
  duckdb-cubit_b200/libcubit_host.so   C++ host mirror of the reference operator API (host/)
The .so files are git-ignored but travel to the GPU box with the snapshot.
nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "build")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_gpu(verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "scan_common.cuh"), os.path.join(CSRC, "col_ref.cuh"),
            os.path.join(CSRC, "table.h"), os.path.join(ROOT, "include", "cubit_gpu.h"),
            os.path.join(ROOT, "include", "cubit_gpu_wire.h")]
    units = ["scan_kernel.cu", "aux_kernels.cu", "column_decode.cu", "wah_decode.cu", "delta_kernels.cu",
             "container_kernels.cu", "probe_dense_kernel.cu", "small_scan_kernels.cu", "lookback_scan_kernel.cu", "cubit_gpu.cu", "cubit_columns.cu", "cubit_delta.cu", "cubit_persist.cu",
             "cubit_query.cu", "cubit_sharded.cu", "cubit_wire.cu"]
    objs = []
    jobs = []
    for u in units:
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJ, u.replace(".cu", ".o"))
        objs.append(obj)
        if _stale(obj, [src] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(lambda c: _run(c, verbose), jobs))
    lib = os.path.join(HERE, "libcubit_gpu.so")
    if jobs or _stale(lib, objs):
        _run([NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
              "-o", lib] + objs + ["-lpthread", "-ldl", "-lrt"], verbose)
    return lib, outs


def build_host(verbose=False):
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp")) if os.path.isdir(HOST) else []
    if not srcs:
        return None
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")]
    hdrs.append(os.path.join(ROOT, "include", "cubit_gpu.h"))
    hdrs.append(os.path.join(ROOT, "include", "cubit_gpu_wire.h"))
    lib = os.path.join(HERE, "libcubit_host.so")
    if _stale(lib, srcs + hdrs):
        _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
              "-o", lib] + srcs + ["-L", HERE, "-lcubit_gpu", "-Wl,-rpath,$ORIGIN"], verbose)
    return lib


def build_all(verbose=False):
    lib, _ = build_gpu(verbose)
    host = build_host(verbose)
    return lib, host


if __name__ == "__main__":
    v = "-q" not in sys.argv
    lib, outs = build_gpu(verbose=v, ptxas_info="--ptxas" in sys.argv)
    if "--ptxas" in sys.argv:
        for o in outs:
            print(o)
    print("built", lib)
    h = build_host(verbose=v)
    if h:
        print("built", h)
