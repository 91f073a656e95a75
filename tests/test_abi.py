"""CPU: the C-ABI library loads and exports every symbol include/cubit_gpu.h declares.
No compute is attempted here (there is no GPU in the build container)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cubit_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cubit_gpu_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(cubit):
    assert header_functions() == sorted(cubit.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(cubit):
    lib = ctypes.CDLL(cubit.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.cubit_gpu_abi_version() == cubit.ABI_VERSION


def test_struct_layout_matches_header(cubit, tmp_path):
    """sizeof/offsetof as the C compiler sees include/cubit_gpu.h == the ctypes mirror"""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "cubit_gpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cubit_result_info), sizeof(cubit_query), sizeof(cubit_pred_group),
         sizeof(cubit_bv_ref), offsetof(cubit_result_info, ms_scan), offsetof(cubit_result_info, d_rowids),
         offsetof(cubit_query, agg_kind));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(cubit.ResultInfo), ctypes.sizeof(cubit.Query), ctypes.sizeof(cubit.PredGroup),
            ctypes.sizeof(cubit.BvRef), cubit.ResultInfo.ms_scan.offset, cubit.ResultInfo.d_rowids.offset,
            cubit.Query.agg_kind.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device(cubit):
    """on a box without a B200 the product path must fail loudly, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert cubit.device_count() == 0
    with pytest.raises(cubit.CubitError) as e:
        cubit.CubitTable(1000)
    assert e.value.code == cubit.ENODEVICE


def test_argument_validation(cubit):
    L = cubit.load_library()
    h = ctypes.c_void_p()
    assert L.cubit_gpu_create(0, 1000, 0, 12345, ctypes.byref(h)) == cubit.EINVAL
    assert b"seg_bits" in L.cubit_gpu_last_error()
    assert L.cubit_gpu_create(0, 0, 0, 65536, ctypes.byref(h)) == cubit.EINVAL
    assert L.cubit_gpu_create(0, 10, 7, 65536, ctypes.byref(h)) == cubit.EINVAL
    assert L.cubit_gpu_create(0, 10, 0, 65536, None) == cubit.EINVAL
    # NULL handles / out-pointers are argument errors on every entry point that can be reached without a device
    n = ctypes.c_uint64(0)
    ix = ctypes.c_int32(0)
    img = ctypes.c_void_p()
    assert L.cubit_gpu_upload_column_validity(None, 0, None, 0) == cubit.EINVAL
    assert L.cubit_gpu_fetch_validity(None, 0, 0, 0, None, None) == cubit.EINVAL
    assert L.cubit_gpu_index_serialize(None, 0, ctypes.byref(img), ctypes.byref(n)) == cubit.EINVAL
    assert L.cubit_gpu_index_deserialize(None, None, 0, ctypes.byref(ix)) == cubit.EINVAL
    assert L.cubit_gpu_alloc_host(16, None) == cubit.EINVAL
    assert L.cubit_gpu_free_host(None) == cubit.OK
    L.cubit_gpu_free_image(None)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under duckdb-cubit_b200/ may reference it"""
    pkg = os.path.join(ROOT, "duckdb-cubit_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "cubit_oracle" not in txt.replace(
                    "oracle/cubit_oracle.c", ""), os.path.join(dirpath, f)
