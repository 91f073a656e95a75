import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _gpu_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def cubit():
    """the product package (directory name has a hyphen)"""
    return importlib.import_module("duckdb-cubit_b200")


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    ids = np.load(os.path.join(ROOT, "tests", "golden", "golden_ids.npz"))
    return g, ids


@pytest.fixture(scope="session")
def lineitem():
    """TPC-H SF0.01 lineitem columns exported from the reference DuckDB (tests/golden/make_golden.py)"""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "tpch_sf001.npz"))
    return {"quantity": z["quantity"].astype(np.int64), "price": z["price"].astype(np.int64),
            "discount": z["discount"].astype(np.int64), "shipdate": z["shipdate"].astype(np.int32),
            "month": z["month"].astype(np.int64)}
