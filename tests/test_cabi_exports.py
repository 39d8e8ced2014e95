"""CPU: the C-ABI library loads and exports every symbol include/lira_b200.h declares; the Python
binding table covers the same set; compute calls fail loudly without a GPU (no fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lira_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lira_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import lira_ann_search_b200 as L
    from lira_ann_search_b200 import _cabi
    lib = _cabi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lira_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == syms
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lira_[a-z0-9_]+)", out))
    assert exported == set(syms)
    assert lib.lira_version() >= 100


def test_no_link_dependency_on_driver_or_torch():
    from lira_ann_search_b200 import _cabi
    out = subprocess.run(["ldd", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "torch" not in out and "libcudart" not in out


def test_compute_fails_loudly_without_gpu():
    import lira_ann_search_b200 as L
    if L._cabi.lib().lira_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(L.LiraError):
        L.LiraIndex.from_cluster_ids(np.zeros((4, 4), np.float32), [[0, 1], [2, 3]])
    with pytest.raises(L.LiraError):
        L.knn(np.zeros((4, 4), np.float32), np.zeros((1, 4), np.float32), 2)
    # straight through the C ABI as well
    lib = L._cabi.lib()
    h = ctypes.c_void_p()
    off = np.array([0, 2, 4], np.int64)
    ids = np.arange(4, dtype=np.int32)
    x = np.zeros((4, 4), np.float32)
    rc = lib.lira_index_create(x.ctypes.data_as(L._cabi.c_f32p), 4, 4, off.ctypes.data_as(L._cabi.c_i64p),
                               ids.ctypes.data_as(L._cabi.c_i32p), 2, 0, 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.lira_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lira-ann-search_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("oracle/", "") or "import oracle" not in src, f
                assert "import oracle" not in src and "liblira_oracle" not in src, f
