"""CPU tier: libcellseg_sm100.so builds here (nvcc cross-compiles sm_100a), loads without a GPU and exports
every entry point include/cellseg_b200.h declares; the ctypes table stays in step with the header; argument
validation answers without touching the device; the product fails loudly when no GPU / no library is there."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cellseg_b200 import build, capi
    path = build.build()
    return capi.bind(ctypes.CDLL(path))


def _declared():
    text = open(os.path.join(ROOT, "include", "cellseg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CS_API\s+[\w\s\*]+?\b(cs_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from cellseg_b200 import capi
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(capi.SIGNATURES) == names, "ctypes table and header disagree"


def test_sass_is_sm100a_only():
    import subprocess
    from cellseg_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_device(lib):
    assert lib.cs_abi_version() == 1
    assert lib.cs_postprocess_workspace_bytes(1, 1024, 1024) > 0
    assert lib.cs_postprocess_workspace_bytes(0, 10, 10) == 0
    assert lib.cs_postprocess_workspace_bytes(1, 70000, 70000) == 0           # beyond 32-bit pixel ids
    assert b"32-bit" in lib.cs_last_error()
    buf = (ctypes.c_uint8 * 16)()
    lab = (ctypes.c_int32 * 16)()
    assert lib.cs_postprocess(buf, 1, 4, 4, 64, 16, 3, lab, None, buf, 16, None) == -1   # connectivity 3
    assert b"connectivity" in lib.cs_last_error()
    assert lib.cs_postprocess(buf, 1, 4, 4, 64, 16, 8, lab, None, buf, 16, None) == -2   # workspace too small
    assert lib.cs_relabel_workspace_bytes(-1) == 0
    assert lib.cs_overlap_workspace_bytes(1024) > 1024 * 12
    assert lib.cs_kernel_launch_count() == 0                                  # nothing was launched


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cellseg_b200 import capi, compute_metric, postprocess
    with pytest.raises(capi.CellSegError):
        postprocess.postprocess_device(torch.zeros((4, 4), dtype=torch.bool))
    with pytest.raises(capi.CellSegError):
        compute_metric.eval_tp_fp_fn(np.ones((4, 4), np.int32), np.ones((4, 4), np.int32))
    with pytest.raises(capi.CellSegError):
        postprocess.label(np.ones((4, 4), bool))                             # *_host entry point: cudaMalloc fails → error code


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from cellseg_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(capi.CellSegError, match="no CPU fallback"):
        capi.load()
