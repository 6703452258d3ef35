"""TEST INFRASTRUCTURE ONLY: load the g++/fiber-emulated build of the kernels and
call the same C ABI with numpy (host) pointers."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

_lib = None


def lib():
    global _lib
    if _lib is None:
        import build_emu
        from cellseg_b200 import capi
        _lib = capi.bind(C.CDLL(build_emu.build()))
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def check(rc):
    if rc != 0:
        raise RuntimeError(f"emu status {rc}: {lib().cs_last_error().decode()}")


def postprocess(mask, hole_area=64, min_size=16, connectivity=8):
    """mask: bool/u8 [N,H,W] or [H,W] → (labels int32, counts int32[N])"""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    squeeze = m.ndim == 2
    if squeeze:
        m = m[None]
    N, H, W = m.shape
    L = lib()
    wsb = L.cs_postprocess_workspace_bytes(N, H, W)
    ws = np.zeros(wsb, dtype=np.uint8)
    labels = np.full((N, H, W), -7, dtype=np.int32)
    n_out = np.zeros(N, dtype=np.int32)
    check(L.cs_postprocess(ptr(m), N, H, W, hole_area, min_size, connectivity, ptr(labels), ptr(n_out), ptr(ws), wsb, None))
    return (labels[0] if squeeze else labels), n_out


def remove_small(mask, min_size, invert):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    squeeze = m.ndim == 2
    if squeeze:
        m = m[None]
    N, H, W = m.shape
    L = lib()
    wsb = L.cs_postprocess_workspace_bytes(N, H, W)
    ws = np.zeros(wsb, dtype=np.uint8)
    out = np.zeros_like(m)
    scratch = np.zeros((N, H, W), dtype=np.int32)
    check(L.cs_remove_small(ptr(m), ptr(out), N, H, W, min_size, int(invert), ptr(scratch), ptr(ws), wsb, None))
    return out[0] if squeeze else out
