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
    check(L.cs_remove_small(ptr(m), ptr(out), N, H, W, min_size, int(invert), ptr(ws), wsb, None))
    return out[0] if squeeze else out


# ---- sliding-window kernels on host pointers (mirrors cellseg_b200.sliding_window wrappers) ----
def scale_lut(img_u8):
    img = np.ascontiguousarray(img_u8)
    N = img.shape[0]
    scratch = np.zeros(N, np.uint32)
    lut = np.zeros((N, 256), np.float32)
    check(lib().cs_scale_lut_u8(ptr(img), N, img[0].size, ptr(scratch), ptr(lut), None))
    return lut


def to_unit_float(img_u8, lut):
    img = np.ascontiguousarray(img_u8)
    N, H, W, _ = img.shape
    out = np.zeros((N, 3, H, W), np.float32)
    check(lib().cs_u8hwc_to_f32chw(ptr(img), N, H, W, ptr(lut), ptr(out), None))
    return out


def plan_arrays(plan):
    from cellseg_b200 import sliding_window as sw
    Hp, Wp = max(plan.H, plan.roi), max(plan.W, plan.roi)
    return (np.asarray(plan.ys, np.int32), np.asarray(plan.xs, np.int32),
            sw._cover_table(plan.ys, plan.roi, Hp), sw._cover_table(plan.xs, plan.roi, Wp))


def gather(src, plan, begin, count, lut=None):
    ys, xs, _, _ = plan_arrays(plan)
    src = np.ascontiguousarray(src)
    out = np.full((count, 3, plan.roi, plan.roi), np.nan, np.float32)
    L = lib()
    if src.dtype == np.uint8:
        check(L.cs_gather_windows_u8(ptr(src), ptr(lut), src.shape[0], plan.H, plan.W, plan.roi, plan.ny, plan.nx,
                                     plan.pad_y, plan.pad_x, ptr(ys), ptr(xs), begin, count, ptr(out), None))
    else:
        check(L.cs_gather_windows(ptr(src), src.shape[0], plan.H, plan.W, plan.roi, plan.ny, plan.nx,
                                  plan.pad_y, plan.pad_x, ptr(ys), ptr(xs), begin, count, ptr(out), None))
    return out


def blend(win_logits, N, plan, imap=None, threshold=0.5, packed=False):
    ys, xs, cy, cx = plan_arrays(plan)
    wl = np.ascontiguousarray(win_logits, np.float32)
    Cc = wl.shape[1]
    blended = np.zeros((N, Cc, plan.H, plan.W), np.float32)
    p1 = np.zeros((N, plan.H, plan.W), np.float32)
    mask = np.zeros((N, plan.H, plan.W), np.uint8)
    L = lib()
    ws = slot = None
    if packed:
        wsb = L.cs_postprocess_workspace_bytes(N, plan.H, plan.W)
        ws = np.zeros(wsb, np.uint8)
        slot = L.cs_postprocess_packed_slot(ptr(ws), wsb, N, plan.H, plan.W)
    im = np.ascontiguousarray(imap, np.float32) if imap is not None else None
    check(L.cs_blend_softmax_threshold(ptr(wl), Cc, N, plan.H, plan.W, plan.roi, plan.ny, plan.nx, plan.pad_y, plan.pad_x,
                                       ptr(ys), ptr(xs), ptr(cy), ptr(cx), ptr(im), threshold, slot, ptr(mask), ptr(p1),
                                       ptr(blended), None))
    out = dict(blended=blended, p1=p1, mask=mask)
    if packed:
        labels = np.zeros((N, plan.H, plan.W), np.int32)
        n_out = np.zeros(N, np.int32)
        check(L.cs_postprocess_packed(N, plan.H, plan.W, 64, 16, 8, ptr(labels), ptr(n_out), ptr(ws), len(ws), None))
        out["labels"] = labels
    return out


def certify_large(mask, hole_area=64, min_size=16):
    """u8/bool [H,W] → (large_bg, large_fg) bool maps of the certified-large pre-pass (certify.cu)."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)[None]
    N, H, W = m.shape
    L = lib()
    wsb = L.cs_postprocess_workspace_bytes(N, H, W)
    ws = np.zeros(wsb, dtype=np.uint8)
    bg = np.zeros_like(m); fg = np.zeros_like(m)
    check(L.cs_certify_large(ptr(m), N, H, W, hole_area, min_size, ptr(bg), ptr(fg), ptr(ws), wsb, None))
    return bg[0].astype(bool), fg[0].astype(bool)
