"""Host-side mirror of the skimage calls at baseline/predict.py:113

    measure.label(morphology.remove_small_objects(morphology.remove_small_holes(mask), 16))

Same names, argument meaning and defaults as skimage for boolean input; the work is done by
libcellseg_sm100.so.  numpy in → numpy out goes through the `*_host` C entry points; CUDA tensors in →
CUDA tensors out stay on device on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _conn(connectivity, default):
    """skimage connectivity (1 = 4-neighbourhood, 2 = 8-neighbourhood) → pixel-neighbour count."""
    c = default if connectivity is None else connectivity
    if c not in (1, 2):
        raise ValueError("connectivity must be 1 or 2 for 2-D images")
    return 4 if c == 1 else 8


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


class DeviceWorkspace:
    """Grow-only device scratch buffer (one per device/stream user)."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes, device):
        import torch
        if self._buf is None or self._buf.numel() < nbytes or self._buf.device != device:
            self._buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self._buf


_default_ws = DeviceWorkspace()


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_batch_u8(mask):
    import torch
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8) if mask.is_contiguous() else mask.contiguous().view(torch.uint8)
    elif mask.dtype != torch.uint8:
        mask = (mask != 0).view(torch.uint8)
    mask = mask.contiguous()
    squeeze = mask.dim() == 2
    if squeeze:
        mask = mask.unsqueeze(0)
    if mask.dim() != 3:
        raise ValueError("mask must be [H,W] or [N,H,W]")
    return mask, squeeze


def postprocess_device(mask, hole_area=64, min_size=16, connectivity=2, workspace=None, out=None):
    """CUDA bool/u8 mask [N,H,W] or [H,W] → (labels int32, counts int32[N]) on the same device."""
    import torch
    if not mask.is_cuda:
        raise capi.CellSegError("postprocess_device needs a CUDA tensor; cellseg_b200 has no CPU path")
    lib = capi.load()
    m, squeeze = _as_batch_u8(mask)
    N, H, W = m.shape
    wsb = lib.cs_postprocess_workspace_bytes(N, H, W)
    if wsb == 0:
        capi.check(lib, -1)
    ws = (workspace or _default_ws).get(wsb, m.device)
    labels = out if out is not None else torch.empty((N, H, W), dtype=torch.int32, device=m.device)
    counts = torch.empty((N,), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        capi.check(lib, lib.cs_postprocess(m.data_ptr(), N, H, W, int(hole_area), int(min_size), _conn(connectivity, 2),
                                           labels.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return (labels[0] if squeeze else labels), counts


def _host_mask(mask):
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8) if np.asarray(mask).dtype != np.uint8 \
        else np.ascontiguousarray(mask)
    if m.ndim != 2:
        raise ValueError("only 2-D masks are supported (the reference path is 2-D)")
    return m


def label(mask, connectivity=None):
    """skimage.measure.label for a boolean 2-D image: int32 ids 1..n in raster order of first pixel."""
    if _is_tensor(mask):
        return postprocess_device(mask, 0, 0, connectivity if connectivity is not None else 2)[0]
    lib = capi.load()
    m = _host_mask(mask)
    out = np.empty(m.shape, dtype=np.int32)
    capi.check(lib, lib.cs_label_host(m.ctypes.data, 1, m.shape[0], m.shape[1], _conn(connectivity, 2), out.ctypes.data, None))
    return out


def _remove_small(mask, size, connectivity, invert):
    if _conn(connectivity, 1) != 4:
        raise NotImplementedError("only skimage's default connectivity=1 is on the reference path")
    lib = capi.load()
    if _is_tensor(mask):
        import torch
        m, squeeze = _as_batch_u8(mask)
        N, H, W = m.shape
        wsb = lib.cs_postprocess_workspace_bytes(N, H, W)
        ws = _default_ws.get(wsb, m.device)
        out = torch.empty_like(m)
        with torch.cuda.device(m.device):
            capi.check(lib, lib.cs_remove_small(m.data_ptr(), out.data_ptr(), N, H, W, int(size), int(invert),
                                                ws.data_ptr(), ws.numel(), _stream_ptr()))
        out = out.view(torch.bool)
        return out[0] if squeeze else out
    m = _host_mask(mask)
    out = np.empty_like(m)
    capi.check(lib, lib.cs_remove_small_host(m.ctypes.data, out.ctypes.data, 1, m.shape[0], m.shape[1], int(size), int(invert)))
    return out.view(np.bool_)


def remove_small_objects(ar, min_size=64, connectivity=1):
    """skimage.morphology.remove_small_objects for a boolean image (components < min_size removed)."""
    return _remove_small(ar, min_size, connectivity, 0)


def remove_small_holes(ar, area_threshold=64, connectivity=1):
    """skimage.morphology.remove_small_holes for a boolean image (holes < area_threshold filled)."""
    return _remove_small(ar, area_threshold, connectivity, 1)


def certified_large(mask, hole_area=64, min_size=16):
    """Diagnostic: (large_bg, large_fg) bool tensors — the pixels the size filters' pre-pass certifies, without labelling, to
    lie in components of at least `hole_area` (background) / `min_size` (foreground) pixels (csrc/certify.cu)."""
    import torch
    if not mask.is_cuda:
        raise capi.CellSegError("certified_large needs a CUDA tensor; cellseg_b200 has no CPU path")
    lib = capi.load()
    m, squeeze = _as_batch_u8(mask)
    N, H, W = m.shape
    wsb = lib.cs_postprocess_workspace_bytes(N, H, W)
    ws = _default_ws.get(wsb, m.device)
    bg, fg = torch.empty_like(m), torch.empty_like(m)
    with torch.cuda.device(m.device):
        capi.check(lib, lib.cs_certify_large(m.data_ptr(), N, H, W, int(hole_area), int(min_size), bg.data_ptr(), fg.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream_ptr()))
    bg, fg = bg.view(torch.bool), fg.view(torch.bool)
    return (bg[0], fg[0]) if squeeze else (bg, fg)
