"""Image I/O for the CLIs without tifffile / skimage.io (neither is installable offline).

Reads what the reference reads (`tif.imread` for .tif/.tiff, `skimage.io.imread` otherwise,
baseline/predict.py:88-91, compute_metric.py:171-178) through Pillow.  Writes TIFFs with a small writer of its own:
the instance map goes out as an int32 TIFF with Adobe-deflate compression — the same TIFF compression code (8) that
`tif.imwrite(..., compression='zlib')` emits at baseline/predict.py:114 — cut into row strips that are deflated in
parallel on the host cores (zlib releases the GIL), because this write sits inside the reference's per-image timer and
is single-threaded there.  tifffile is used when present.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_SAMPLE_FORMAT = {"u": 1, "i": 2, "f": 3}
_pool = None


def _executor() -> ThreadPoolExecutor:
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1)))
    return _pool


def imread(path: str) -> np.ndarray:
    if path.endswith(".npy"):
        return np.load(path)
    try:
        import tifffile
        if path.endswith((".tif", ".tiff")):
            return tifffile.imread(path)
    except ImportError:
        pass
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        a = np.array(im)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8) * 255
    return a


def write_tiff(path: str, arr: np.ndarray, compress: bool = True, level: int = 6, rows_per_strip: int | None = None) -> None:
    """Classic little-endian baseline TIFF: [H,W] or [H,W,C] of uint8/uint16/uint32/int8/int16/int32/float32, one IFD,
    row strips, optional Adobe-deflate (compression 8, no predictor — what tifffile's compression='zlib' writes)."""
    a = np.ascontiguousarray(arr)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3 or a.dtype.kind not in _SAMPLE_FORMAT or a.dtype.itemsize > 4:
        raise ValueError(f"write_tiff: unsupported array {arr.shape} {arr.dtype}")
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)
    H, W, C = a.shape
    if rows_per_strip is None:
        rows_per_strip = max(1, min(H, (1 << 20) // max(1, W * C * a.dtype.itemsize)))      # ≈ 1 MiB of pixels per strip
    bounds = [(r, min(H, r + rows_per_strip)) for r in range(0, H, rows_per_strip)]
    if compress:
        strips = list(_executor().map(lambda b: zlib.compress(a[b[0]:b[1]].tobytes(), level), bounds))
    else:
        strips = [a[r0:r1].tobytes() for r0, r1 in bounds]
    n = len(strips)
    offsets, pos = [], 8
    for s in strips:
        offsets.append(pos)
        pos += len(s) + (len(s) & 1)                         # word-align every strip
    if pos + 4096 + 8 * n >= 1 << 32:
        raise ValueError("write_tiff: image too large for classic TIFF")
    extra = bytearray()                                      # out-of-line tag values, placed after the strips
    extra_base = pos

    def outofline(fmt: str, vals) -> int:
        off = extra_base + len(extra)
        extra.extend(struct.pack("<" + fmt * len(vals), *vals))
        if len(extra) & 1:
            extra.append(0)
        return off

    bits = a.dtype.itemsize * 8
    entries = []                                             # (tag, type, count, value-or-offset)

    def short_tag(tag, vals):
        if len(vals) <= 2:
            entries.append((tag, 3, len(vals), struct.unpack("<I", struct.pack("<HH", *(list(vals) + [0])[:2]))[0]))
        else:
            entries.append((tag, 3, len(vals), outofline("H", vals)))

    def long_tag(tag, vals):
        entries.append((tag, 4, len(vals), vals[0] if len(vals) == 1 else outofline("I", vals)))

    long_tag(256, [W]); long_tag(257, [H])
    short_tag(258, [bits] * C)
    short_tag(259, [8 if compress else 1])
    short_tag(262, [2 if C in (3, 4) and a.dtype.kind == "u" else 1])
    long_tag(273, offsets)
    short_tag(277, [C])
    long_tag(278, [rows_per_strip])
    long_tag(279, [len(s) for s in strips])
    short_tag(284, [1])
    if C == 2 or C > 3 or (C == 3 and a.dtype.kind != "u"):
        nextra = C - (3 if C >= 3 and a.dtype.kind == "u" else 1)
        short_tag(338, [0] * nextra)
    short_tag(339, [_SAMPLE_FORMAT[a.dtype.kind]] * C)
    entries.sort(key=lambda e: e[0])
    ifd_off = extra_base + len(extra)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        for s in strips:
            fh.write(s)
            if len(s) & 1:
                fh.write(b"\0")
        fh.write(bytes(extra))
        fh.write(struct.pack("<H", len(entries)))
        for tag, typ, cnt, val in entries:
            fh.write(struct.pack("<HHII", tag, typ, cnt, val))
        fh.write(struct.pack("<I", 0))


def imwrite_labels(path: str, labels: np.ndarray) -> None:
    """int32 instance map → deflate-compressed TIFF (predict.py:114)."""
    try:
        import tifffile
        tifffile.imwrite(path, labels, compression="zlib")
        return
    except ImportError:
        pass
    write_tiff(path, np.ascontiguousarray(labels, dtype=np.int32), compress=True)


def imsave(path: str, img: np.ndarray) -> None:
    """skimage.io.imsave stand-in for the overlay image (predict.py:122)."""
    if path.lower().endswith((".tif", ".tiff")):
        write_tiff(path, img, compress=False)
        return
    from PIL import Image
    Image.fromarray(img).save(path)
