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


_TIFF_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q"}


def read_tiff(path: str) -> np.ndarray:
    """First image of a classic (non-Big) TIFF stored in strips or tiles, chunky or planar, 8/16/32/64-bit unsigned, signed or
    float samples, uncompressed, deflate (8 / 32946) or PackBits (32773), optional horizontal predictor — the layouts of
    microscopy exports (16-bit multi-channel included) that Pillow refuses.  → [H,W] or [H,W,C] (planar files come back
    channels-last, as tifffile's imread followed by the reference's own use expects H,W,C)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    bo = {b"II": "<", b"MM": ">"}.get(buf[:2])
    if bo is None or struct.unpack(bo + "H", buf[2:4])[0] != 42:
        raise ValueError(f"read_tiff: {path} is not a classic TIFF (BigTIFF and other containers are not supported)")
    ifd = struct.unpack(bo + "I", buf[4:8])[0]
    n = struct.unpack(bo + "H", buf[ifd:ifd + 2])[0]
    tags = {}
    for i in range(n):
        tag, typ, cnt, val = struct.unpack(bo + "HHI4s", buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i])
        fmt = _TIFF_TYPES.get(typ)
        if fmt is None:
            continue
        size = struct.calcsize(bo + fmt) * cnt
        raw = val[:size] if size <= 4 else buf[struct.unpack(bo + "I", val)[0]:][:size]
        tags[tag] = struct.unpack(bo + fmt * cnt, raw)
    W, H = tags[256][0], tags[257][0]
    C = tags.get(277, (1,))[0]
    bits = tags.get(258, (1,))
    if len(set(bits)) != 1 or bits[0] not in (8, 16, 32, 64):
        raise ValueError(f"read_tiff: unsupported bits per sample {bits} in {path}")
    kind = {1: "u", 2: "i", 3: "f"}.get(tags.get(339, (1,))[0], "u")
    dt = np.dtype(f"{bo}{kind}{bits[0] // 8}")
    comp = tags.get(259, (1,))[0]
    pred = tags.get(317, (1,))[0]
    planar = tags.get(284, (1,))[0] == 2 and C > 1
    if tags.get(262, (1,))[0] == 3:
        raise ValueError(f"read_tiff: palette TIFF {path}: convert it to RGB first")

    def decode(raw: bytes, nbytes: int) -> bytes:
        if comp == 1:
            return raw[:nbytes]
        if comp in (8, 32946):
            return zlib.decompress(raw)
        if comp == 32773:
            out, i = bytearray(), 0
            while i < len(raw) and len(out) < nbytes:
                c = raw[i]; i += 1
                if c < 128:
                    out += raw[i:i + c + 1]; i += c + 1
                elif c > 128:
                    out += raw[i:i + 1] * (257 - c); i += 1
            return bytes(out)
        raise ValueError(f"read_tiff: compression {comp} in {path} is not supported (LZW/JPEG need tifffile)")

    spp = 1 if planar else C                 # samples per pixel inside one chunk
    planes = C if planar else 1
    out = np.zeros((planes, H, W, spp), dtype=dt.newbyteorder("="))
    if 322 in tags:                          # tiles
        tw, tl = tags[322][0], tags[323][0]
        offs, cnts = tags[324], tags[325]
        tx, ty = -(-W // tw), -(-H // tl)
        for k, (o, c) in enumerate(zip(offs, cnts)):
            pl, r = divmod(k, tx * ty)
            y0, x0 = (r // tx) * tl, (r % tx) * tw
            a = np.frombuffer(decode(buf[o:o + c], tl * tw * spp * dt.itemsize), dtype=dt, count=tl * tw * spp).reshape(tl, tw, spp)
            if pred == 2:
                a = np.cumsum(a, axis=1, dtype=dt)
            out[pl, y0:y0 + tl, x0:x0 + tw] = a[:H - y0, :W - x0]
    else:
        rps = min(tags.get(278, (H,))[0], H)
        offs, cnts = tags[273], tags[279]
        per_plane = -(-H // rps)
        for k, (o, c) in enumerate(zip(offs, cnts)):
            pl, r = divmod(k, per_plane)
            y0 = r * rps
            rows = min(rps, H - y0)
            a = np.frombuffer(decode(buf[o:o + c], rows * W * spp * dt.itemsize), dtype=dt, count=rows * W * spp).reshape(rows, W, spp)
            if pred == 2:
                a = np.cumsum(a, axis=1, dtype=dt)
            out[pl, y0:y0 + rows] = a
    if planar:
        out = np.moveaxis(out[..., 0], 0, -1)
    else:
        out = out[0]
    return out[..., 0] if out.shape[-1] == 1 else out


def imread(path: str) -> np.ndarray:
    """`tif.imread` / `skimage.io.imread` stand-in (baseline/predict.py:88-91).  TIFFs go through `read_tiff` (Pillow is the
    fallback for LZW / JPEG compression); palette images come back as RGB(A) like skimage's reader returns them."""
    if path.endswith(".npy"):
        return np.load(path)
    is_tiff = path.lower().endswith((".tif", ".tiff"))
    try:
        import tifffile
        if is_tiff:
            return tifffile.imread(path)
    except ImportError:
        pass
    tiff_err = None
    if is_tiff:
        # the built-in reader first: Pillow silently reduces 16-bit multi-channel TIFFs to 8 bits per sample
        try:
            return read_tiff(path)
        except Exception as e:               # LZW / JPEG / BigTIFF …: let Pillow try
            tiff_err = e
    from PIL import Image
    limit, Image.MAX_IMAGE_PIXELS = Image.MAX_IMAGE_PIXELS, None      # whole-slide images exceed Pillow's bomb guard; restored below
    try:
        with Image.open(path) as im:
            if im.mode in ("P", "PA"):
                im = im.convert("RGBA" if (im.mode == "PA" or "transparency" in im.info) else "RGB")
            if is_tiff and im.mode == "RGB" and max(im.tag_v2.get(258, (8,))) > 8:
                raise ValueError(f"cannot decode {path} without losing precision: Pillow would reduce its {im.tag_v2.get(258)}-bit "
                                 f"samples to 8 bits; built-in TIFF reader: {tiff_err}")
            a = np.array(im)
    except ValueError:
        raise
    except Exception as e:
        raise ValueError(f"cannot decode {path}: Pillow: {e}" + (f"; built-in TIFF reader: {tiff_err}" if tiff_err else "")) from e
    finally:
        Image.MAX_IMAGE_PIXELS = limit
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


def imwrite_labels(path: str, labels: np.ndarray, dtype=np.int32) -> None:
    """Instance map → deflate-compressed TIFF (predict.py:114).  The sample type is int32 — what `scipy.ndimage.label`, to which
    `skimage.measure.label` hands boolean input, returns and what the reference's sample label files hold
    (data/Train_Labeled/labels/*.tiff).  A scikit-image build whose `label` returns intp would make the reference write int64
    samples; the pixel values are the same, and `dtype=np.int64` is not offered because classic TIFF readers (Pillow included)
    do not read 64-bit samples."""
    labels = np.ascontiguousarray(labels, dtype=dtype)
    try:
        import tifffile
        tifffile.imwrite(path, labels, compression="zlib")
        return
    except ImportError:
        pass
    write_tiff(path, labels, compress=True)


def imsave(path: str, img: np.ndarray) -> None:
    """skimage.io.imsave stand-in for the overlay image (predict.py:122)."""
    if path.lower().endswith((".tif", ".tiff")):
        write_tiff(path, img, compress=False)
        return
    from PIL import Image
    Image.fromarray(img).save(path)
