"""Image I/O for the CLIs without tifffile / skimage.io (neither is installable offline).

Reads what the reference reads (`tif.imread` for .tif/.tiff, `skimage.io.imread` otherwise,
baseline/predict.py:88-91, compute_metric.py:171-178) through Pillow, and writes the instance map as
an int32 TIFF with Adobe-deflate compression — the same TIFF compression code (8) that
`tif.imwrite(..., compression='zlib')` emits at baseline/predict.py:114.  tifffile is used when present.
"""
from __future__ import annotations

import numpy as np


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


def imwrite_labels(path: str, labels: np.ndarray) -> None:
    """int32 instance map → deflate-compressed TIFF."""
    try:
        import tifffile
        tifffile.imwrite(path, labels, compression="zlib")
        return
    except ImportError:
        pass
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(labels, dtype=np.int32), mode="I").save(path, format="TIFF", compression="tiff_adobe_deflate")


def imsave(path: str, img: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(img).save(path)
