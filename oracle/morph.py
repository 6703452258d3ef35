"""Oracle for the mask → instance-map chain of baseline/predict.py:113

    measure.label(morphology.remove_small_objects(morphology.remove_small_holes(mask), 16))

skimage (un-pinned in requirements.txt:3) is absent offline; for boolean input each of the three
calls reduces to scipy.ndimage.label + numpy.bincount, restated here.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

_CONN4 = ndi.generate_binary_structure(2, 1)
_CONN8 = np.ones((3, 3), dtype=bool)


def _structure(connectivity: int):
    if connectivity in (1, 4):
        return _CONN4
    if connectivity in (2, 8):
        return _CONN8
    raise ValueError("connectivity must be 1/4 or 2/8")


def remove_small_objects(mask: np.ndarray, min_size: int = 64, connectivity: int = 1) -> np.ndarray:
    """skimage.morphology.remove_small_objects on a bool array: components (default 4-connected)
    with strictly fewer than `min_size` pixels are cleared."""
    out = np.array(mask, dtype=bool, copy=True)
    if min_size == 0:
        return out
    cc, _ = ndi.label(out, _structure(connectivity))
    sizes = np.bincount(cc.ravel())
    too_small = sizes < min_size
    out[too_small[cc]] = False
    return out


def remove_small_holes(mask: np.ndarray, area_threshold: int = 64, connectivity: int = 1) -> np.ndarray:
    """skimage.morphology.remove_small_holes: remove_small_objects on the complement."""
    inv = ~np.asarray(mask, dtype=bool)
    inv = remove_small_objects(inv, area_threshold, connectivity)
    return ~inv


def label(mask: np.ndarray, connectivity: int = 2) -> np.ndarray:
    """skimage.measure.label on a bool array (default full connectivity): ids 1..n in raster
    order of each component's first pixel; int32 like the reference's TIFF output."""
    cc, _ = ndi.label(np.asarray(mask, dtype=bool), _structure(connectivity))
    return cc.astype(np.int32)


def postprocess_mask(mask: np.ndarray, hole_area: int = 64, min_size: int = 16, connectivity: int = 2) -> np.ndarray:
    """The full chain of baseline/predict.py:113 on one 2-D bool mask."""
    m = np.asarray(mask, dtype=bool)
    if hole_area > 0:
        m = remove_small_holes(m, hole_area)
    if min_size > 0:
        m = remove_small_objects(m, min_size)
    return label(m, connectivity)


def label_cv2(mask: np.ndarray, connectivity: int = 8) -> np.ndarray:
    """Independent cross-check: OpenCV's Wu two-pass labelling produces the same raster-ordered ids."""
    import cv2
    _, lab = cv2.connectedComponentsWithAlgorithm(np.asarray(mask, dtype=np.uint8), connectivity, cv2.CV_32S, cv2.CCL_WU)
    return lab.astype(np.int32)


def relabel_sequential(label_field: np.ndarray, offset: int = 1):
    """skimage.segmentation.relabel_sequential: non-zero labels mapped to offset.. in ascending
    order, 0 kept; returns (relabeled, forward_map, inverse_map) with arrays for the maps."""
    lf = np.asarray(label_field)
    vals = np.unique(lf)
    nz = vals[vals != 0]
    new_max = offset + len(nz) - 1 if len(nz) else 0
    out_dtype = lf.dtype
    if new_max > np.iinfo(out_dtype).max:
        out_dtype = np.min_scalar_type(new_max)
    fw = np.zeros(int(vals.max()) + 1 if len(vals) else 1, dtype=out_dtype)
    fw[nz] = np.arange(offset, offset + len(nz), dtype=out_dtype)
    inv = np.zeros(new_max + 1, dtype=lf.dtype)
    inv[offset:offset + len(nz)] = nz
    return fw[lf], fw, inv
