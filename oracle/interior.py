"""Oracle for data/pre_process_3class.py:31-56 `create_interior_map` (instance map → {0 background, 1 interior, 2 boundary}).

skimage is absent offline; its three calls are restated with scipy.ndimage exactly as skimage executes them:
  segmentation.find_boundaries(lab, mode='inner')      connectivity 1:  dilation(lab, cross) != erosion(lab, cross), then & (lab != 0);
                                                       skimage's grey dilation / erosion reflect at the image edge, so the edge itself
                                                       never makes a boundary
  morphology.binary_dilation(b, disk(1))               cross-shaped 3x3 footprint, outside = False
  morphology.remove_small_objects(m, min_size=16)      4-connected components with < 16 pixels cleared (oracle/morph.py)
The reference converts the labels with `.astype(np.int16)` first (pre_process_3class.py:98): ids ≥ 32768 wrap (negative ids are
"not > 0", id 65536 becomes background) — kept.  TEST INFRASTRUCTURE — see oracle/__init__.py."""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

from . import morph

_CROSS = ndi.generate_binary_structure(2, 1)


def find_boundaries_inner(lab: np.ndarray) -> np.ndarray:
    mx = ndi.grey_dilation(lab, footprint=_CROSS, mode="reflect")
    mn = ndi.grey_erosion(lab, footprint=_CROSS, mode="reflect")
    return (mx != mn) & (lab != 0)


def create_interior_map(inst_map: np.ndarray) -> np.ndarray:
    """pre_process_3class.py:31-56 on `inst_map` as the reference passes it (already int16)."""
    inst_map = np.asarray(inst_map)
    boundary = find_boundaries_inner(inst_map)
    boundary = ndi.binary_dilation(boundary, structure=_CROSS, border_value=0)
    interior_temp = np.logical_and(~boundary, inst_map > 0)
    interior_temp = morph.remove_small_objects(interior_temp, 16)
    interior = np.zeros(inst_map.shape, dtype=np.uint8)
    interior[interior_temp] = 1
    interior[boundary] = 2
    return interior


def find_boundaries_inner_direct(lab: np.ndarray) -> np.ndarray:
    """Independent cross-check of find_boundaries_inner: a foreground pixel with a 4-neighbour (inside the image) of another value."""
    H, W = lab.shape
    p = np.pad(lab, 1, mode="edge")
    c = p[1:-1, 1:-1]
    diff = (p[:-2, 1:-1] != c) | (p[2:, 1:-1] != c) | (p[1:-1, :-2] != c) | (p[1:-1, 2:] != c)
    return diff & (lab != 0)
