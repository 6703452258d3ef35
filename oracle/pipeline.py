"""Oracle restatement of the reference's per-image pipeline, baseline/predict.py:94-113, on the CPU:

    channel fix-up + percentile normalisation (:94-104) → /max, float32 NCHW (:107-108) →
    sliding_window_inference(roi 256², sw_batch_size 4) (:109) → softmax (:110) → class-1 > 0.5 (:111-113) →
    remove_small_holes → remove_small_objects(16) → measure.label (:113)

built from the other oracle modules (sliding_window.py, morph.py).  The predictor is an argument (a torch module
on the CPU; the UNet itself is PyTorch on both sides and out of scope).  TEST INFRASTRUCTURE and the reported
CPU baseline of bench.py — see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import morph
from . import sliding_window as osw


def rescale_intensity_u8(img: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """skimage.exposure.rescale_intensity(img, in_range=(lo, hi), out_range='uint8'): clip, scale to [0, 1] in
    float64, ×255 (+0), cast (the reference then calls .astype(np.uint8), predict.py:21)."""
    x = np.clip(img, float(lo), float(hi))
    x = (x - float(lo)) / (float(hi) - float(lo))
    return np.asarray(x * 255.0 + 0.0)


def normalize_channel(img: np.ndarray, lower: float = 1, upper: float = 99) -> np.ndarray:
    """predict.py:14-21."""
    non_zero_vals = img[np.nonzero(img)]
    percentiles = np.percentile(non_zero_vals, [lower, upper])
    if percentiles[1] - percentiles[0] > 0.001:
        img_norm = rescale_intensity_u8(img, percentiles[0], percentiles[1])
    else:
        img_norm = img
    return img_norm.astype(np.uint8)


def preprocess(img_data: np.ndarray) -> np.ndarray:
    """predict.py:94-104 → u8[H,W,3]."""
    if img_data.ndim == 2:
        img_data = np.repeat(np.expand_dims(img_data, axis=-1), 3, axis=-1)
    elif img_data.ndim == 3 and img_data.shape[-1] > 3:
        img_data = img_data[:, :, :3]
    pre = np.zeros(img_data.shape, dtype=np.uint8)
    for i in range(3):
        ch = img_data[:, :, i]
        if len(ch[np.nonzero(ch)]) > 0:
            pre[:, :, i] = normalize_channel(ch, 1, 99)
    return pre


@torch.no_grad()
def segment(pre_img_u8: np.ndarray, predictor, roi: int = 256, sw_batch_size: int = 4, mode: str = "constant",
            gaussian_variant: str = "A") -> np.ndarray:
    """predict.py:107-113 for one pre-normalised u8[H,W,3] image → int32 instance map."""
    x = osw.scale_to_unit(pre_img_u8)
    out = osw.sliding_window_inference(x, (roi, roi), sw_batch_size, predictor, mode=mode, gaussian_variant=gaussian_variant)
    return morph.postprocess_mask(osw.foreground_mask(out), 64, 16, 2)
