"""Oracle restatement (numpy/scipy) of the reference's F1 metric, baseline/compute_metric.py.

Function-by-function, with the reference lines each one follows.  Pinned against outputs of the
reference file itself executed in the build container: tests/golden/metric_*.npz, produced by
oracle/make_golden.py (which imports /root/reference/baseline/compute_metric.py with stub
skimage/tifffile modules).  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

from .morph import relabel_sequential


def label_overlap(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """compute_metric.py:39-70 — overlap[i, j] = #pixels with x == i and y == j, uint64,
    shape (1 + x.max(), 1 + y.max())."""
    x = np.asarray(x).ravel().astype(np.int64)
    y = np.asarray(y).ravel().astype(np.int64)
    nx, ny = int(x.max()) + 1, int(y.max()) + 1
    flat = np.bincount(x * ny + y, minlength=nx * ny)
    return flat.reshape(nx, ny).astype(np.uint64)


def intersection_over_union(masks_true: np.ndarray, masks_pred: np.ndarray) -> np.ndarray:
    """compute_metric.py:21-37 — IoU of every (true, pred) label pair incl. background row/col; 0/0 → 0."""
    overlap = label_overlap(masks_true, masks_pred)
    n_pred = overlap.sum(axis=0, keepdims=True)
    n_true = overlap.sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = overlap / (n_pred + n_true - overlap)
    iou[np.isnan(iou)] = 0.0
    return iou


def true_positive(iou: np.ndarray, th: float) -> int:
    """compute_metric.py:83-105 — Hungarian assignment on -(iou>=th) - iou/(2*n_min); count matches >= th."""
    n_min = min(iou.shape[0], iou.shape[1])
    costs = -(iou >= th).astype(float) - iou / (2 * n_min)
    ti, pi = linear_sum_assignment(costs)
    return int((iou[ti, pi] >= th).sum())


def eval_tp_fp_fn(masks_true: np.ndarray, masks_pred: np.ndarray, threshold: float = 0.5):
    """compute_metric.py:107-122 — note the quirk: an empty prediction yields (0, 0, 0), not fn = n_gt."""
    n_gt, n_seg = int(np.max(masks_true)), int(np.max(masks_pred))
    if n_seg > 0:
        iou = intersection_over_union(masks_true, masks_pred)[1:, 1:]
        tp = true_positive(iou, threshold)
        return tp, n_seg - tp, n_gt - tp
    return 0, 0, 0


def remove_boundary_cells(mask: np.ndarray) -> np.ndarray:
    """compute_metric.py:124-133 — drop every label with a pixel in the 2-pixel frame (mutating `mask`
    like the reference), then relabel_sequential."""
    h, w = mask.shape
    frame = np.ones((h, w))
    frame[2:h - 2, 2:w - 2] = 0
    touching = np.unique(mask * frame)
    for lab in touching[1:]:          # the reference skips the first unique value whatever it is
        mask[mask == lab] = 0
    return relabel_sequential(mask)[0]


def dice(gt: np.ndarray, seg: np.ndarray) -> float:
    """compute_metric.py:72-81."""
    g, s = int(np.count_nonzero(gt)), int(np.count_nonzero(seg))
    if g == 0 and s == 0:
        return 1.0
    if g == 0 and s > 0:
        return 0.0
    return 2 * int(np.count_nonzero(np.logical_and(gt, seg))) / (g + s)


def evaluate_pair(gt: np.ndarray, seg: np.ndarray, threshold: float = 0.5, count_bd_cells: bool = False, roi_size: int = 2000):
    """compute_metric.py:179-252 for one (gt, seg) pair → dict with the CSV row's fields."""
    dice_score = dice(gt > 0, seg > 0)
    if np.prod(gt.shape) < 25000000:
        if not count_bd_cells:
            gt = remove_boundary_cells(gt.astype(np.int32))
            seg = remove_boundary_cells(seg.astype(np.int32))
        gt = relabel_sequential(gt)[0]
        seg = relabel_sequential(seg)[0]
        n_true, n_pred = int(np.max(gt)), int(np.max(seg))
        tp, fp, fn = eval_tp_fp_fn(gt, seg, threshold)
    else:
        if count_bd_cells:
            raise NameError("gt_roi")   # the reference's large-image branch fails this way (:223-227) → 'failed' list
        H, W = gt.shape
        n_h, n_w = -(-H // roi_size), -(-W // roi_size)
        gt_pad = np.zeros((n_h * roi_size, n_w * roi_size), dtype=gt.dtype)
        seg_pad = np.zeros_like(gt_pad)
        gt_pad[:H, :W] = gt
        seg_pad[:H, :W] = seg
        tp = fp = fn = n_true = n_pred = 0
        for i in range(n_h):
            for j in range(n_w):
                sl = (slice(roi_size * i, roi_size * (i + 1)), slice(roi_size * j, roi_size * (j + 1)))
                g = relabel_sequential(remove_boundary_cells(gt_pad[sl]))[0]
                s = relabel_sequential(remove_boundary_cells(seg_pad[sl]))[0]
                n_true += int(np.max(g)); n_pred += int(np.max(s))
                a, b, c = eval_tp_fp_fn(g, s, threshold)
                tp += a; fp += b; fn += c
    if tp == 0:
        precision = recall = f1 = 0
    else:
        precision, recall = tp / n_pred, tp / n_true
        f1 = 2 * (precision * recall) / (precision + recall)
    return dict(true_num=n_true, pred_num=n_pred, tp=tp, fn=fn, fp=fp, precision=np.round(precision, 4),
                recall=np.round(recall, 4), dice=np.round(dice_score, 4), f1=np.round(f1, 4))
