"""Oracle restatement of monai.inferers.sliding_window_inference as the reference calls it
(baseline/predict.py:109: roi (256,256), sw_batch_size 4, defaults overlap=0.25, mode="constant",
sigma_scale=0.125, padding_mode="constant", cval=0) — SURVEY.md §8 a-2 / a-2g.

monai is not installable offline and the reference vendors no copy, so this follows MONAI's published
algorithm (0.9–1.1 code path: pad → scan intervals → dense_patch_slices → per-window
`out += w * pred; cnt += w` in ascending window order → `out / cnt` → crop).  "Parity unpinned" against
monai proper; see oracle/__init__.py.  Plain torch on CPU, Python loops.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def scan_interval(image_size: int, roi: int, overlap: float) -> int:
    """monai.inferers.utils._get_scan_interval for one axis."""
    if roi == image_size:
        return int(roi)
    interval = int(roi * (1 - overlap))
    return interval if interval > 0 else 1


def window_starts(image_size: int, roi: int, overlap: float) -> list[int]:
    """monai.data.utils.dense_patch_slices for one axis (image_size is the PADDED size >= roi)."""
    interval = scan_interval(image_size, roi, overlap)
    num = int(math.ceil(float(image_size) / interval))
    scan_dim = next((d for d in range(num) if d * interval + roi >= image_size), None)
    scan_num = scan_dim + 1 if scan_dim is not None else 1
    starts = []
    for idx in range(scan_num):
        s = idx * interval
        s -= max(s + roi - image_size, 0)
        starts.append(s)
    return starts


def gaussian_1d_erf(sigma: float, truncated: float = 4.0) -> torch.Tensor:
    """monai.networks.layers.gaussian_1d(approx="erf", normalize=False)."""
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float32)
    t = 0.70710678 / abs(sigma)
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def importance_map(roi: int, mode: str = "constant", sigma_scale: float = 0.125, variant: str = "A") -> torch.Tensor:
    """monai.data.utils.compute_importance_map for a square 2-D ROI, followed by the clamp
    sliding_window_inference applies (`min = max(map[map != 0].min(), 1e-3)`).

    variant "A" = monai 0.9–1.1: unit impulse at roi//2 filtered by a separable erf-Gaussian,
    divided by its max.  variant "B" = monai >= 1.2: analytic exp(-d²/2σ²) about (roi-1)/2."""
    if mode == "constant":
        m = torch.ones((roi, roi), dtype=torch.float32)
    elif mode == "gaussian":
        sigma = sigma_scale * roi
        if variant == "A":
            imp = torch.zeros((1, 1, roi, roi), dtype=torch.float32)
            imp[0, 0, roi // 2, roi // 2] = 1.0
            k = gaussian_1d_erf(sigma)
            pad = k.numel() // 2
            imp = torch.nn.functional.conv2d(imp, k.view(1, 1, -1, 1), padding=(pad, 0))
            imp = torch.nn.functional.conv2d(imp, k.view(1, 1, 1, -1), padding=(0, pad))
            m = imp[0, 0]
            m = m / m.max()
        else:
            x = torch.arange(-(roi - 1) / 2.0, (roi - 1) / 2.0 + 1.0, dtype=torch.float32)
            g = torch.exp(x ** 2 / (-2.0 * sigma ** 2))
            m = g[:, None] * g[None, :]
    else:
        raise ValueError(mode)
    min_non_zero = max(m[m != 0].min().item(), 1e-3)
    return torch.clamp(m.to(torch.float32), min=min_non_zero)


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor, overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, gaussian_variant: str = "A",
                             return_window_logits: bool = False):
    """inputs [B,C,H,W] float32 → blended predictor output [B,C_out,H,W]."""
    roi = int(roi_size[0]) if not isinstance(roi_size, int) else int(roi_size)
    assert isinstance(roi_size, int) or roi_size[0] == roi_size[1], "square ROI (the reference uses 256x256)"
    B, _, H0, W0 = inputs.shape
    pads = []
    for D in (H0, W0):
        diff = max(roi - D, 0)
        half = diff // 2
        pads.append((half, diff - half))
    x = torch.nn.functional.pad(inputs, (pads[1][0], pads[1][1], pads[0][0], pads[0][1]), mode="constant", value=0.0)
    H, W = x.shape[2:]
    ys, xs = window_starts(H, roi, overlap), window_starts(W, roi, overlap)
    slices = [(y, xx) for y in ys for xx in xs]
    num_win = len(slices)
    total = num_win * B
    w = importance_map(roi, mode, sigma_scale, gaussian_variant)
    out = cnt = None
    all_logits = []
    for g in range(0, total, sw_batch_size):
        rng = range(g, min(g + sw_batch_size, total))
        data = torch.cat([x[idx // num_win: idx // num_win + 1, :, slices[idx % num_win][0]:slices[idx % num_win][0] + roi,
                            slices[idx % num_win][1]:slices[idx % num_win][1] + roi] for idx in rng])
        seg = predictor(data)
        if return_window_logits:
            all_logits.append(seg.clone())
        if out is None:
            out = torch.zeros((B, seg.shape[1], H, W), dtype=torch.float32)
            cnt = torch.zeros((B, 1, H, W), dtype=torch.float32)   # per-image weight sum
        for idx in rng:
            b = idx // num_win
            y, xx = slices[idx % num_win]
            out[b, :, y:y + roi, xx:xx + roi] += w * seg[idx - g]
            cnt[b, 0, y:y + roi, xx:xx + roi] += w
    out = out / cnt
    out = out[:, :, pads[0][0]:pads[0][0] + H0, pads[1][0]:pads[1][0] + W0]
    if return_window_logits:
        return out, torch.cat(all_logits)
    return out


def blend_from_window_logits(win_logits: torch.Tensor, B: int, H0: int, W0: int, roi: int, overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, gaussian_variant: str = "A") -> torch.Tensor:
    """Steps (5)-(6) of the algorithm on precomputed window logits [B*n_win, C, roi, roi]
    (tier-1 parity feeds identical logits to the oracle and the kernel)."""
    it = iter(range(win_logits.shape[0]))
    return sliding_window_inference(torch.zeros((B, 1, H0, W0)), (roi, roi), 1,
                                    lambda d: win_logits[next(it)][None], overlap, mode, sigma_scale, gaussian_variant)


def scale_to_unit(pre_img_u8: np.ndarray) -> torch.Tensor:
    """baseline/predict.py:107-108: u8 HWC / global max in float64, to float32 NCHW."""
    with np.errstate(divide="ignore", invalid="ignore"):
        a = pre_img_u8 / np.max(pre_img_u8)
    return torch.from_numpy(np.expand_dims(a, 0)).permute(0, 3, 1, 2).type(torch.FloatTensor)


def foreground_mask(blended_logits: torch.Tensor) -> np.ndarray:
    """baseline/predict.py:110-113: softmax over classes, class-1 probability > 0.5."""
    p = torch.nn.functional.softmax(blended_logits, dim=1)
    return p[0, 1].cpu().numpy() > 0.5
