"""Seeded synthetic inputs for tests and bench (no datasets offline).

Everything here is numpy on the host and deterministic in `seed`.
"""
from __future__ import annotations

import numpy as np


def _smooth(a: np.ndarray, sigma: float) -> np.ndarray:
    import scipy.ndimage as ndi
    return ndi.gaussian_filter(a, sigma, mode="wrap")


def blob_mask(H: int, W: int, seed: int = 0, sigma: float = 6.0, thresh: float = 0.5) -> np.ndarray:
    """Gaussian-filtered N(0,1) noise thresholded at +thresh std → ≈31 % foreground blobs
    (the config-3 post-processing workload)."""
    rng = np.random.default_rng(seed)
    f = _smooth(rng.standard_normal((H, W)).astype(np.float32), sigma)
    return f > thresh * f.std()


def noise_mask(H: int, W: int, p: float, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).random((H, W)) < p


def spiral_mask(H: int, W: int, gap: int = 2) -> np.ndarray:
    """One long 1-pixel-wide rectangular spiral: worst case for union-find depth."""
    m = np.zeros((H, W), dtype=bool)
    y = x = 0
    top, bottom, left, right = 0, H - 1, 0, W - 1
    m[0, 0] = True
    while True:
        if x >= right:
            break
        m[y, x:right + 1] = True; x = right; top = y + gap
        if y >= bottom:
            break
        m[y:bottom + 1, x] = True; y = bottom; right = x - gap
        if x <= left:
            break
        m[y, left:x + 1] = True; x = left; bottom = y - gap
        if y <= top:
            break
        m[top:y + 1, x] = True; y = top; left = x + gap
    return m


def comb_mask(H: int, W: int, vertical: bool = True) -> np.ndarray:
    m = np.zeros((H, W), dtype=bool)
    if vertical:
        m[:, ::2] = True; m[0, :] = True
    else:
        m[::2, :] = True; m[:, 0] = True
    return m


def checkerboard(H: int, W: int) -> np.ndarray:
    yy, xx = np.mgrid[0:H, 0:W]
    return ((yy + xx) & 1).astype(bool)


def modality_image(H: int, W: int, seed: int = 0, modality: int | None = None) -> np.ndarray:
    """Modality-like synthetic microscopy image (brightfield / fluorescence / phase / DIC).
    Returns u8[H,W,3] except phase-contrast which is 2-D u8[H,W] (exercises predict.py:94-95)."""
    rng = np.random.default_rng(seed)
    modality = seed % 4 if modality is None else modality
    n_cells = max(4, H * W // 4000)
    cy = rng.integers(0, H, n_cells); cx = rng.integers(0, W, n_cells)
    ry = rng.uniform(6, 14, n_cells); rx = rng.uniform(6, 14, n_cells)
    cells = np.zeros((H, W), dtype=np.float32)
    for y, x, a, b in zip(cy, cx, ry, rx):
        y0, y1 = max(0, int(y - 3 * a)), min(H, int(y + 3 * a) + 1)
        x0, x1 = max(0, int(x - 3 * b)), min(W, int(x + 3 * b) + 1)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        g = np.exp(-(((yy - y) / a) ** 2 + ((xx - x) / b) ** 2))
        cells[y0:y1, x0:x1] = np.maximum(cells[y0:y1, x0:x1], g)
    if modality == 0:  # brightfield: bright background, darker cells, slight colour cast
        img = np.stack([200 + rng.normal(0, 8, (H, W)) - 60 * cells * c for c in (1.0, 0.9, 0.8)], -1)
        return np.clip(img, 0, 255).astype(np.uint8)
    if modality == 1:  # fluorescence: dark Poisson background, signal in one channel only
        ch = rng.poisson(6, (H, W)).astype(np.float32) + 220 * cells
        img = np.zeros((H, W, 3), dtype=np.float32); img[..., 1] = ch
        return np.clip(img, 0, 255).astype(np.uint8)
    if modality == 2:  # phase contrast: single channel, dark cells with bright halo
        halo = np.clip(_smooth(cells, 2.0) - cells, 0, None)
        img = 128 + rng.normal(0, 10, (H, W)) - 70 * cells + 160 * halo
        return np.clip(img, 0, 255).astype(np.uint8)
    # DIC: directional derivative of the blobs
    d = np.gradient(cells, axis=0) + np.gradient(cells, axis=1)
    img = 128 + rng.normal(0, 6, (H, W)) + 400 * d
    return np.repeat(np.clip(img, 0, 255).astype(np.uint8)[..., None], 3, -1)


def lattice_cells(H: int, W: int, seed: int = 0, pitch: int = 29, radius: int = 11, jitter: int = 2) -> np.ndarray:
    """Instance map of disks on a jittered lattice (config 4: ~20k cells per 4096²). int32, ids 1..n."""
    rng = np.random.default_rng(seed)
    lab = np.zeros((H, W), dtype=np.int32)
    yy, xx = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    disk = (yy * yy + xx * xx) <= radius * radius
    k = 0
    for y in range(pitch // 2, H, pitch):
        for x in range(pitch // 2, W, pitch):
            k += 1
            jy, jx = rng.integers(-jitter, jitter + 1, 2)
            y0, x0 = y + jy - radius, x + jx - radius
            ys0, xs0 = max(0, y0), max(0, x0)
            ys1, xs1 = min(H, y0 + 2 * radius + 1), min(W, x0 + 2 * radius + 1)
            if ys1 <= ys0 or xs1 <= xs0:
                continue
            sub = disk[ys0 - y0:ys1 - y0, xs0 - x0:xs1 - x0]
            view = lab[ys0:ys1, xs0:xs1]
            view[sub & (view == 0)] = k
    return lab
