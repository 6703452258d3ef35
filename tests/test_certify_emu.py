"""CPU tier for the certified-large pre-pass (certify.cu) and the chain built on it.

The pre-pass may certify as much or as little as it likes — the chain's results must not depend on it — but what it
certifies must be right: every certified pixel lies in a 4-connected component of at least `threshold` pixels of its
polarity.  On cell-like masks it must also be useful (most of the large components' pixels certified), otherwise the tile
kernel is back to labelling everything."""
import os

import numpy as np
import pytest
import scipy.ndimage as ndi
from hypothesis import given, settings, strategies as st

import emu_lib
from cases import all_cases
from cellseg_b200 import synth
from oracle import morph

C4 = ndi.generate_binary_structure(2, 1)
CASES = all_cases()


def large_pixels(m, thr):
    cc, n = ndi.label(m, C4)
    sizes = np.bincount(cc.ravel(), minlength=n + 1)
    big = sizes >= thr
    big[0] = False
    return big[cc]


def unet_masks(golden_dir):
    return np.unpackbits(np.load(os.path.join(golden_dir, "unet_masks_4x1024.npz"))["packed"]).reshape(4, 1024, 1024).astype(bool)


@pytest.mark.parametrize("name,mask", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("thr", [(64, 16), (5, 200), (300, 2), (1, 1), (17, 100)])
def test_certified_is_subset_of_large(name, mask, thr):
    bg, fg = emu_lib.certify_large(mask, *thr)
    m = np.asarray(mask, bool)
    assert not (bg & m).any() and not (fg & ~m).any()
    if thr[0] > 1:
        assert not (bg & ~large_pixels(~m, thr[0])).any()
    else:
        assert not bg.any()
    if thr[1] > 1:
        assert not (fg & ~large_pixels(m, thr[1])).any()
    else:
        assert not fg.any()


def test_coverage_on_cell_like_masks(golden_dir):
    for m in (synth.blob_mask(512, 640, 1), synth.lattice_cells(512, 640, seed=2, pitch=90, radius=14) > 0):
        bg, fg = emu_lib.certify_large(m)
        assert bg.sum() >= 0.93 * large_pixels(~m, 64).sum()
        assert fg.sum() >= 0.90 * large_pixels(m, 16).sum()
    # salt-like masks of the bench's random-init predictor: no solid 8x8 background block anywhere, the single-seed floods
    # must still certify a good part of the percolating background (the tile kernel's closure rounds take it from there)
    m = unet_masks(golden_dir)[0][:256, :384]
    bg, fg = emu_lib.certify_large(m)
    assert bg.sum() >= 0.5 * large_pixels(~m, 64).sum()
    assert not (bg & ~large_pixels(~m, 64)).any() and not (fg & ~large_pixels(m, 16)).any()


def mixed_mask(h, w, seed):
    """blobs + salt + pepper + thin bridges, with small holes / specks placed on tile borders (64-row / 128-column lines)."""
    rng = np.random.default_rng(seed)
    m = synth.blob_mask(h, w, seed, sigma=float(rng.uniform(2.5, 8)))
    m ^= rng.random((h, w)) < rng.choice([0.0, 0.01, 0.08])          # salt and pepper
    for _ in range(12):                                             # small features straddling tile borders
        y = int(rng.choice([63, 64, 127, 128])) if h > 130 else int(rng.integers(0, h))
        x = int(rng.choice([127, 128, 255, 256])) if w > 260 else int(rng.integers(0, w))
        dy, dx = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        y0, x0 = max(0, y - dy // 2), max(0, x - dx // 2)
        m[y0:y0 + dy, x0:x0 + dx] = rng.random() < 0.5
    return m


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 200), st.integers(1, 400), st.integers(0, 2**31 - 1), st.sampled_from([(64, 16), (64, 16), (10, 40), (257, 3), (30, 0), (0, 25)]))
def test_chain_on_mixed_masks(h, w, seed, thr):
    m = mixed_mask(h, w, seed)
    lab, n = emu_lib.postprocess(m, thr[0], thr[1], 8)
    assert np.array_equal(lab, morph.postprocess_mask(m, thr[0], thr[1], 2))


def test_chain_on_predictor_masks(golden_dir):
    m = unet_masks(golden_dir)
    for i, (y, x) in enumerate(((0, 0), (300, 200), (512, 100), (700, 600))):
        sub = m[i][y:y + 300, x:x + 390]
        lab, _ = emu_lib.postprocess(sub, 64, 16, 8)
        assert np.array_equal(lab, morph.postprocess_mask(sub))
