"""CPU tier: sliding-window kernels (fiber-emulated build) against the torch restatement of MONAI's
algorithm in oracle/sliding_window.py, plus the host-side window-grid logic."""
import numpy as np
import pytest
import torch

import emu_lib
from cellseg_b200 import sliding_window as sw, synth
from oracle import sliding_window as osw, morph


@pytest.mark.parametrize("D,roi", [(1024, 256), (2048, 256), (4096, 256), (512, 256), (10000, 256), (500, 256),
                                   (256, 256), (200, 256), (257, 256), (300, 64), (64, 64), (65, 64), (1, 64)])
def test_window_grid_matches_oracle(D, roi):
    p = sw.make_plan(D, D, roi, 0.25)
    Dp = max(D, roi)
    assert p.ys == osw.window_starts(Dp, roi, 0.25)
    assert p.pad_y == max(roi - D, 0) // 2
    cov = sw._cover_table(p.ys, roi, Dp)
    for c in {0, 1, Dp // 2, Dp - 1, min(Dp - 1, roi), min(Dp - 1, roi - 1)}:
        idx = [i for i, s in enumerate(p.ys) if s <= c < s + roi]
        assert cov[c] >> 8 == idx[0] and cov[c] & 255 == len(idx)


def test_survey_window_counts():
    # 1024² → 5x5, 2048² → 11², 4096² → 21², 512² → 3², 10000² → 52²; shifted last starts
    assert [sw.make_plan(d, d).ny for d in (1024, 2048, 4096, 512, 10000)] == [5, 11, 21, 3, 52]
    assert sw.make_plan(10000, 10000).ys[-3:] == [9408, 9600, 9744]
    assert sw.make_plan(512, 512).ys == [0, 192, 256]
    assert sw.make_plan(500, 500).ys == [0, 192, 244]


@pytest.mark.parametrize("maxv", [255, 200, 1, 0])
def test_scale_lut_and_unit_float(maxv):
    rng = np.random.default_rng(maxv)
    img = rng.integers(0, maxv + 1, (2, 37, 53, 3), dtype=np.uint8)
    img[0, 3, 4, 1] = maxv
    img[1] = np.minimum(img[1], maxv // 2)
    lut = emu_lib.scale_lut(img)
    out = emu_lib.to_unit_float(img, lut)
    for n in range(2):
        ref = osw.scale_to_unit(img[n]).numpy()[0]
        assert np.array_equal(out[n], ref, equal_nan=True)


@pytest.mark.parametrize("H,W,roi", [(100, 130, 64), (64, 64, 64), (50, 70, 64), (40, 200, 64), (129, 131, 64), (300, 280, 256)])
def test_gather_matches_monai_slicing(H, W, roi):
    rng = np.random.default_rng(H * W)
    img = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    lut = emu_lib.scale_lut(img)
    x = torch.cat([osw.scale_to_unit(img[n]) for n in range(2)])
    seen = []
    osw.sliding_window_inference(x, (roi, roi), 4, lambda d: (seen.append(d.clone()), d)[1])
    ref = torch.cat(seen).numpy()
    plan = sw.make_plan(H, W, roi, 0.25)
    assert ref.shape[0] == 2 * plan.n_win
    assert np.array_equal(emu_lib.gather(img, plan, 0, 2 * plan.n_win, lut), ref)
    assert np.array_equal(emu_lib.gather(x.numpy(), plan, 0, 2 * plan.n_win), ref)
    if plan.n_win > 2:
        assert np.array_equal(emu_lib.gather(img, plan, 3, 2, lut), ref[3:5])


@pytest.mark.parametrize("H,W,roi,mode,variant", [
    (100, 130, 64, "constant", "A"), (50, 70, 64, "constant", "A"), (125, 126, 64, "constant", "A"),
    (100, 130, 64, "gaussian", "A"), (129, 200, 64, "gaussian", "B"), (300, 280, 256, "gaussian", "A"),
    (260, 300, 256, "constant", "A")])
def test_blend_softmax_threshold(H, W, roi, mode, variant):
    N = 2
    plan = sw.make_plan(H, W, roi, 0.25)
    g = torch.Generator().manual_seed(H + W)
    wl = torch.randn((N * plan.n_win, 3, roi, roi), generator=g) * 2.0
    ref = osw.blend_from_window_logits(wl, N, H, W, roi, 0.25, mode, gaussian_variant=variant)
    imap = sw.importance_map(roi, mode, variant=variant)
    oimap = osw.importance_map(roi, mode, variant=variant)
    if imap is not None:
        assert torch.equal(imap, oimap)
    out = emu_lib.blend(wl.numpy(), N, plan, None if imap is None else imap.numpy(), packed=True)
    assert np.array_equal(out["blended"], ref.numpy())             # same op order ⇒ bit-exact blend
    p_ref = torch.softmax(ref, dim=1)[:, 1].numpy()
    assert np.allclose(out["p1"], p_ref, rtol=1e-5, atol=1e-7)     # north_star tolerance (libm exp differs by ulps)
    sure = np.abs(p_ref - 0.5) > 1e-6
    assert np.array_equal(out["mask"].astype(bool)[sure], (p_ref > 0.5)[sure])
    for n in range(N):                                                # packed mask → instance map == oracle on the same mask
        assert np.array_equal(out["labels"][n], morph.postprocess_mask(out["mask"][n].astype(bool)))


def test_image_smaller_than_roi_is_padded_like_monai():
    img = synth.modality_image(40, 50, seed=0)
    lut = emu_lib.scale_lut(img[None])
    plan = sw.make_plan(40, 50, 64, 0.25)
    assert (plan.pad_y, plan.pad_x, plan.n_win) == (12, 7, 1)
    win = emu_lib.gather(img[None], plan, 0, 1, lut)
    assert np.all(win[0, :, :12] == 0) and np.all(win[0, :, 52:] == 0) and np.all(win[0, :, :, :7] == 0)
    ref = osw.scale_to_unit(img).numpy()[0]
    assert np.array_equal(win[0, :, 12:52, 7:57], ref)
