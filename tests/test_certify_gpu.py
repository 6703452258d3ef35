"""GPU tier of tests/test_certify_emu.py: the certified-large pre-pass through the C ABI on a B200, and the chain on masks
built to stress it (features on tile borders, salt masks of the bench's predictor, odd thresholds, full-size batches)."""
import os

import numpy as np
import pytest

from cases import all_cases
from cellseg_b200 import synth
from oracle import morph
from test_certify_emu import large_pixels, mixed_mask, unet_masks

pytestmark = pytest.mark.gpu
CASES = all_cases()


@pytest.fixture(scope="module")
def pp():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    from cellseg_b200 import postprocess
    return postprocess


def _dev(mask):
    import torch
    return torch.from_numpy(np.ascontiguousarray(mask)).cuda()


@pytest.mark.parametrize("name,mask", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("thr", [(64, 16), (5, 200), (300, 2)])
def test_certified_is_subset_of_large(pp, name, mask, thr):
    bg, fg = (t.cpu().numpy() for t in pp.certified_large(_dev(mask), *thr))
    m = np.asarray(mask, bool)
    assert not (bg & ~large_pixels(~m, thr[0])).any()
    assert not (fg & ~large_pixels(m, thr[1])).any()


def test_coverage(pp, golden_dir):
    m = synth.blob_mask(2048, 2048, 1)
    bg, fg = (t.cpu().numpy() for t in pp.certified_large(_dev(m)))
    assert bg.sum() >= 0.93 * large_pixels(~m, 64).sum() and fg.sum() >= 0.90 * large_pixels(m, 16).sum()
    m = unet_masks(golden_dir)[1]
    bg, fg = (t.cpu().numpy() for t in pp.certified_large(_dev(m)))
    assert bg.sum() >= 0.5 * large_pixels(~m, 64).sum()
    assert not (bg & ~large_pixels(~m, 64)).any() and not (fg & ~large_pixels(m, 16)).any()


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("thr", [(64, 16), (10, 40), (257, 3), (30, 0), (0, 25)])
def test_chain_on_mixed_masks(pp, seed, thr):
    rng = np.random.default_rng(seed)
    m = mixed_mask(int(rng.integers(1, 700)), int(rng.integers(1, 900)), seed)
    lab, _ = pp.postprocess_device(_dev(m), thr[0], thr[1], 2)
    assert np.array_equal(lab.cpu().numpy(), morph.postprocess_mask(m, thr[0], thr[1], 2))


def test_chain_on_predictor_masks_full_batch(pp, golden_dir):
    m = unet_masks(golden_dir)
    lab, n = pp.postprocess_device(_dev(m))
    lab = lab.cpu().numpy()
    for i in range(4):
        ref = morph.postprocess_mask(m[i])
        assert np.array_equal(lab[i], ref) and int(n[i]) == ref.max()
