"""CPU tier: the post-processing kernels (compiled for the fiber emulator, tests/emu) against the
oracle.  Checks kernel LOGIC bit-exactly; the same cases run on the real library in
test_postprocess_gpu.py."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import emu_lib
from cases import all_cases
from oracle import morph

CASES = all_cases()


@pytest.mark.parametrize("name,mask", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("conn", [8, 4])
def test_label_matches_scipy(name, mask, conn):
    lab, n = emu_lib.postprocess(mask, 0, 0, conn)
    ref = morph.label(mask, conn)
    assert lab.dtype == np.int32
    assert np.array_equal(lab, ref)
    assert n[0] == ref.max()


@pytest.mark.parametrize("name,mask", CASES, ids=[c[0] for c in CASES])
def test_chain_matches_oracle(name, mask):
    lab, n = emu_lib.postprocess(mask, 64, 16, 8)
    ref = morph.postprocess_mask(mask, 64, 16, 2)
    assert np.array_equal(lab, ref)
    assert n[0] == ref.max()


@pytest.mark.parametrize("name,mask", CASES[:12] + CASES[24:], ids=[c[0] for c in CASES[:12] + CASES[24:]])
@pytest.mark.parametrize("thr,invert", [(16, 0), (64, 1), (2, 0), (200, 1), (1000, 0)])
def test_remove_small(name, mask, thr, invert):
    out = emu_lib.remove_small(mask, thr, invert).astype(bool)
    ref = morph.remove_small_holes(mask, thr) if invert else morph.remove_small_objects(mask, thr)
    assert np.array_equal(out, ref)


def test_cv2_crosscheck_of_oracle():
    for name, mask in CASES[:8]:
        for conn in (4, 8):
            assert np.array_equal(morph.label(mask, conn), morph.label_cv2(mask, conn)), name


def test_batch_independent_images():
    rng = np.random.default_rng(5)
    masks = np.stack([rng.random((150, 270)) < p for p in (0.3, 0.5, 0.62, 0.0, 1.0)])
    lab, n = emu_lib.postprocess(masks, 64, 16, 8)
    for i in range(len(masks)):
        ref = morph.postprocess_mask(masks[i])
        assert np.array_equal(lab[i], ref)
        assert n[i] == ref.max()


def test_larger_blob_image():
    from cellseg_b200 import synth
    m = synth.blob_mask(1024, 1024, 3)
    lab, n = emu_lib.postprocess(m, 64, 16, 8)
    assert np.array_equal(lab, morph.postprocess_mask(m))


def test_properties():
    from cellseg_b200 import synth
    m = synth.noise_mask(300, 420, 0.58, 77)
    lab, n = emu_lib.postprocess(m, 64, 16, 8)
    ids = np.unique(lab)
    assert ids[0] == 0 and np.array_equal(ids[1:], np.arange(1, n[0] + 1))
    first = np.array([np.flatnonzero(lab.ravel() == k)[0] for k in range(1, n[0] + 1)])
    assert np.all(np.diff(first) > 0)                       # raster-monotone first pixels
    # idempotence: relabelling the foreground of the result changes nothing
    lab2, _ = emu_lib.postprocess(lab > 0, 64, 16, 8)
    assert np.array_equal(lab, lab2)
    sizes = np.bincount(lab.ravel())[1:]
    assert sizes.min() >= 1


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 140), st.integers(1, 280), st.floats(0.05, 0.95), st.integers(0, 2**31 - 1))
def test_random_shapes(h, w, p, seed):
    m = np.random.default_rng(seed).random((h, w)) < p
    lab, n = emu_lib.postprocess(m, 64, 16, 8)
    assert np.array_equal(lab, morph.postprocess_mask(m))
    lab4, _ = emu_lib.postprocess(m, 0, 0, 4)
    assert np.array_equal(lab4, morph.label(m, 4))


def test_argument_errors():
    L = emu_lib.lib()
    assert L.cs_postprocess(None, 1, 4, 4, 64, 16, 8, None, None, None, 0, None) < 0
    assert b"null" in L.cs_last_error()
    m = np.zeros((1, 8, 8), np.uint8); lab = np.zeros((1, 8, 8), np.int32)
    assert L.cs_postprocess(emu_lib.ptr(m), 1, 8, 8, 64, 16, 6, emu_lib.ptr(lab), None, None, 0, None) == -1
    assert L.cs_postprocess(emu_lib.ptr(m), 1, 8, 8, 64, 16, 8, emu_lib.ptr(lab), None, emu_lib.ptr(m), 8, None) == -2
    assert L.cs_postprocess_workspace_bytes(0, 8, 8) == 0
