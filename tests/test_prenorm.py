"""Pre-normalisation (predict.py:14-21, 94-104): host logic on the CPU (histogram → numpy-exact percentiles → LUT),
kernels on the GPU against the oracle's numpy restatement."""
import numpy as np
import pytest

from oracle import pipeline as opl


@pytest.mark.parametrize("dtype,hi", [(np.uint8, 255), (np.uint16, 65535), (np.uint16, 900)])
def test_percentiles_from_histogram_equal_numpy(dtype, hi):
    from cellseg_b200.predict import percentiles_from_histogram
    rng = np.random.default_rng(int(hi))
    for n in (1, 2, 3, 7, 100, 101, 9973, 250000):
        a = rng.integers(0, hi + 1, n).astype(dtype)
        a[rng.random(n) < 0.3] = 0
        bins = 256 if dtype == np.uint8 else 65536
        got = percentiles_from_histogram(np.bincount(a, minlength=bins))
        nz = a[np.nonzero(a)]
        if len(nz) == 0:
            assert got is None
        else:
            assert np.array_equal(got, np.percentile(nz, [1, 99]))           # float64, bit for bit


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_lut_reproduces_normalize_channel(dtype):
    from cellseg_b200.predict import prenorm_lut
    rng = np.random.default_rng(3)
    bins = 256 if dtype == np.uint8 else 65536
    for scale in (bins - 1, 40, 1):
        img = (rng.gamma(2.0, scale / 6.0, (120, 130, 3))).clip(0, bins - 1).astype(dtype)
        img[..., 1] = 0                                                       # channel without signal stays zero
        img[..., 2] = img[0, 0, 2] if scale == 1 else img[..., 2]            # constant channel → the `else` branch
        hist = np.stack([np.bincount(img[..., c].ravel(), minlength=bins) for c in range(3)])
        lut = prenorm_lut(hist, dtype)
        ref = opl.preprocess(img)
        got = np.stack([lut[c][img[..., c]] for c in range(3)], -1)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape,dtype", [((37, 53, 3), np.uint8), ((64, 64, 3), np.uint8), ((5, 3, 3), np.uint8), ((41, 67), np.uint8),
                                         ((33, 31, 4), np.uint8), ((29, 45, 3), np.uint16), ((30, 30), np.uint16)])
def test_prenorm_kernels_on_the_emulator(shape, dtype):
    """Histogram and table application (fiber-emulated build): interleaved RGB in 16-pixel groups with a per-pixel tail, gray, RGBA,
    uint16 — against numpy."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import emu_lib
    from emu_lib import ptr, check
    L = emu_lib.lib()
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    bins = 256 if dtype == np.uint8 else 65536
    img = np.ascontiguousarray(rng.integers(0, 40 if dtype == np.uint8 else 20000, shape).astype(dtype))  # u8: few values, many collisions; u16: both sides of the shared-memory range
    H, W = shape[:2]
    ch = 1 if img.ndim == 2 else shape[2]
    hist = np.full((3, bins), 7, np.uint32)
    check(L.cs_prenorm_hist(ptr(img), img.dtype.itemsize, H, W, ch, ptr(hist), None))
    flat = img.reshape(H * W, ch)
    for c in range(min(ch, 3)):
        assert np.array_equal(hist[c], np.bincount(flat[:, c], minlength=bins))
    lut = rng.integers(0, 256, (3, bins), dtype=np.uint8)
    out = np.full((H, W, 3), 9, np.uint8)
    check(L.cs_prenorm_apply(ptr(img), img.dtype.itemsize, H, W, ch, ptr(lut), ptr(out), None))
    ref = np.stack([lut[c][flat[:, 0 if ch == 1 else c]] for c in range(3)], -1).reshape(H, W, 3)
    assert np.array_equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["u8_rgb", "u16_gray", "u8_gray", "u8_rgba", "u16_rgb", "fixture_png", "fixture_tif"])
def test_device_prenorm_equals_oracle(case, golden_dir):
    import os
    from cellseg_b200 import synth, tiffio
    from cellseg_b200.predict import preprocess_image_device
    rng = np.random.default_rng(5)
    if case == "u8_rgb":
        img = synth.modality_image(700, 900, seed=4, modality=0)
    elif case == "u16_gray":
        img = (rng.gamma(2.0, 900.0, (513, 777))).clip(0, 65535).astype(np.uint16)
    elif case == "u8_gray":
        img = synth.modality_image(600, 500, seed=2, modality=2)
    elif case == "u8_rgba":
        img = rng.integers(0, 256, (300, 301, 4), dtype=np.uint8)
    elif case == "u16_rgb":
        img = rng.integers(0, 4096, (257, 300, 3)).astype(np.uint16); img[..., 0] = 0
    elif case == "fixture_png":
        img = tiffio.imread(os.path.join(golden_dir, "ref_data", "cell_00849.png"))
    else:
        img = tiffio.imread(os.path.join(golden_dir, "ref_data", "cell_00528.tif"))
    got = preprocess_image_device(img).cpu().numpy()
    assert np.array_equal(got, opl.preprocess(img))
