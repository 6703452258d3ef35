"""The tifffile-free TIFF writer (predict.py:114's `tif.imwrite(..., compression='zlib')`): files must read back
identically through an independent decoder (Pillow / OpenCV), strips deflated in parallel or not."""
import os
import time

import numpy as np
import pytest


@pytest.mark.parametrize("shape,dtype,compress", [((300, 400), np.int32, True), ((1, 1), np.int32, True), ((257, 129), np.uint8, True),
                                                  ((200, 200, 3), np.uint16, False), ((64, 70, 3), np.uint8, True), ((5000, 37), np.int32, True),
                                                  ((90, 80), np.uint16, True), ((33, 44), np.float32, False)])
def test_roundtrip_through_pillow(tmp_path, shape, dtype, compress):
    from PIL import Image
    from cellseg_b200 import tiffio
    rng = np.random.default_rng(1)
    a = (rng.integers(0, 256 if dtype == np.uint8 else 50000, shape) if dtype != np.float32 else rng.random(shape) * 100).astype(dtype)
    p = str(tmp_path / "x.tiff")
    tiffio.write_tiff(p, a, compress=compress)
    if a.ndim == 3 and dtype == np.uint16:
        import cv2                                           # Pillow has no 16-bit RGB mode; OpenCV's libtiff does
        back = cv2.imread(p, cv2.IMREAD_UNCHANGED)[..., ::-1]
    else:
        back = np.array(Image.open(p))
    assert back.dtype == a.dtype and np.array_equal(back, a)


def test_label_tiff_matches_reference_fixture_format(tmp_path, golden_dir):
    """Same compression / layout tags as the reference's own label files (Adobe deflate, one sample per pixel)."""
    from PIL import Image
    from cellseg_b200 import tiffio
    ref = Image.open(os.path.join(golden_dir, "ref_data", "cell_00849_label.tiff"))
    lab = np.array(ref).astype(np.int32)
    p = str(tmp_path / "cell_00849_label.tiff")
    tiffio.imwrite_labels(p, lab)
    mine = Image.open(p)
    assert np.array_equal(np.array(mine), lab) and np.array(mine).dtype == np.int32
    for tag in (259, 277):                                   # Compression (8 = Adobe deflate), SamplesPerPixel
        assert mine.tag_v2[tag] == ref.tag_v2[tag], tag
    assert tiffio.imread(p).dtype == np.int32


def test_parallel_strips_are_faster_than_one(tmp_path):
    from cellseg_b200 import tiffio
    if (os.cpu_count() or 1) < 4:
        pytest.skip("needs a few cores")
    yy, xx = np.mgrid[0:4096, 0:4096]
    lab = ((yy // 29) * 142 + xx // 29).astype(np.int32) * ((yy % 29 < 22) & (xx % 29 < 22))
    t0 = time.perf_counter(); tiffio.write_tiff(str(tmp_path / "a.tiff"), lab, rows_per_strip=4096); t1 = time.perf_counter()
    tiffio.write_tiff(str(tmp_path / "b.tiff"), lab); t2 = time.perf_counter()
    assert np.array_equal(tiffio.imread(str(tmp_path / "b.tiff")), lab)
    assert (t2 - t1) < (t1 - t0)
