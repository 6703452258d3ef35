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
    one, many = [], []
    for _ in range(3):                                       # best of three: the CI box is shared, a single timing can be unlucky
        t0 = time.perf_counter(); tiffio.write_tiff(str(tmp_path / "a.tiff"), lab, rows_per_strip=4096); t1 = time.perf_counter()
        tiffio.write_tiff(str(tmp_path / "b.tiff"), lab); t2 = time.perf_counter()
        one.append(t1 - t0); many.append(t2 - t1)
    assert np.array_equal(tiffio.imread(str(tmp_path / "b.tiff")), lab)
    assert min(many) < min(one)


@pytest.mark.parametrize("shape,dtype,compress", [((70, 90, 3), np.uint16, True), ((70, 90, 3), np.uint16, False), ((40, 33, 4), np.uint16, True),
                                                  ((64, 64), np.int32, True), ((31, 17, 2), np.uint8, True), ((20, 30), np.float32, True)])
def test_builtin_reader_reads_what_pillow_cannot(tmp_path, shape, dtype, compress):
    """16-bit multi-channel TIFFs (fluorescence exports) are not decodable by Pillow; imread falls back to read_tiff."""
    from cellseg_b200 import tiffio
    rng = np.random.default_rng(2)
    a = (rng.random(shape) * 1000).astype(dtype)
    p = str(tmp_path / "m.tif")
    tiffio.write_tiff(p, a, compress=compress, rows_per_strip=7)
    assert np.array_equal(tiffio.read_tiff(p), a) and tiffio.read_tiff(p).dtype == a.dtype
    assert np.array_equal(tiffio.imread(p), a)


def test_builtin_reader_against_opencv_written_files(tmp_path):
    """Files from an independent writer (libtiff through OpenCV): LZW is refused with a clear message, deflate / none decode."""
    import cv2
    from cellseg_b200 import tiffio
    rng = np.random.default_rng(3)
    a = rng.integers(0, 60000, (50, 61, 3)).astype(np.uint16)
    for comp, ok in ((1, True), (8, True), (32946, True), (5, False)):
        p = str(tmp_path / f"c{comp}.tif")
        assert cv2.imwrite(p, a[..., ::-1], [cv2.IMWRITE_TIFF_COMPRESSION, comp])
        if ok:
            assert np.array_equal(tiffio.read_tiff(p), a), comp
        else:
            with pytest.raises(ValueError, match="not supported"):
                tiffio.read_tiff(p)


def test_palette_png_comes_back_as_rgb(tmp_path):
    from PIL import Image
    from cellseg_b200 import tiffio
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 4, (40, 50, 3)).astype(np.uint8) * 60
    pal = Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=64)
    p = str(tmp_path / "p.png")
    pal.save(p)
    limit = Image.MAX_IMAGE_PIXELS
    got = tiffio.imread(p)
    assert got.shape == (40, 50, 3) and np.array_equal(got, np.array(pal.convert("RGB")))
    assert Image.MAX_IMAGE_PIXELS == limit                   # the bomb guard is only lifted for the duration of the read
