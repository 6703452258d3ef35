"""Three-class training target (data/pre_process_3class.py:31-56): oracle self-checks on the CPU, the kernels on the emulator
(CPU tier) and through the product module on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from oracle import interior as oi
from oracle import morph


def _cases():
    from cellseg_b200 import synth
    rng = np.random.default_rng(0)
    c = [("lattice", synth.lattice_cells(150, 210, seed=1)),
         ("lattice_touching", synth.lattice_cells(130, 260, seed=2, pitch=20, radius=11, jitter=1)),     # cells overlap: shared boundaries
         ("noise_labels", (rng.integers(0, 4, (70, 90)) * rng.integers(0, 2, (70, 90))).astype(np.int32)),
         ("edge_cells", np.pad(np.full((20, 30), 7, np.int32), 0)),                                      # one cell filling the image: no boundary at the edge
         ("empty", np.zeros((33, 65), np.int32)), ("one_px", np.eye(1, dtype=np.int32))]
    big = synth.lattice_cells(90, 90, seed=3).astype(np.int64)
    big[big > 0] += 32760                                                                               # ids around the int16 wrap (reference casts to int16)
    c.append(("int16_wrap", big.astype(np.int32)))
    blobs = morph.label(synth.blob_mask(140, 200, 4), 2)
    c.append(("blobs", blobs))
    return c


@pytest.mark.parametrize("name,lab", _cases())
def test_oracle_boundary_definitions_agree(name, lab):
    l16 = lab.astype(np.int16)
    assert np.array_equal(oi.find_boundaries_inner(l16), oi.find_boundaries_inner_direct(l16))
    out = oi.create_interior_map(l16)
    assert out.dtype == np.uint8 and set(np.unique(out)) <= {0, 1, 2}
    assert not np.any((out > 0) & (l16 <= 0) & (out == 1))            # interior only inside cells


def test_fixture_label_statistics(golden_dir):
    """The reference's own sample label (data/Train_Labeled/labels/cell_00849_label.tiff, 187 cells): every cell keeps an id, boundary
    pixels exist for every cell, and interior components smaller than 16 pixels are gone."""
    from PIL import Image
    import scipy.ndimage as ndi
    lab = np.array(Image.open(os.path.join(golden_dir, "ref_data", "cell_00849_label.tiff"))).astype(np.int16)
    out = oi.create_interior_map(lab)
    assert (out[lab == 0] != 1).all()
    cc, n = ndi.label(out == 1)
    assert n > 0 and np.bincount(cc.ravel())[1:].min() >= 16
    assert all((out[lab == i] == 2).any() for i in np.unique(lab)[1:])


@pytest.mark.parametrize("name,lab", _cases())
def test_kernels_on_emulator_bit_exact(name, lab):
    import emu_lib
    from emu_lib import ptr, check
    L = emu_lib.lib()
    H, W = lab.shape
    inst = np.ascontiguousarray(lab, np.int32)
    b = np.zeros((H, W), np.uint8); it = np.zeros((H, W), np.uint8)
    check(L.cs_interior_boundary(ptr(inst), 1, H, W, ptr(b), ptr(it), None))
    l16 = inst.astype(np.int16)
    ref_b = oi.find_boundaries_inner(l16)
    import scipy.ndimage as ndi
    ref_b = ndi.binary_dilation(ref_b, structure=ndi.generate_binary_structure(2, 1))
    assert np.array_equal(b.astype(bool), ref_b)
    assert np.array_equal(it.astype(bool), ~ref_b & (l16 > 0))
    kept = emu_lib.remove_small(it, 16, 0)
    out = np.zeros((H, W), np.uint8)
    kept8 = np.ascontiguousarray(kept, np.uint8)
    check(L.cs_interior_combine(ptr(b), ptr(kept8), H * W, ptr(out), None))
    assert np.array_equal(out, oi.create_interior_map(l16))


@pytest.mark.gpu
@pytest.mark.parametrize("name,lab", _cases())
def test_create_interior_map_gpu_bit_exact(name, lab):
    import torch
    from cellseg_b200 import pre_process_3class as pp
    l16 = lab.astype(np.int16)
    ref = oi.create_interior_map(l16)
    assert np.array_equal(pp.create_interior_map(l16), ref)
    got = pp.create_interior_map(torch.from_numpy(lab).cuda())
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.gpu
def test_cli_writes_reference_layout(tmp_path, golden_dir):
    import shutil
    from PIL import Image
    from cellseg_b200 import pre_process_3class as pp
    from oracle import pipeline as opl
    src = tmp_path / "Train_Labeled"
    (src / "images").mkdir(parents=True); (src / "labels").mkdir()
    ref = os.path.join(golden_dir, "ref_data")
    for f in ("cell_00849.png", "cell_00528.tif"):
        shutil.copy(os.path.join(ref, f), src / "images" / f)
        shutil.copy(os.path.join(ref, f.split(".")[0] + "_label.tiff"), src / "labels" / (f.split(".")[0] + "_label.tiff"))
    out = tmp_path / "pre"
    pp.main(["-i", str(src), "-o", str(out)])
    assert sorted(os.listdir(out / "images")) == ["cell_00528.png", "cell_00849.png"]
    assert sorted(os.listdir(out / "labels")) == ["cell_00528_label.png", "cell_00849_label.png"]
    for stem, ext in (("cell_00849", ".png"), ("cell_00528", ".tif")):
        from cellseg_b200 import tiffio
        img = tiffio.imread(str(src / "images" / (stem + ext)))
        assert np.array_equal(np.array(Image.open(out / "images" / (stem + ".png"))), opl.preprocess(img))
        gt = tiffio.imread(str(src / "labels" / (stem + "_label.tiff")))
        assert np.array_equal(np.array(Image.open(out / "labels" / (stem + "_label.png"))), oi.create_interior_map(gt.astype(np.int16)))
