"""GPU tier for the F1-metric path: libcellseg_sm100.so (through the C ABI, via cellseg_b200.compute_metric)
against the golden vectors produced by the REFERENCE's own baseline/compute_metric.py
(tests/golden/metric_*.npz, oracle/make_golden.py) and against the oracle at BASELINE config-4 size."""
import json
import os

import numpy as np
import pytest

from oracle import metric as om
from oracle import morph

pytestmark = pytest.mark.gpu
FIX = ("cell_00225", "cell_00528", "cell_00849")


@pytest.fixture(scope="module")
def cm():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    from cellseg_b200 import compute_metric
    return compute_metric


@pytest.fixture(scope="module")
def fixtures(golden_dir):
    return np.load(os.path.join(golden_dir, "metric_fixtures.npz"))


@pytest.fixture(scope="module")
def synthetic(golden_dir):
    return np.load(os.path.join(golden_dir, "metric_synthetic.npz"))


def _fixture_label(golden_dir, name):
    from PIL import Image
    return np.array(Image.open(os.path.join(golden_dir, "ref_data", name + "_label.tiff"))).astype(np.int32)


@pytest.mark.parametrize("name", FIX)
def test_reference_fixtures(cm, golden_dir, fixtures, name):
    g = _fixture_label(golden_dir, name)
    work = g.copy()
    rb = cm.remove_boundary_cells(work)
    assert rb.dtype == np.int32
    assert np.array_equal(rb, fixtures[name + "_rb"]) and np.array_equal(work, fixtures[name + "_rb_mutated"])
    assert tuple(cm.eval_tp_fp_fn(rb, rb, 0.5)) == tuple(fixtures[name + "_self"])
    rolled = cm.remove_boundary_cells(np.roll(g, 3, axis=1).copy())
    for th in (0.5, 0.7, 0.9):
        assert tuple(cm.eval_tp_fp_fn(rb, rolled, th)) == tuple(fixtures[f"{name}_roll3_th{th}"])
    assert cm.dice(g > 0, np.roll(g, 3, axis=1) > 0) == float(fixtures[name + "_dice_roll3"])


def test_synthetic_pairs(cm, synthetic):
    for name in synthetic["names"]:
        gt, seg = synthetic[name + "_gt"], synthetic[name + "_seg"]
        assert cm.dice(gt > 0, seg > 0) == float(synthetic[name + "_dice"])
        a, b = gt.astype(np.int32).copy(), seg.astype(np.int32).copy()
        ra, rb = cm.remove_boundary_cells(a), cm.remove_boundary_cells(b)
        assert np.array_equal(ra, synthetic[name + "_gt_rb"]) and np.array_equal(rb, synthetic[name + "_seg_rb"])
        assert np.array_equal(a, synthetic[name + "_gt_mut"]) and np.array_equal(b, synthetic[name + "_seg_mut"])
        for th in (0.3, 0.5, 0.7):
            assert tuple(cm.eval_tp_fp_fn(ra, rb, th)) == tuple(synthetic[f"{name}_tpfpfn_{th}"]), (name, th)
        assert tuple(cm.eval_tp_fp_fn(gt, seg, 0.5)) == tuple(synthetic[f"{name}_raw_tpfpfn_0.5"])
        if name + "_overlap" in synthetic.files:
            ov = cm._label_overlap(ra, rb)
            assert ov.dtype == np.uint and np.array_equal(ov, synthetic[name + "_overlap"])
            iou = cm._intersection_over_union(ra, rb)
            assert np.array_equal(iou, synthetic[name + "_iou"])              # float64, bit-exact
            for th in (0.3, 0.5, 0.7):
                assert cm._true_positive(iou[1:, 1:], th) == om.true_positive(iou[1:, 1:], th)


def test_cli_rows_including_25mpx_patch_branch(cm, golden_dir):
    """Every CSV row the reference's main() produced, incl. the 5003x5001 image that takes the 2000² patch branch."""
    from cellseg_b200 import synth
    spec = json.load(open(os.path.join(golden_dir, "metric_cli.json")))
    for th, rows in spec["csv"].items():
        for row in rows[1:]:
            key = row[0].split("_label")[0]
            H, W = spec["specs"][key]
            r = cm.evaluate_pair(synth.lattice_cells(H, W, seed=spec["gt_seed"]), synth.lattice_cells(H, W, seed=spec["seg_seed"]), float(th))
            assert [r["true_num"], r["pred_num"], r["tp"], r["fn"], r["fp"]] == [int(v) for v in row[1:6]], (key, th)
            assert float(np.round(r["dice"], 4)) == float(row[8])


def test_config4_full_size_against_oracle(cm):
    """BASELINE config 4: ~20k cells per 4096² map.  SURVEY.md §8d known answer: tp = 19 785 at th = 0.5."""
    from cellseg_b200 import synth
    gt, seg = synth.lattice_cells(4096, 4096, seed=0), synth.lattice_cells(4096, 4096, seed=1)
    a, b = gt.copy(), seg.copy()
    ra, rb = cm.remove_boundary_cells(a), cm.remove_boundary_cells(b)
    oa, ob = gt.copy(), seg.copy()
    assert np.array_equal(ra, om.remove_boundary_cells(oa)) and np.array_equal(rb, om.remove_boundary_cells(ob))
    assert np.array_equal(a, oa) and np.array_equal(b, ob)
    assert (ra.max(), rb.max()) == (19826, 19827)
    tp, fp, fn = cm.eval_tp_fp_fn(ra, rb, 0.5)
    assert (tp, fp, fn) == (19785, 19827 - 19785, 19826 - 19785)
    # sparse triples against a bincount of the pair codes (size-independent check of the histogram)
    import torch
    g, p = torch.from_numpy(ra).cuda(), torch.from_numpy(rb).cuda()
    sp = cm._sparse_overlap_dev(g, p, int(ra.max()), int(rb.max()))
    n = sp["n"]
    keys = sp["pg"][:n].cpu().numpy().astype(np.int64) * (int(rb.max()) + 1) + sp["pp"][:n].cpu().numpy()
    cnt = sp["pc"][:n].cpu().numpy().view(np.uint32)
    order = np.argsort(keys)
    codes, ref_cnt = np.unique(ra.astype(np.int64) * (int(rb.max()) + 1) + rb, return_counts=True)
    assert np.array_equal(keys[order], codes) and np.array_equal(cnt[order], ref_cnt)
    assert int(cnt.sum()) == 4096 * 4096                                       # checksum of counts
    r = cm.evaluate_pair(gt, seg, 0.5)
    o = om.evaluate_pair(gt.copy(), seg.copy(), 0.5)
    assert [r[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")] == [o[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")]


def test_quirks_and_edges(cm):
    for arr in (np.full((4, 7), 5, np.int32), np.array([[3, 3, 9], [9, 9, 9]], np.int32), np.full((3, 3), 0, np.int32),
                np.array([[0, 2, 2, 4]], np.int32)):
        a, b = arr.copy(), arr.copy()
        assert np.array_equal(cm.remove_boundary_cells(a), om.remove_boundary_cells(b)), arr
        assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    lab = rng.choice(np.array([0, 3, 4, 9, 100, 1000, 65000], np.int32), size=(50, 70))
    assert np.array_equal(cm.relabel_sequential(lab), morph.relabel_sequential(lab)[0])
    z = np.zeros((30, 30), np.int32)
    assert np.array_equal(cm.relabel_sequential(z), z)
    assert cm.eval_tp_fp_fn(lab, z, 0.5) == (0, 0, 0)
    assert cm.dice(z > 0, z > 0) == 1.0 and cm.dice(z > 0, lab > 0) == 0.0


def test_random_label_noise_stresses_hash(cm):
    """Per-pixel random labels: every pixel starts a new run and the shared hash overflows into the global table."""
    rng = np.random.default_rng(5)
    gt = rng.integers(0, 3000, (512, 768)).astype(np.int32)
    seg = rng.integers(0, 2500, (512, 768)).astype(np.int32)
    ov = cm._label_overlap(gt, seg)
    assert np.array_equal(ov, om.label_overlap(gt, seg))
