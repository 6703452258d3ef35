"""CPU tier for the F1-metric path:
  1. the oracle restatement (oracle/metric.py) against golden vectors produced by the REFERENCE's own
     baseline/compute_metric.py (tests/golden/metric_*.npz, oracle/make_golden.py);
  2. the metric kernels (fiber-emulated build) and the host logic of cellseg_b200.compute_metric
     against the same vectors.  The product module is pointed at the emulated library by monkeypatching
     its device helpers — test-only; the package itself has no CPU path."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import emu_lib
from oracle import metric as om
from oracle import morph


@pytest.fixture(scope="module")
def fixtures(golden_dir):
    return np.load(os.path.join(golden_dir, "metric_fixtures.npz"))


@pytest.fixture(scope="module")
def synthetic(golden_dir):
    return np.load(os.path.join(golden_dir, "metric_synthetic.npz"))


def _fixture_label(golden_dir, name):
    from PIL import Image
    return np.array(Image.open(os.path.join(golden_dir, "ref_data", name + "_label.tiff"))).astype(np.int32)


FIX = ("cell_00225", "cell_00528", "cell_00849")


@pytest.fixture()
def cm(monkeypatch):
    """cellseg_b200.compute_metric wired to the emulated library and CPU tensors (test-only)."""
    from cellseg_b200 import capi, compute_metric
    monkeypatch.setattr(capi, "_lib", emu_lib.lib())
    monkeypatch.setattr(compute_metric, "_torch", lambda: torch)
    monkeypatch.setattr(compute_metric, "_stream", lambda t: None)
    monkeypatch.setattr(compute_metric, "_to_device_i32",
                        lambda a: a.to(torch.int32) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)))
    return compute_metric


# ---- 1. oracle vs the reference's own outputs -----------------------------------------------------
@pytest.mark.parametrize("name", FIX)
def test_oracle_on_reference_fixtures(golden_dir, fixtures, name):
    g = _fixture_label(golden_dir, name)
    assert len(np.unique(g)) - 1 == fixtures[name + "_nlabels"]
    work = g.copy()
    rb = om.remove_boundary_cells(work)
    assert np.array_equal(rb, fixtures[name + "_rb"]) and np.array_equal(work, fixtures[name + "_rb_mutated"])
    assert tuple(om.eval_tp_fp_fn(rb, rb, 0.5)) == tuple(fixtures[name + "_self"])
    rolled = om.remove_boundary_cells(np.roll(g, 3, axis=1).copy())
    for th in (0.5, 0.7, 0.9):
        assert tuple(om.eval_tp_fp_fn(rb, rolled, th)) == tuple(fixtures[f"{name}_roll3_th{th}"])
    assert om.dice(g > 0, np.roll(g, 3, axis=1) > 0) == float(fixtures[name + "_dice_roll3"])


def test_survey_known_answers(fixtures):
    # cell_00849 has 187 labels, 151 after boundary removal, eval(g, g) = (151, 0, 0)
    assert fixtures["cell_00849_nlabels"] == 187 and fixtures["cell_00849_rb"].max() == 151
    assert tuple(fixtures["cell_00849_self"]) == (151, 0, 0)


def test_oracle_on_synthetic_pairs(synthetic):
    for name in synthetic["names"]:
        gt, seg = synthetic[name + "_gt"], synthetic[name + "_seg"]
        assert om.dice(gt > 0, seg > 0) == float(synthetic[name + "_dice"])
        a, b = gt.astype(np.int32).copy(), seg.astype(np.int32).copy()
        ra, rb = om.remove_boundary_cells(a), om.remove_boundary_cells(b)
        assert np.array_equal(ra, synthetic[name + "_gt_rb"]) and np.array_equal(rb, synthetic[name + "_seg_rb"])
        assert np.array_equal(a, synthetic[name + "_gt_mut"]) and np.array_equal(b, synthetic[name + "_seg_mut"])
        for th in (0.3, 0.5, 0.7):
            assert tuple(om.eval_tp_fp_fn(ra, rb, th)) == tuple(synthetic[f"{name}_tpfpfn_{th}"]), (name, th)
        assert tuple(om.eval_tp_fp_fn(gt, seg, 0.5)) == tuple(synthetic[f"{name}_raw_tpfpfn_0.5"])
        if name + "_overlap" in synthetic.files:
            assert np.array_equal(om.label_overlap(ra, rb), synthetic[name + "_overlap"])
            assert np.array_equal(om.intersection_over_union(ra, rb), synthetic[name + "_iou"])


def test_oracle_cli_rows(golden_dir):
    from cellseg_b200 import synth
    spec = json.load(open(os.path.join(golden_dir, "metric_cli.json")))
    for th, rows in spec["csv"].items():
        for row in rows[1:]:
            key = row[0].split("_label")[0]
            if key == "huge":
                continue                                       # 25 Mpx case: checked on the GPU tier
            H, W = spec["specs"][key]
            r = om.evaluate_pair(synth.lattice_cells(H, W, seed=spec["gt_seed"]), synth.lattice_cells(H, W, seed=spec["seg_seed"]), float(th))
            got = [r["true_num"], r["pred_num"], r["tp"], r["fn"], r["fp"], r["precision"], r["recall"], r["dice"], r["f1"]]
            assert [float(v) for v in got] == [float(v) for v in row[1:]], (key, th)


# ---- 2. kernels + host logic vs the reference's own outputs ---------------------------------------------
@pytest.mark.parametrize("name", FIX)
def test_kernels_on_reference_fixtures(cm, golden_dir, fixtures, name):
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


def test_kernels_on_synthetic_pairs(cm, synthetic):
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


def test_evaluate_pair_matches_reference_cli_rows(cm, golden_dir):
    from cellseg_b200 import synth
    spec = json.load(open(os.path.join(golden_dir, "metric_cli.json")))
    for th, rows in spec["csv"].items():
        for row in rows[1:]:
            key = row[0].split("_label")[0]
            if key == "huge":
                continue
            H, W = spec["specs"][key]
            r = cm.evaluate_pair(synth.lattice_cells(H, W, seed=spec["gt_seed"]), synth.lattice_cells(H, W, seed=spec["seg_seed"]), float(th))
            assert [r["true_num"], r["pred_num"], r["tp"], r["fn"], r["fp"]] == [int(v) for v in row[1:6]]
            assert float(np.round(r["dice"], 4)) == float(row[8])


def test_patch_branch_logic_small_roi(cm):
    """The ≥25 Mpx branch with a small roi_size, product vs oracle (same tiling code path)."""
    from cellseg_b200 import synth
    gt, seg = synth.lattice_cells(230, 310, seed=3), synth.lattice_cells(230, 310, seed=4)
    n_h, n_w = -(-230 // 100), -(-310 // 100)
    gp = np.zeros((n_h * 100, n_w * 100), np.int32); sp = np.zeros_like(gp)
    gp[:230, :310] = gt; sp[:230, :310] = seg
    tot_o = np.zeros(5, int); tot_k = np.zeros(5, int)
    for i in range(n_h):
        for j in range(n_w):
            sl = (slice(100 * i, 100 * (i + 1)), slice(100 * j, 100 * (j + 1)))
            go = morph.relabel_sequential(om.remove_boundary_cells(gp[sl].copy()))[0]
            so = morph.relabel_sequential(om.remove_boundary_cells(sp[sl].copy()))[0]
            tot_o += np.array([go.max(), so.max(), *om.eval_tp_fp_fn(go, so, 0.5)])
            gk = cm.relabel_sequential(cm.remove_boundary_cells(gp[sl].copy()))
            sk = cm.relabel_sequential(cm.remove_boundary_cells(sp[sl].copy()))
            assert np.array_equal(gk, go) and np.array_equal(sk, so)
            tot_k += np.array([gk.max(), sk.max(), *cm.eval_tp_fp_fn(gk, sk, 0.5)])
    assert np.array_equal(tot_o, tot_k)


def test_remove_boundary_quirks(cm):
    # all-frame image without background: the smallest label survives (np.unique(mask*bd)[1:] skips it)
    for arr in (np.full((4, 7), 5, np.int32), np.array([[3, 3, 9], [9, 9, 9]], np.int32), np.full((3, 3), 0, np.int32),
                np.array([[0, 2, 2, 4]], np.int32)):
        a, b = arr.copy(), arr.copy()
        assert np.array_equal(cm.remove_boundary_cells(a), om.remove_boundary_cells(b)), arr
        assert np.array_equal(a, b)
    # dtype follows relabel_sequential (input dtype kept when it fits) and views are mutated in place
    big = np.zeros((40, 60), np.uint16); big[10:20, 10:20] = 700; big[0:5, 30:40] = 9; big[25:30, 25:30] = 12
    view = big[:, 5:55]
    ref_view = big.copy()[:, 5:55]
    out = cm.remove_boundary_cells(view)
    ref = om.remove_boundary_cells(ref_view)
    assert out.dtype == ref.dtype == np.uint16 and np.array_equal(out, ref) and np.array_equal(view, ref_view)


def test_relabel_sequential_and_empty(cm):
    rng = np.random.default_rng(0)
    lab = rng.choice(np.array([0, 3, 4, 9, 100, 1000, 65000], np.int32), size=(50, 70))
    assert np.array_equal(cm.relabel_sequential(lab), morph.relabel_sequential(lab)[0])
    z = np.zeros((30, 30), np.int32)
    assert np.array_equal(cm.relabel_sequential(z), z)
    assert cm.eval_tp_fp_fn(lab, z, 0.5) == (0, 0, 0)
    assert cm.dice(z > 0, z > 0) == 1.0 and cm.dice(z > 0, lab > 0) == 0.0


@pytest.mark.parametrize("H,W", [(5, 1), (1, 7), (4, 4), (9, 37), (33, 130), (64, 1028), (7, 1027)])
def test_row_oriented_label_kernels_on_odd_geometries(cm, H, W):
    """k_label_max / k_label_flags / k_remap / k_count3 walk rows in 4-pixel groups (16-byte loads when pitch and address allow):
    widths that are not multiples of 4, single rows / columns, rows longer than one CTA's reach, blobs touching the frame."""
    rng = np.random.default_rng(H * 1000 + W)
    lab = np.zeros((H, W), np.int32)
    for k in range(1, 1 + max(1, H * W // 40)):            # rectangles with distinct ids, later ones overwrite earlier ones
        y, x = rng.integers(0, H), rng.integers(0, W)
        lab[y:y + rng.integers(1, 6), x:x + rng.integers(1, 9)] = k * 3
    seg = np.roll(lab, 1, axis=1)
    work, ref_work = lab.copy(), lab.copy()
    assert np.array_equal(cm.remove_boundary_cells(work), om.remove_boundary_cells(ref_work)) and np.array_equal(work, ref_work)
    assert np.array_equal(cm.relabel_sequential(lab), morph.relabel_sequential(lab)[0])
    assert cm.dice(lab > 0, seg > 0) == om.dice(lab > 0, seg > 0)
    assert tuple(cm.eval_tp_fp_fn(lab, seg, 0.5)) == tuple(om.eval_tp_fp_fn(lab, seg, 0.5))
    r = cm.evaluate_pair(lab, seg, 0.5)
    o = om.evaluate_pair(lab.copy(), seg.copy(), 0.5)
    assert [r[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")] == [o[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")]
    assert round(float(r["dice"]), 4) == round(float(o["dice"]), 4)


def test_max_matching_matches_hungarian():
    from cellseg_b200 import compute_metric
    rng = np.random.default_rng(1)
    for _ in range(30):
        iou = rng.random((rng.integers(1, 9), rng.integers(1, 9))) * (rng.random((1, 1)) < 0.9)
        iou[rng.random(iou.shape) < 0.5] = 0
        for th in (0.0, 0.2, 0.5):
            assert compute_metric._true_positive(iou, th) == om.true_positive(iou, th)


def test_mismatched_pair_is_a_python_error_not_a_device_read(cm, tmp_path, capsys, monkeypatch):
    """A (gt, seg) pair of different shapes must raise before any kernel sees it (the kernels walk both maps with one
    geometry); the CLI records that file as failed and carries on, like the reference's per-file try/except
    (compute_metric.py:170, 253-257)."""
    from cellseg_b200 import synth, tiffio
    a, b = synth.lattice_cells(96, 120, seed=1), synth.lattice_cells(96, 128, seed=1)
    for fn in (cm.evaluate_pair, cm.eval_tp_fp_fn, cm._intersection_over_union, cm._label_overlap):
        with pytest.raises(ValueError, match="differ in shape"):
            fn(a, b)
    with pytest.raises(ValueError, match="differ in shape"):
        cm.dice(a > 0, b > 0)
    with pytest.raises(ValueError, match="2-D"):
        cm.evaluate_pair(a[None], a[None])
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "gt")); os.makedirs(os.path.join(d, "seg"))
    tiffio.imwrite_labels(os.path.join(d, "gt", "bad.tiff"), a); tiffio.imwrite_labels(os.path.join(d, "seg", "bad_label.tiff"), b)
    tiffio.imwrite_labels(os.path.join(d, "gt", "ok.tiff"), a); tiffio.imwrite_labels(os.path.join(d, "seg", "ok_label.tiff"), a)
    monkeypatch.setattr(sys, "argv", ["compute_metric", "-g", os.path.join(d, "gt"), "-s", os.path.join(d, "seg"), "--gt_suffix", ".tiff",
                                      "-o", d, "-n", "mm"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    cm.main()
    out = capsys.readouterr().out
    assert "failed cases: ['bad_label.tiff']" in out
    rows = open(os.path.join(d, "mm-0.5.csv")).read().strip().split("\n")
    assert len(rows) == 2 and rows[1].startswith("ok_label.tiff")


# ---- 3. multi-process CLI: files dealt out over ranks, rows gathered over gloo ---------------------------------
def _write_pairs(d, n):
    from cellseg_b200 import synth, tiffio
    os.makedirs(os.path.join(d, "gt")); os.makedirs(os.path.join(d, "seg"))
    for i in range(n):
        tiffio.imwrite_labels(os.path.join(d, "gt", f"img{i}.tiff"), synth.lattice_cells(96, 120, seed=i))
        tiffio.imwrite_labels(os.path.join(d, "seg", f"img{i}_label.tiff"), synth.lattice_cells(96, 120, seed=100 + i))


def _cli_worker(rank, world, port, d, q):
    import faulthandler
    faulthandler.enable()
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests", "emu"))
        import cellseg_b200  # noqa: F401
        import emu_lib as el
        from cellseg_b200 import capi, compute_metric as cmod
        capi._lib = el.lib()                         # test-only wiring of the product module to the emulated kernels (as the cm fixture)
        cmod._torch = lambda: torch
        cmod._stream = lambda t: None
        cmod._to_device_i32 = lambda a: a.to(torch.int32) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32))
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        sys.argv = ["compute_metric", "-g", os.path.join(d, "gt"), "-s", os.path.join(d, "seg"), "--gt_suffix", ".tiff", "-o", d, "-n", f"w{world}"]
        cmod.main()
        q.put((rank, True, ""))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_cli_file_sharded_over_gloo_ranks_equals_single_process(tmp_path):
    import sys as _sys
    import torch.multiprocessing as mp
    emu_lib.lib()
    d = str(tmp_path)
    _write_pairs(d, 5)
    ctx = mp.get_context("spawn")
    csv = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = 29900 + os.getpid() % 90 + world
        procs = [ctx.Process(target=_cli_worker, args=(r, world, port, d, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = []
        try:
            for _ in procs:
                res.append(q.get(timeout=240))
        except Exception:                                    # a rank that never reports: show what the others said and how they ended
            for p in procs:
                p.join(timeout=5)
            codes = [p.exitcode for p in procs]
            for p in procs:
                if p.is_alive():
                    p.terminate()
            raise AssertionError(f"world {world}: only {len(res)} of {len(procs)} ranks reported; exit codes {codes}; reports: {res}")
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok, _ in res), res
        csv[world] = open(os.path.join(d, f"w{world}-0.5.csv")).read()
    assert csv[1] == csv[2] and csv[1].count("\n") == 6


# ---- property test (hypothesis): random label maps of any small shape through the emulated metric kernels ---------------------------
from hypothesis import given, settings, strategies as hst


@settings(max_examples=30, deadline=None)
@given(h=hst.integers(1, 48), w=hst.integers(1, 70), nlab=hst.integers(1, 40), seed=hst.integers(0, 10**6), th=hst.sampled_from([0.5, 0.7, 0.9]),
       sparse=hst.booleans())
def test_property_metric_path_equals_oracle_on_random_maps(h, w, nlab, seed, th, sparse):
    """Random blocky label maps (non-sequential ids, labels touching the frame, ties at exact halves likely at th = 0.5): boundary
    removal + relabel, overlap / IoU matrices, tp/fp/fn and the whole evaluate_pair against the oracle restatement of the reference."""
    import emu_lib as el
    from cellseg_b200 import capi, compute_metric as cmod
    saved = (capi._lib, cmod._torch, cmod._stream, cmod._to_device_i32)
    capi._lib = el.lib()
    cmod._torch = lambda: torch
    cmod._stream = lambda t: None
    cmod._to_device_i32 = lambda a: a.to(torch.int32) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32))
    try:
        rng = np.random.default_rng(seed)
        ids = rng.choice(np.arange(1, 400), size=nlab, replace=False)

        def blocky():
            m = np.zeros((h, w), np.int32)
            for k in ids:
                if sparse and rng.random() < 0.5:
                    continue
                y, x = rng.integers(0, h), rng.integers(0, w)
                m[y:y + rng.integers(1, 7), x:x + rng.integers(1, 9)] = k
            return m
        gt, seg = blocky(), blocky()
        if rng.random() < 0.3:
            seg = np.where(rng.random((h, w)) < 0.5, gt, seg).astype(np.int32)      # correlated maps: real overlaps
        a, b = gt.copy(), gt.copy()
        assert np.array_equal(cmod.remove_boundary_cells(a), om.remove_boundary_cells(b)) and np.array_equal(a, b)
        assert np.array_equal(cmod.relabel_sequential(seg), morph.relabel_sequential(seg)[0])
        if seg.max() > 0:
            assert tuple(cmod.eval_tp_fp_fn(gt, seg, th)) == tuple(om.eval_tp_fp_fn(gt, seg, th))
        assert np.array_equal(cmod._label_overlap(gt, seg), om.label_overlap(gt, seg))
        assert np.array_equal(cmod._intersection_over_union(gt, seg), om.intersection_over_union(gt, seg))
        r, o = cmod.evaluate_pair(gt, seg, th), om.evaluate_pair(gt.copy(), seg.copy(), th)
        assert [r[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")] == [o[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")]
        assert round(float(r["dice"]), 4) == round(float(o["dice"]), 4)
    finally:
        capi._lib, cmod._torch, cmod._stream, cmod._to_device_i32 = saved
