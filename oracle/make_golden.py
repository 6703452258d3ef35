"""Generate tests/golden/metric_*.npz by running the REFERENCE's own baseline/compute_metric.py
(imported from /root/reference in the build container; it cannot travel to the GPU box, the vectors do).

skimage / tifffile are not installable offline, so the import is given stub modules:
  skimage.segmentation.relabel_sequential → oracle.morph.relabel_sequential (our restatement),
  tifffile.imread / skimage.io.imread      → PIL / numpy readers.
Everything else (numba _label_overlap, _intersection_over_union, scipy Hungarian _true_positive,
eval_tp_fp_fn, remove_boundary_cells, dice, and main() with its ≥25 Mpx tiling branch and CSV
rounding) is the reference's unmodified code.

    python -m oracle.make_golden            # from the repo root, in the build container
"""
from __future__ import annotations

import csv
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/baseline/compute_metric.py"
OUT = os.path.join(ROOT, "tests", "golden")


def _read_image(path):
    if path.endswith(".npy"):
        return np.load(path)
    from PIL import Image
    return np.array(Image.open(path))


def load_reference():
    from oracle import morph
    sk = types.ModuleType("skimage")
    seg = types.ModuleType("skimage.segmentation")
    seg.relabel_sequential = morph.relabel_sequential
    io = types.ModuleType("skimage.io")
    io.imread = _read_image
    sk.segmentation, sk.io = seg, io
    tif = types.ModuleType("tifffile")
    tif.imread = _read_image
    sys.modules.update({"skimage": sk, "skimage.segmentation": seg, "skimage.io": io, "tifffile": tif})
    spec = importlib.util.spec_from_file_location("ref_compute_metric", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_pairs():
    """(name, gt, seg) cases from SURVEY.md §8d config 4, at sizes small enough to commit."""
    from cellseg_b200 import synth
    H = W = 384
    gt = synth.lattice_cells(H, W, seed=0)
    seg = synth.lattice_cells(H, W, seed=1)
    cases = [("lattice", gt, seg)]
    d = seg.copy(); d[np.isin(d, [5, 17, 40, 41])] = 0
    cases.append(("dropped_pred_cells", gt, d))
    e = gt.copy(); e[np.isin(e, [3, 9, 77])] = 0
    cases.append(("dropped_gt_cells", e, seg))
    m = seg.copy(); m[m == 30] = 29; m[m == 58] = 57
    cases.append(("merged_pred_cells", gt, m))
    s = seg.copy()
    for lab in (20, 66, 100):
        ys, xs = np.nonzero(s == lab)
        s[ys[xs > xs.mean()], xs[xs > xs.mean()]] = s.max() + 1
    cases.append(("split_pred_cells", gt, s))
    # exact-half split: a 10x10 GT square predicted as two 5x10 halves plus an equal-size spill → IoU ties at 0.5
    g2 = np.zeros((64, 64), np.int32); p2 = np.zeros((64, 64), np.int32)
    g2[10:20, 10:20] = 1; p2[10:20, 10:15] = 1; p2[10:20, 15:20] = 2
    g2[30:40, 30:40] = 2; p2[30:40, 30:40] = 3; p2[30:35, 40:50] = 3      # IoU = 100/150
    g2[45:55, 10:20] = 3; p2[45:55, 10:20] = 4; p2[45:55, 20:30] = 4      # IoU = 100/200 = 0.5 exactly
    cases.append(("iou_half_ties", g2, p2))
    cases.append(("empty_pred", gt, np.zeros_like(gt)))
    cases.append(("empty_gt", np.zeros_like(gt), seg))
    nl = gt.copy(); nl[nl > 0] = nl[nl > 0] * 7 + 100            # non-sequential label values
    cases.append(("non_sequential_labels", nl, seg))
    return cases


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    from PIL import Image
    # ---- real fixtures of the reference repo (data/Train_Labeled/labels) -------------------------
    fx = {}
    for name in ("cell_00225", "cell_00528", "cell_00849"):
        g = np.array(Image.open(os.path.join(OUT, "ref_data", name + "_label.tiff"))).astype(np.int32)
        fx[name + "_nlabels"] = len(np.unique(g)) - 1
        work = g.copy()
        rb = ref.remove_boundary_cells(work)
        fx[name + "_rb"] = rb.astype(np.int32)
        fx[name + "_rb_mutated"] = work
        fx[name + "_self"] = np.array(ref.eval_tp_fp_fn(rb, rb, 0.5), dtype=np.int64)
        rolled = ref.remove_boundary_cells(np.roll(g, 3, axis=1).copy())
        for th in (0.5, 0.7, 0.9):
            fx[f"{name}_roll3_th{th}"] = np.array(ref.eval_tp_fp_fn(rb, rolled, th), dtype=np.int64)
        fx[name + "_dice_roll3"] = np.float64(ref.dice(g > 0, np.roll(g, 3, axis=1) > 0))
    np.savez_compressed(os.path.join(OUT, "metric_fixtures.npz"), **fx)
    # ---- synthetic pairs ---------------------------------------------------------------------------
    sy = {}
    names = []
    for name, gt, seg in synthetic_pairs():
        names.append(name)
        sy[name + "_gt"], sy[name + "_seg"] = gt, seg
        sy[name + "_dice"] = np.float64(ref.dice(gt > 0, seg > 0))
        a, b = gt.astype(np.int32).copy(), seg.astype(np.int32).copy()
        ra, rb = ref.remove_boundary_cells(a), ref.remove_boundary_cells(b)
        sy[name + "_gt_rb"], sy[name + "_seg_rb"] = ra.astype(np.int32), rb.astype(np.int32)
        sy[name + "_gt_mut"], sy[name + "_seg_mut"] = a, b
        for th in (0.3, 0.5, 0.7):
            sy[f"{name}_tpfpfn_{th}"] = np.array(ref.eval_tp_fp_fn(ra, rb, th), dtype=np.int64)
        sy[f"{name}_raw_tpfpfn_0.5"] = np.array(ref.eval_tp_fp_fn(gt, seg, 0.5), dtype=np.int64)
        if ra.max() > 0 and rb.max() > 0 and ra.max() * rb.max() < 40000:
            sy[name + "_overlap"] = ref._label_overlap(ra, rb).astype(np.uint64)
            sy[name + "_iou"] = ref._intersection_over_union(ra, rb)
    sy["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "metric_synthetic.npz"), **sy)
    # ---- the reference CLI end to end (incl. the >= 25 Mpx patch branch and CSV rounding) ---------------
    from cellseg_b200 import synth
    with tempfile.TemporaryDirectory() as td:
        gdir, sdir = os.path.join(td, "gt"), os.path.join(td, "seg")
        os.makedirs(gdir); os.makedirs(sdir)
        specs = {"small": (384, 384), "wide": (300, 1000), "huge": (5003, 5001)}
        for k, (H, W) in specs.items():
            np.save(os.path.join(gdir, k + ".npy"), synth.lattice_cells(H, W, seed=10))
            np.save(os.path.join(sdir, k + "_label.npy"), synth.lattice_cells(H, W, seed=11))
        rows = {}
        for th in (0.5, 0.7):
            argv = sys.argv
            sys.argv = ["compute_metric", "-g", gdir, "-s", sdir, "--gt_suffix", ".npy", "--seg_suffix", "_label.npy",
                        "-thre", str(th), "-o", td, "-n", "golden"]
            try:
                ref.main()
            finally:
                sys.argv = argv
            with open(os.path.join(td, f"golden-{th}.csv")) as fh:
                rows[str(th)] = list(csv.reader(fh))
        import json
        with open(os.path.join(OUT, "metric_cli.json"), "w") as fh:
            json.dump({"specs": specs, "gt_seed": 10, "seg_seed": 11, "csv": rows}, fh, indent=1)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
