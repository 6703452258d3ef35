"""GPU tier: the two drop-in CLIs end to end on the reference's own sample images (data/Train_Labeled, copied to
tests/golden/ref_data): predict → <name>_label.tiff (int32, deflate) → compute_metric → CSV with the reference's columns."""
import csv
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from oracle import morph, pipeline as opl
from oracle import sliding_window as osw

pytestmark = pytest.mark.gpu


def test_predict_then_compute_metric(tmp_path, golden_dir, monkeypatch, capsys):
    from cellseg_b200 import compute_metric, predict, tiffio
    from cellseg_b200.unet import build_unet
    inp, out, gtd = tmp_path / "inputs", tmp_path / "outputs", tmp_path / "gt"
    inp.mkdir(); gtd.mkdir()
    ref = os.path.join(golden_dir, "ref_data")
    for f in ("cell_00849.png", "cell_00528.tif"):
        shutil.copy(os.path.join(ref, f), inp / f)
    for f in ("cell_00849_label.tiff", "cell_00528_label.tiff"):
        shutil.copy(os.path.join(ref, f), gtd / f)
    # a checkpoint in the reference's format (predict.py:80-81), seeded like the reference's training script
    ckpt = tmp_path / "work_dir"
    ckpt.mkdir()
    torch.manual_seed(2022)
    model = build_unet(3)
    torch.save({"model_state_dict": model.state_dict()}, ckpt / "best_Dice_model.pth")
    torch.backends.cudnn.allow_tf32 = False
    predict.main(["-i", str(inp), "-o", str(out), "--model_path", str(ckpt), "--model_name", "unet", "--show_overlay"])
    log = capsys.readouterr().out
    assert log.count("Prediction finished:") == 2
    names = sorted(os.listdir(out))
    assert names == ["cell_00528_label.tiff", "cell_00849_label.tiff", "overlay_cell_00528.tif", "overlay_cell_00849.png"]
    model = model.eval().cuda()
    for stem, ext in (("cell_00849", ".png"), ("cell_00528", ".tif")):
        lab = tiffio.imread(str(out / f"{stem}_label.tiff"))
        assert lab.dtype == np.int32
        img = tiffio.imread(str(inp / (stem + ext)))
        pre = opl.preprocess(img)
        assert lab.shape == pre.shape[:2]
        # oracle pipeline with the same GPU predictor (whole-call tier: equal unless a pixel sits on the 0.5 knife edge)
        with torch.no_grad():
            blended = osw.sliding_window_inference(osw.scale_to_unit(pre), (256, 256), 4, lambda d: model(d.contiguous().cuda()).cpu())
        p = torch.softmax(blended, dim=1)[0, 1].numpy()
        if np.abs(p - 0.5).min() > 1e-4:
            assert np.array_equal(lab, morph.postprocess_mask(p > 0.5))
        assert lab.max() == len(np.unique(lab)) - 1                        # ids are 1..n
    # metric CLI on the produced files, reference flags and CSV layout (compute_metric.py:136-147, 157-167, 259-263)
    monkeypatch.setattr(sys, "argv", ["compute_metric", "-g", str(gtd), "-s", str(out), "--gt_suffix", "_label.tiff", "--seg_suffix", "_label.tiff",
                                      "-o", str(tmp_path), "-n", "demo", "-thre", "0.5", "0.7"])
    compute_metric.main()
    for th in ("0.5", "0.7"):
        rows = list(csv.reader(open(tmp_path / f"demo-{th}.csv")))
        assert rows[0] == ["names", "true_num", "pred_num", "correct_num(TP)", "missed_num(FN)", "wrong_num(FP)", "precision", "recall", "dice", "F1"]
        assert [r[0] for r in rows[1:]] == ["cell_00528_label.tiff", "cell_00849_label.tiff"]
        from oracle import metric as om
        for r in rows[1:]:
            gt = tiffio.imread(str(gtd / r[0])); seg = tiffio.imread(str(out / r[0]))
            o = om.evaluate_pair(gt.copy(), seg.copy(), float(th))
            assert [int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5])] == [o["true_num"], o["pred_num"], o["tp"], o["fn"], o["fp"]]
            assert [float(r[6]), float(r[7]), float(r[8]), float(r[9])] == [float(o["precision"]), float(o["recall"]), float(o["dice"]), float(o["f1"])]
