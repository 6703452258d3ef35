"""Odd-shaped calls of the post-processing chain (certify, tile kernels incl. stop / resume, merges, look-back scan, final), the
TMA-staged blend kernel and the strip path on one device — the target of compute-sanitizer runs:
    compute-sanitizer --tool memcheck|initcheck|synccheck python tools/sanitize_post.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cellseg_b200
from cellseg_b200 import postprocess, sliding_window as sw, synth
from oracle import morph

rng = np.random.default_rng(0)
masks = [synth.blob_mask(130, 259, 1), synth.noise_mask(65, 129, 0.5, 2), synth.noise_mask(200, 300, 0.2, 3), np.zeros((70, 140), bool),
         np.ones((64, 128), bool), synth.blob_mask(257, 130, 4) ^ (rng.random((257, 130)) < 0.05), synth.noise_mask(1, 300, 0.6, 5),
         synth.noise_mask(300, 1, 0.6, 6)]
for m in masks:
    for thr in ((64, 16), (10, 40), (300, 3), (0, 0)):
        lab, n = postprocess.postprocess_device(torch.from_numpy(m).cuda(), thr[0], thr[1], 2)
        assert np.array_equal(lab.cpu().numpy(), morph.postprocess_mask(m, thr[0], thr[1], 2)), (m.shape, thr)
    postprocess.certified_large(torch.from_numpy(m).cuda())
batch = np.stack([synth.noise_mask(150, 270, p, i) for i, p in enumerate((0.2, 0.5, 0.62))])
lab, n = postprocess.postprocess_device(torch.from_numpy(batch).cuda())
for i in range(3):
    assert np.array_equal(lab[i].cpu().numpy(), morph.postprocess_mask(batch[i]))
# blend, both kernels
for hw in ((300, 420), (200, 200), (513, 1030)):
    plan = sw.make_plan(*hw, 256, 0.25)
    wl = torch.randn((2 * plan.n_win, 3, 256, 256), device="cuda")
    os.environ.pop("CELLSEG_BLEND_TMA", None)
    a = sw.blend(wl, 2, plan, None, want=("mask",))["mask"].clone()
    os.environ["CELLSEG_BLEND_TMA"] = "1"
    b = sw.blend(wl, 2, plan, None, want=("mask",))["mask"].clone()
    os.environ.pop("CELLSEG_BLEND_TMA", None)
    assert torch.equal(a, b)
torch.cuda.synchronize()
print("sanitize_post: all calls completed")
