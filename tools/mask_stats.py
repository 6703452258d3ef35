"""What do the bench's foreground masks (random-init UNet on synthetic images) look like?  fg fraction, component counts,
segments per 32-pixel word — to know which regime the tile kernel is in."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cellseg_b200 import sliding_window as sw, postprocess
from cellseg_b200.unet import build_unet
torch.backends.cudnn.benchmark = True
imgs = torch.from_numpy(bench.synthetic_batch(8, 1024, 0)).cuda()
eng = sw.SegmentationEngine(build_unet(3, seed=2022).eval().cuda(), chunk=100)
logits, b = eng.window_logits(imgs)
mask = sw.blend(logits, 8, b["plan"], None, want=("mask",))["mask"].bool()
for n in range(8):
    m = mask[n]
    raw = int(postprocess.label(m, 1).max()); bg = int(postprocess.label(~m, 1).max())
    fin = int(eng(imgs[n:n + 1])[1][0])
    w = m.view(1024, 32, 32)
    starts = (w & ~torch.nn.functional.pad(w, (1, 0))[..., :32]).sum(-1)      # segments per word (approx: word-local)
    print(f"img {n} (modality {n % 4}): fg {m.float().mean().item():.3f}  4-conn fg comps {raw}  bg comps {bg}  final instances {fin}  "
          f"segments/word mean {starts.float().mean().item():.2f}  words with >=3 segs {(starts >= 3).float().mean().item():.3f}")
