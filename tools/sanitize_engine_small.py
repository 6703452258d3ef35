"""One engine call on a single-window image: compute-sanitizer initcheck target (which kernel, if any, reads uninitialised memory)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cellseg_b200
from cellseg_b200 import sliding_window as sw, synth, unet
m = unet.build_unet(3, seed=2022).eval().cuda()
img = torch.from_numpy(synth.modality_image(200, 200, seed=1, modality=0)[None]).cuda()
labels, n = sw.SegmentationEngine(m, chunk=5)(img)
torch.cuda.synchronize()
print("engine call done", int(n[0]))
