"""A few eager (no graph) calls of the fused engine on a small synthetic batch: target of ncu captures of every libcellseg
kernel of the step (predictor glue, blend, post-processing) at the bench's shapes.  usage: engine_once.py [images] [calls] [chunk]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cellseg_b200
from cellseg_b200 import sliding_window as sw
from cellseg_b200.unet import build_unet
torch.backends.cudnn.benchmark = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
imgs = torch.from_numpy(bench.synthetic_batch(n, 1024, 0)).cuda()
eng = sw.SegmentationEngine(build_unet(3, seed=2022).eval().cuda(), chunk=chunk)
for _ in range(calls):
    eng(imgs)
torch.cuda.synchronize()
