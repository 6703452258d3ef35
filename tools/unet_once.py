"""A few forwards of the channels-last predictor runner over 100 windows (target of ncu captures of the glue kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cellseg_b200
from cellseg_b200 import unet
torch.backends.cudnn.benchmark = True
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = unet.build_unet(3, seed=2022).eval().cuda()
x = torch.rand(100, 3, 256, 256, device="cuda")
with torch.no_grad():
    for _ in range(n):
        m(x)
torch.cuda.synchronize()
