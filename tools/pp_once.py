"""Run the mask→instance-map chain a few times on one workload (for ncu). Usage: pp_once.py [batch|wsi|noise] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellseg_b200 import postprocess, synth

which = sys.argv[1] if len(sys.argv) > 1 else "batch"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "batch":
    m = np.stack([synth.blob_mask(1024, 1024, s) for s in range(64)])
elif which == "wsi":
    m = synth.blob_mask(10000, 10000, 0)[None]
else:
    m = synth.noise_mask(4096, 4096, 0.45, 1)[None]
d = torch.from_numpy(np.ascontiguousarray(m)).cuda()
out = torch.empty(d.shape, dtype=torch.int32, device="cuda")
for _ in range(iters):
    postprocess.postprocess_device(d, out=out)
torch.cuda.synchronize()
print("done", which, iters)
