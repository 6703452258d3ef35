"""Where the (out-of-scope, PyTorch/cuDNN) UNet forward spends its time: top CUDA kernels of one 100-window forward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import cellseg_b200
from cellseg_b200.unet import build_unet
torch.backends.cudnn.benchmark = True
m = build_unet(3, seed=2022).eval().cuda()
cl = len(sys.argv) > 1 and sys.argv[1] == "cl"
x = torch.rand(100, 3, 256, 256, device="cuda")
if cl:
    x = x.contiguous(memory_format=torch.channels_last); m = m.to(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m(x)
    e1.record(); torch.cuda.synchronize()
    print("ms per 100-window forward:", e0.elapsed_time(e1) / 5)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(x); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
