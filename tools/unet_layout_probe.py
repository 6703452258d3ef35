"""Which memory format should the (out-of-scope, cuDNN) convolutions of the predictor see?  Times a 100-window forward
for NCHW / channels_last with the fused and the plain norm blocks and lists the top CUDA kernels of each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import cellseg_b200
from cellseg_b200 import unet
torch.backends.cudnn.benchmark = True
x0 = torch.rand(100, 3, 256, 256, device="cuda")

def run(tag, fused, cl, rows=14, fast=False):
    unet.FUSED_ADN = fused; unet.FAST_NHWC = fast
    m = unet.build_unet(3, seed=2022).eval().cuda()
    x = x0
    if cl:
        x = x0.contiguous(memory_format=torch.channels_last); m = m.to(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m(x)
        e1.record(); torch.cuda.synchronize()
        print(f"== {tag}: {e0.elapsed_time(e1) / 5:.3f} ms per 100-window forward")
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(x); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=rows, max_name_column_width=70))

run("channels-last runner (libcellseg stem/head/norm + cuDNN NHWC)", True, False, 30, fast=True)
run("fused NCHW", True, False)
if len(sys.argv) > 1:
    run("plain NCHW", False, False, 8)
    run("plain channels_last", False, True, 20)
