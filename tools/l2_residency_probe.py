"""Does a level-0 conv → statistics → apply chain run faster per window when the sub-batch's activations (1 MiB per window and
tensor) stay L2-resident?  Times the chain for several sub-batch sizes; prints us per window."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import cellseg_b200
from cellseg_b200 import capi
torch.backends.cudnn.benchmark = True
lib = capi.load()
a = torch.tensor([0.25], device="cuda")
for c, hw in ((16, 128), (32, 64)):
    w = (torch.randn(c, c, 3, 3, device="cuda") * 0.1).contiguous(memory_format=torch.channels_last)
    for n in (16, 25, 32, 50, 64, 100, 200, 400):
        x = torch.randn(n, c, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        ws = torch.empty((n, lib.cs_nhwc_stat_chunks(hw * hw, c), c, 2), device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        def chain(x):
            y = F.conv2d(x, w, None, 1, 1)
            capi.check(lib, lib.cs_instnorm_prelu_nhwc(y.data_ptr(), x.data_ptr(), y.data_ptr(), a.data_ptr(), None, n, hw * hw, c, 1e-5, ws.data_ptr(), None, 0, 0, st))
            return y
        with torch.no_grad():
            y = x
            for _ in range(3): y = chain(y)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = x
                for _ in range(8): y = chain(y)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
        print(f"C={c} {hw}x{hw}  sub-batch {n:4d} ({n * c * hw * hw * 4 / 2**20:6.1f} MiB per tensor): {e0.elapsed_time(e1) / 5 / 8 / n * 1000:7.2f} us per window per (conv + norm block)")
