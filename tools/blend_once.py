"""Run the fused blend kernel a few times on random window logits (for ncu / timing). Usage: blend_once.py [N] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cellseg_b200 import capi, sliding_window as sw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
plan = sw.make_plan(1024, 1024, 256, 0.25)
wl = torch.randn((N * plan.n_win, 3, 256, 256), device="cuda")
lib = capi.load()
wsb = lib.cs_postprocess_workspace_bytes(N, 1024, 1024)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
slot = lib.cs_postprocess_packed_slot(ws.data_ptr(), ws.numel(), N, 1024, 1024)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for a, b in ev:
    a.record(); sw.blend(wl, N, plan, None, want=(), packed_ptr=slot); b.record()
torch.cuda.synchronize()
ts = [a.elapsed_time(b) for a, b in ev]
byts = wl.numel() * 4 + N * 1024 * 1024 / 8
print("blend ms", [round(t, 4) for t in ts], "GB/s best", round(byts / min(ts) / 1e6, 1))
