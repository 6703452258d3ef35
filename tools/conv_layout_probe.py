"""Per-convolution timing of the predictor's layers (100 windows) in NCHW vs channels_last, cudnn.benchmark on."""
import sys, os, torch
import torch.nn.functional as F
torch.backends.cudnn.benchmark = True
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
layers = [  # name, transposed, cin, cout, k, stride, in_hw
    ("d0.u0 3->16 s2", 0, 3, 16, 3, 2, 256), ("d0.u1 16->16", 0, 16, 16, 3, 1, 128),
    ("d1.u0 16->32 s2", 0, 16, 32, 3, 2, 128), ("d1.u1 32->32", 0, 32, 32, 3, 1, 64),
    ("d2.u0 32->64 s2", 0, 32, 64, 3, 2, 64), ("d2.u1 64->64", 0, 64, 64, 3, 1, 32),
    ("d3.u0 64->128 s2", 0, 64, 128, 3, 2, 32), ("d3.u1 128->128", 0, 128, 128, 3, 1, 16),
    ("b.u0 128->256", 0, 128, 256, 3, 1, 16), ("b.u1 256->256", 0, 256, 256, 3, 1, 16), ("b.res 1x1 128->256", 0, 128, 256, 1, 1, 16),
    ("u3.T 384->64", 1, 384, 64, 3, 2, 16), ("u3.c 64->64", 0, 64, 64, 3, 1, 32),
    ("u2.T 128->32", 1, 128, 32, 3, 2, 32), ("u2.c 32->32", 0, 32, 32, 3, 1, 64),
    ("u1.T 64->16", 1, 64, 16, 3, 2, 64), ("u1.c 16->16", 0, 16, 16, 3, 1, 128),
    ("u0.T 32->3", 1, 32, 3, 3, 2, 128), ("u0.c 3->3", 0, 3, 3, 3, 1, 256),
]
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1000
tot = {"nchw": 0.0, "cl": 0.0, "best": 0.0}
with torch.no_grad():
    for name, tr, cin, cout, k, s, hw in layers:
        x = torch.randn(N, cin, hw, hw, device="cuda")
        w = torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), device="cuda") * 0.05
        p = (k - 1) // 2
        res = {}
        for fmt in ("nchw", "cl"):
            xx = x.contiguous(memory_format=torch.channels_last) if fmt == "cl" else x
            ww = w.contiguous(memory_format=torch.channels_last) if fmt == "cl" else w
            if tr: fn = lambda: F.conv_transpose2d(xx, ww, None, s, p, s - 1)
            else: fn = lambda: F.conv2d(xx, ww, None, s, p)
            res[fmt] = timeit(fn)
        out_hw = hw * s if tr else hw // s
        mb = (cin * hw * hw + cout * out_hw * out_hw) * 4 * N / 1e6
        floor = mb / 6459.3 * 1000
        mult = 2 if ("s2" in name and not tr) else 1     # the strided down convs appear twice (unit0 + residual)
        for f in res: tot[f] += res[f] * mult
        tot["best"] += min(res.values()) * mult
        print(f"{name:22s} nchw {res['nchw']:8.1f} us   cl {res['cl']:8.1f} us   x{mult}  traffic {mb:7.1f} MB  hbm floor {floor:6.1f} us")
print("totals (us):", {k: round(v, 1) for k, v in tot.items()})
