"""Pre-normalisation kernels (predict.py:14-21, 94-104) timed on their own: histogram + LUT application on
u8 RGB / u8 gray / u16 RGB images, CUDA events, next to the numpy restatement on the host.  Usage: prenorm_bench.py [edge] [--cpu]"""
import sys, os, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cellseg_b200
from cellseg_b200 import capi, synth

edge = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8192
lib = capi.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(0)
out = {"edge": edge, "cases": {}}
cases = {"u8_rgb": synth.modality_image(edge, edge, seed=0, modality=0),
         "u8_gray": synth.modality_image(edge, edge, seed=1, modality=2),
         "u16_rgb": rng.gamma(2.0, 600.0, (edge // 2, edge // 2, 3)).clip(0, 65535).astype(np.uint16)}
for name, img in cases.items():
    H, W = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    eb = img.dtype.itemsize
    bins = 256 if eb == 1 else 65536
    src = torch.from_numpy(np.ascontiguousarray(img).view(np.uint8)).cuda()
    hist = torch.empty((3, bins), dtype=torch.int32, device="cuda")
    lut = torch.randint(0, 256, (3, bins), dtype=torch.uint8, device="cuda")
    dst = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    res = {}
    for what in ("hist", "apply"):
        def run():
            if what == "hist":
                capi.check(lib, lib.cs_prenorm_hist(src.data_ptr(), eb, H, W, ch, hist.data_ptr(), st))
            else:
                capi.check(lib, lib.cs_prenorm_apply(src.data_ptr(), eb, H, W, ch, lut.data_ptr(), dst.data_ptr(), st))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = src.numel() + (dst.numel() if what == "apply" else 0)
        res[what] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1)}
    ref = np.stack([np.bincount(img.reshape(H * W, ch)[:, c], minlength=bins) for c in range(min(ch, 3))])
    res["hist_matches_numpy"] = bool(np.array_equal(hist.cpu().numpy().view(np.uint32)[:min(ch, 3)], ref))
    if "--cpu" in sys.argv and name == "u8_rgb":
        from oracle import pipeline as opl
        t0 = time.perf_counter(); opl.preprocess(img); res["numpy_s"] = round(time.perf_counter() - t0, 2)
    out["cases"][name] = res
print(json.dumps(out))
