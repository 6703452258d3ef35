"""Quick device timing of the mask→instance-map chain (CUDA events). Not the contract bench."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellseg_b200 import postprocess, synth

def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2], ts[0]

def run(name, masks):
    m = torch.from_numpy(np.ascontiguousarray(masks)).cuda()
    if m.dim() == 2: m = m[None]
    N, H, W = m.shape
    out = torch.empty((N, H, W), dtype=torch.int32, device="cuda")
    med, best = timeit(lambda: postprocess.postprocess_device(m, out=out))
    px = N * H * W
    if os.environ.get("PP_KERNELS"):               # per-kernel device times (CUDA events around every launch; separate passes)
        from cellseg_b200 import capi
        capi.profile_enable(True)
        for _ in range(5):
            postprocess.postprocess_device(m, out=out)
        torch.cuda.synchronize()
        prof = capi.profile_read(); capi.profile_enable(False)
        print("   ", {k: round(ms / 5, 4) for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0])}, flush=True)
    print(json.dumps({"case": name, "N": N, "H": H, "W": W, "ms_med": round(med, 4), "ms_best": round(best, 4),
                      "Gpx_s": round(px / med / 1e6, 2), "GBs_at_9Bpx": round(9 * px / med / 1e6, 1),
                      "GBs_at_5Bpx": round(5 * px / med / 1e6, 1)}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["batch", "wsi", "noise"]
    if "batch" in which:
        run("blob 64x1024^2", np.stack([synth.blob_mask(1024, 1024, s) for s in range(64)]))
        run("blob 1x1024^2", synth.blob_mask(1024, 1024, 0))
    if "sparse" in which or "batch" in which:
        # a sparse microscopy field: ~25 cells of radius 14 per 1024², about half of the 64x128 tiles are empty (k_tile's fast path)
        run("sparse cells 64x1024^2", np.stack([synth.lattice_cells(1024, 1024, seed=s, pitch=200, radius=14) > 0 for s in range(64)]))
    if "unet" in which or "batch" in which:
        # masks of the bench's random-init predictor (4 saved 1024² masks, tiled): salt-like, ~1000 specks per 64x128 tile
        f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "unet_masks_4x1024.npz")
        m4 = np.unpackbits(np.load(f)["packed"]).reshape(4, 1024, 1024).astype(bool)
        run("unet-like 64x1024^2", np.tile(m4, (16, 1, 1)))
    if "wsi" in which:
        run("blob 10000^2", synth.blob_mask(10000, 10000, 0))
        run("blob 4096^2", synth.blob_mask(4096, 4096, 0))
    if "noise" in which:
        for p in (0.3, 0.45, 0.6):
            run(f"noise p={p} 4096^2", synth.noise_mask(4096, 4096, p, 1))
