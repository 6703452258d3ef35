"""Small, odd-shaped calls of every predictor-glue kernel and of the fused engine — the target of compute-sanitizer runs
(memcheck / racecheck / synccheck):  compute-sanitizer --tool memcheck python tools/sanitize_glue.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cellseg_b200
from cellseg_b200 import capi, sliding_window as sw, synth, unet
lib = capi.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
m = unet.build_unet(3, seed=2022).eval().cuda()
kernels_only = len(sys.argv) > 1 and sys.argv[1] == "kernels"      # skip the cuDNN-heavy part (slow under initcheck)
with torch.no_grad():
    for shape in (() if kernels_only else ((3, 3, 256, 256), (2, 3, 64, 96), (1, 3, 34, 70))):
        y = m(torch.rand(*shape, device="cuda"))
        assert y.shape == shape and torch.isfinite(y).all()
    eng = sw.SegmentationEngine(m, chunk=5)
    for hw in (() if kernels_only else ((300, 420), (200, 200), (513, 257))):
        img = torch.from_numpy(synth.modality_image(*hw, seed=1, modality=0)[None]).cuda()
        labels, n = eng(img)
        assert labels.shape == (1,) + hw
    # stem (both sources) and head on a small window batch
    g = torch.Generator(device="cuda").manual_seed(0)
    win = torch.rand(2, 3, 32, 48, device="cuda", generator=g)
    w0 = np.random.default_rng(0).standard_normal((16, 3, 3, 3)).astype(np.float32); wr = w0[::-1].copy()
    A = torch.empty(2, 16 * 24, 16, device="cuda"); R = torch.empty_like(A); pa = torch.empty(2, 2, 16, 2, device="cuda")
    capi.check(lib, lib.cs_unet_stem(win.data_ptr(), 2, 32, 48, w0.ctypes.data, wr.ctypes.data, A.data_ptr(), R.data_ptr(), pa.data_ptr(), st))
    yA = torch.empty_like(A)
    a1 = torch.tensor([0.25], device="cuda")
    capi.check(lib, lib.cs_nhwc_norm_prelu(A.data_ptr(), R.data_ptr(), yA.data_ptr(), a1.data_ptr(), None, 2, 16 * 24, 16, pa.data_ptr(), 256, 1e-5, st))
    wt = np.random.default_rng(1).standard_normal((32, 3, 3, 3)).astype(np.float32)
    w3 = np.random.default_rng(2).standard_normal((3, 3, 3, 3)).astype(np.float32); b3 = np.zeros(3, np.float32)
    z = torch.empty(2, 3, 32, 48, device="cuda"); pz = torch.empty(2, lib.cs_unet_head_partial_chunks(16, 24), 3, 2, device="cuda"); out = torch.empty_like(z)
    capi.check(lib, lib.cs_unet_head_convt(yA.data_ptr(), A.data_ptr(), 2, 16, 24, wt.ctypes.data, z.data_ptr(), pz.data_ptr(), st))
    capi.check(lib, lib.cs_unet_head_final(z.data_ptr(), pz.data_ptr(), 2, 32, 48, a1.data_ptr(), w3.ctypes.data, b3.ctypes.data, 1e-5, out.data_ptr(), st))
    assert torch.isfinite(out).all()
    # norm block variants on odd sizes
    a = torch.tensor([0.25], device="cuda")
    for n, c, h, w in ((2, 16, 7, 9), (1, 64, 5, 3), (1, 256, 3, 3), (3, 4, 33, 17)):
        x = torch.randn(n, h * w, c, device="cuda"); res = torch.randn_like(x); y = torch.empty_like(x)
        ws = torch.empty((n, lib.cs_nhwc_stat_chunks(h * w, c), c, 2), device="cuda")
        capi.check(lib, lib.cs_instnorm_prelu_nhwc(x.data_ptr(), res.data_ptr(), y.data_ptr(), a.data_ptr(), None, n, h * w, c, 1e-5, ws.data_ptr(), None, 0, 0, st))
        capi.check(lib, lib.cs_instnorm_prelu_nhwc(x.data_ptr(), res.data_ptr(), y.data_ptr(), a.data_ptr(), None, n, h * w, c, 1e-5, None, None, 0, 0, st))
        wide = torch.randn(n, h * w, 2 * c, device="cuda"); both = torch.empty(n, h * w, 3 * c, device="cuda")
        capi.check(lib, lib.cs_instnorm_prelu_nhwc_slices(wide.data_ptr(), 2 * c, c, wide.data_ptr(), 2 * c, 0, y.data_ptr(), a.data_ptr(), None, n, h * w, c, 1e-5,
                                                          ws.data_ptr(), both.data_ptr(), 3 * c, 2 * c, st))
        if c <= 64:
            xs = torch.randn(n, h + 1, w + 1, 4 * c, device="cuda"); ys = torch.empty(n, 4 * h * w, c, device="cuda")
            ws4 = torch.empty((n, lib.cs_nhwc_stat_chunks(h * w, 4 * c), 4 * c, 2), device="cuda")
            capi.check(lib, lib.cs_instnorm_prelu_nhwc_shuffle(xs.data_ptr(), ys.data_ptr(), a.data_ptr(), n, h, w, c, 1e-5, ws4.data_ptr(), st))
torch.cuda.synchronize()
print("sanitize_glue: all calls completed")
