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
with torch.no_grad():
    for shape in ((3, 3, 256, 256), (2, 3, 64, 96), (1, 3, 34, 70)):
        y = m(torch.rand(*shape, device="cuda"))
        assert y.shape == shape and torch.isfinite(y).all()
    eng = sw.SegmentationEngine(m, chunk=5)
    for hw in ((300, 420), (200, 200), (513, 257)):
        img = torch.from_numpy(synth.modality_image(*hw, seed=1, modality=0)[None]).cuda()
        labels, n = eng(img)
        assert labels.shape == (1,) + hw
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
