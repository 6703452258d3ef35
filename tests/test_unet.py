"""The predictor (monai-free rebuild of the reference's UNet, baseline/predict.py:45-52): topology facts from
SURVEY.md §8 a-3 on the CPU; on the GPU the fused InstanceNorm+PReLU(+residual) kernel against the PyTorch modules."""
import pytest
import torch


def test_topology_matches_monai_unet():
    from cellseg_b200.unet import build_unet
    m = build_unet(3, seed=2022).eval()
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 1626072 and len(sd) == 63
    for k in ("model.0.conv.unit0.conv.weight", "model.0.conv.unit0.adn.A.weight", "model.0.residual.weight",
              "model.2.0.conv.weight", "model.2.1.conv.unit0.conv.weight"):
        assert k in sd, k
    assert any(k.startswith("model.1.submodule.") for k in sd)
    assert sd["model.2.0.conv.weight"].shape == (32, 3, 3, 3)          # ConvTranspose2d(32 → 3)
    with torch.no_grad():
        y = m(torch.rand(1, 3, 256, 256))
    assert y.shape == (1, 3, 256, 256)


@pytest.mark.gpu
@pytest.mark.parametrize("planes,hw", [((7, 16), (128, 128)), ((3, 256), (16, 16)), ((5, 3), (256, 256)), ((2, 64), (32, 32)), ((1, 5), (6, 10))])
def test_fused_adn_kernel_matches_torch(planes, hw):
    import ctypes as C
    from cellseg_b200 import capi
    lib = capi.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(*planes, *hw, device="cuda", generator=g) * 3 + 0.7
    res = torch.randn(*planes, *hw, device="cuda", generator=g)
    a = torch.tensor([0.25], device="cuda")
    ref = torch.nn.functional.prelu(torch.nn.functional.instance_norm(x, eps=1e-5), a)
    for r, want in ((None, ref), (res, ref + res)):
        y = torch.empty_like(x)
        capi.check(lib, lib.cs_instnorm_prelu(x.data_ptr(), None if r is None else r.data_ptr(), y.data_ptr(), a.data_ptr(), None, planes[1],
                                              planes[0] * planes[1], hw[0] * hw[1], 1e-5, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert torch.allclose(y, want, rtol=1e-5, atol=2e-6)
    rb = torch.randn(planes[1], device="cuda", generator=g)
    y = torch.empty_like(x)
    capi.check(lib, lib.cs_instnorm_prelu(x.data_ptr(), res.data_ptr(), y.data_ptr(), a.data_ptr(), rb.data_ptr(), planes[1],
                                          planes[0] * planes[1], hw[0] * hw[1], 1e-5, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.allclose(y, ref + res + rb.view(1, -1, 1, 1), rtol=1e-5, atol=4e-6)
    z = x.clone()                                                     # in place
    capi.check(lib, lib.cs_instnorm_prelu(z.data_ptr(), None, z.data_ptr(), a.data_ptr(), None, planes[1], planes[0] * planes[1], hw[0] * hw[1], 1e-5,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.allclose(z, ref, rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_fused_unet_matches_unfused():
    from cellseg_b200 import unet
    torch.backends.cudnn.allow_tf32 = False
    m = unet.build_unet(3, seed=2022).eval().cuda()
    x = torch.rand(6, 3, 256, 256, device="cuda")
    with torch.no_grad():
        unet.FAST_NHWC = False             # the NCHW module graph with / without the fused norm kernel
        try:
            unet.FUSED_ADN = True
            a = m(x)
            unet.FUSED_ADN = False
            b = m(x)
        finally:
            unet.FUSED_ADN = unet.FAST_NHWC = True
    print("fused vs unfused UNet: max abs diff", (a - b).abs().max().item(), "scale", b.abs().mean().item())
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.gpu
@pytest.mark.parametrize("N,C,H,W", [(3, 16, 128, 128), (2, 32, 64, 64), (2, 64, 32, 32), (2, 128, 16, 16), (2, 256, 16, 16), (1, 8, 5, 7)])
def test_nhwc_norm_kernels_match_torch(N, C, H, W):
    import ctypes as Ct
    from cellseg_b200 import capi
    lib = capi.load()
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(N, C, H, W, device="cuda", generator=g) * 2 + torch.randn(1, C, 1, 1, device="cuda", generator=g) * 3
    res = torch.randn(N, C, H, W, device="cuda", generator=g)
    rb = torch.randn(C, device="cuda", generator=g)
    a = torch.tensor([0.25], device="cuda")
    xn, rn = _nhwc(x), _nhwc(res)
    px = 2048 // (C // 4)
    part = torch.empty((N, lib.cs_nhwc_stat_chunks(H * W, C), C, 2), device="cuda")
    capi.check(lib, lib.cs_nhwc_stats(xn.data_ptr(), N, H * W, C, part.data_ptr(), st))
    y = torch.empty_like(xn)
    capi.check(lib, lib.cs_nhwc_norm_prelu(xn.data_ptr(), rn.data_ptr(), y.data_ptr(), a.data_ptr(), rb.data_ptr(), N, H * W, C, part.data_ptr(), px, 1e-5, st))
    ref = torch.nn.functional.prelu(torch.nn.functional.instance_norm(x, eps=1e-5), a) + res + rb.view(1, -1, 1, 1)
    assert torch.allclose(y, _nhwc(ref), rtol=1e-5, atol=5e-6)


@pytest.mark.gpu
def test_stem_and_head_kernels_match_torch():
    import ctypes as Ct
    import torch.nn.functional as F
    from cellseg_b200 import capi
    lib = capi.load()
    torch.backends.cudnn.allow_tf32 = False
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(7)
    N, H, W = 3, 256, 256
    x = torch.rand(N, 3, H, W, device="cuda", generator=g)
    w0 = torch.randn(16, 3, 3, 3, device="cuda", generator=g) * 0.3
    wr = torch.randn(16, 3, 3, 3, device="cuda", generator=g) * 0.3
    A = torch.empty(N, H // 2, W // 2, 16, device="cuda"); R = torch.empty_like(A)
    part = torch.empty(N, (H // 2) * (W // 2) // 256, 16, 2, device="cuda")
    w0h, wrh = w0.cpu().numpy(), wr.cpu().numpy()          # host copies: these weights travel as kernel parameters
    capi.check(lib, lib.cs_unet_stem(x.data_ptr(), N, H, W, w0h.ctypes.data, wrh.ctypes.data, A.data_ptr(), R.data_ptr(), part.data_ptr(), st))
    assert torch.allclose(A, _nhwc(F.conv2d(x, w0, None, 2, 1)), rtol=1e-5, atol=1e-5)
    assert torch.allclose(R, _nhwc(F.conv2d(x, wr, None, 2, 1)), rtol=1e-5, atol=1e-5)
    # the opt-in tensor-core stem (mma.sync TF32, 3xTF32 split) must produce the same convolutions and statistics
    import os
    A2 = torch.empty_like(A); R2 = torch.empty_like(A); part2 = torch.empty_like(part)
    os.environ["CELLSEG_STEM_MMA"] = "1"
    try:
        capi.check(lib, lib.cs_unet_stem(x.data_ptr(), N, H, W, w0h.ctypes.data, wrh.ctypes.data, A2.data_ptr(), R2.data_ptr(), part2.data_ptr(), st))
    finally:
        del os.environ["CELLSEG_STEM_MMA"]
    assert torch.allclose(A2, A, rtol=1e-5, atol=1e-5) and torch.allclose(R2, R, rtol=1e-5, atol=1e-5)
    assert torch.allclose(part2, part, rtol=1e-3, atol=1e-4)
    a = torch.tensor([0.2], device="cuda")
    y = torch.empty_like(A)
    capi.check(lib, lib.cs_nhwc_norm_prelu(A.data_ptr(), None, y.data_ptr(), a.data_ptr(), None, N, (H // 2) * (W // 2), 16, part.data_ptr(), 256, 1e-5, st))
    assert torch.allclose(y, _nhwc(F.prelu(F.instance_norm(F.conv2d(x, w0, None, 2, 1), eps=1e-5), a)), rtol=1e-4, atol=1e-5)
    # head
    xa = torch.randn(N, 16, 128, 128, device="cuda", generator=g); xb = torch.randn(N, 16, 128, 128, device="cuda", generator=g)
    wt = torch.randn(32, 3, 3, 3, device="cuda", generator=g) * 0.2
    w3 = torch.randn(3, 3, 3, 3, device="cuda", generator=g) * 0.3
    b3 = torch.randn(3, device="cuda", generator=g)
    zt = F.conv_transpose2d(torch.cat([xa, xb], 1), wt, None, 2, 1, 1)
    yt = F.prelu(F.instance_norm(zt, eps=1e-5), a)
    ref = F.conv2d(yt, w3, b3, 1, 1) + yt
    z = torch.empty(N, 3, H, W, device="cuda"); pz = torch.empty(N, lib.cs_unet_head_partial_chunks(128, 128), 3, 2, device="cuda"); out = torch.empty_like(z)
    xan, xbn = _nhwc(xa), _nhwc(xb)
    wth, w3h, b3h = wt.cpu().numpy(), w3.cpu().numpy(), b3.cpu().numpy()
    capi.check(lib, lib.cs_unet_head_convt(xan.data_ptr(), xbn.data_ptr(), N, 128, 128, wth.ctypes.data, z.data_ptr(), pz.data_ptr(), st))
    assert torch.allclose(z, zt, rtol=1e-4, atol=2e-5)
    capi.check(lib, lib.cs_unet_head_final(z.data_ptr(), pz.data_ptr(), N, H, W, a.data_ptr(), w3h.ctypes.data, b3h.ctypes.data, 1e-5, out.data_ptr(), st))
    assert torch.allclose(out, ref, rtol=2e-4, atol=5e-5)


@pytest.fixture(params=["pair-auto", "pair-always", "pair-always-lag64", "pair-always-lag-all"])
def norm_pair_mode(request, monkeypatch):
    """The statistics + apply pair as two launches (auto: small tensors) or as the single fused launch (k_nhwc_pair), with the apply CTAs
    trailing closely (they do wait) or by the whole batch."""
    monkeypatch.setenv("CELLSEG_NORM_PAIR", "1" if request.param == "pair-auto" else "2")
    if request.param.endswith("lag64"):
        monkeypatch.setenv("CELLSEG_PAIR_LAG", "64")
    if request.param.endswith("lag-all"):
        monkeypatch.setenv("CELLSEG_PAIR_LAG", "100000000")
    return request.param


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [256, 64])
def test_channels_last_runner_matches_module_graph(hw, norm_pair_mode):
    from cellseg_b200 import unet
    torch.backends.cudnn.allow_tf32 = False
    m = unet.build_unet(3, seed=2022).eval().cuda()
    x = torch.rand(5, 3, hw, hw, device="cuda")
    with torch.no_grad():
        a = m(x)
        assert m._runner, "the channels-last runner must be the CUDA inference path"
        unet.FAST_NHWC = unet.FUSED_ADN = False
        try:
            b = m(x)
        finally:
            unet.FAST_NHWC = unet.FUSED_ADN = True
    print("channels-last runner vs module graph: max abs diff", (a - b).abs().max().item(), "scale", b.abs().mean().item())
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("N,C,H,W", [(3, 16, 128, 128), (2, 32, 64, 64), (2, 64, 32, 32), (2, 128, 16, 16), (2, 256, 16, 16), (1, 8, 5, 7),
                                     (2, 16, 100, 100), (1, 4, 3, 1), (1, 16, 160, 160)])
def test_single_call_norm_block_matches_torch(N, C, H, W, norm_pair_mode):
    """cs_instnorm_prelu_nhwc: statistics + apply pair with a workspace; one-cluster-per-sample single pass without one."""
    import ctypes as Ct
    from cellseg_b200 import capi
    lib = capi.load()
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(N, C, H, W, device="cuda", generator=g) * 2 + torch.randn(1, C, 1, 1, device="cuda", generator=g) * 3
    res = torch.randn(N, C, H, W, device="cuda", generator=g)
    rb = torch.randn(C, device="cuda", generator=g)
    a = torch.tensor([0.25], device="cuda")
    xn, rn = _nhwc(x), _nhwc(res)
    ws = torch.empty((N, max(1, lib.cs_nhwc_stat_chunks(H * W, C)), C, 2), device="cuda")
    ref = torch.nn.functional.prelu(torch.nn.functional.instance_norm(x, eps=1e-5), a)
    y = torch.empty_like(xn)
    capi.check(lib, lib.cs_instnorm_prelu_nhwc(xn.data_ptr(), rn.data_ptr(), y.data_ptr(), a.data_ptr(), rb.data_ptr(), N, H * W, C, 1e-5, ws.data_ptr(), None, 0, 0, st))
    assert torch.allclose(y, _nhwc(ref + res + rb.view(1, -1, 1, 1)), rtol=1e-5, atol=5e-6)
    z = xn.clone()                                                    # in place, no residual
    capi.check(lib, lib.cs_instnorm_prelu_nhwc(z.data_ptr(), None, z.data_ptr(), a.data_ptr(), None, N, H * W, C, 1e-5, ws.data_ptr(), None, 0, 0, st))
    assert torch.allclose(z, _nhwc(ref), rtol=1e-5, atol=5e-6)
    both = torch.full((N, H * W, 2 * C + 8), 7.0, device="cuda")         # concat slot only (no dense output)
    capi.check(lib, lib.cs_instnorm_prelu_nhwc(xn.data_ptr(), rn.data_ptr(), None, a.data_ptr(), rb.data_ptr(), N, H * W, C, 1e-5, ws.data_ptr(),
                                               both.data_ptr(), 2 * C + 8, C + 4, st))
    assert torch.equal(both[..., C + 4:2 * C + 4], y.view(N, H * W, C)) and bool((both[..., :C + 4] == 7).all()) and bool((both[..., 2 * C + 4:] == 7).all())
    if H * W * C <= 262144:                                           # no workspace → thread-block cluster kernel (DSMEM exchange)
        y2 = torch.empty_like(xn)
        capi.check(lib, lib.cs_instnorm_prelu_nhwc(xn.data_ptr(), rn.data_ptr(), y2.data_ptr(), a.data_ptr(), rb.data_ptr(), N, H * W, C, 1e-5, None, None, 0, 0, st))
        assert torch.allclose(y2, _nhwc(ref + res + rb.view(1, -1, 1, 1)), rtol=1e-5, atol=5e-6)
    else:
        assert lib.cs_instnorm_prelu_nhwc(xn.data_ptr(), None, y.data_ptr(), a.data_ptr(), None, N, H * W, C, 1e-5, None, None, 0, 0, st) != 0


def test_channels_last_runner_support_matrix():
    """The fast CUDA runner is only claimed for the reference's topology; anything else runs the module graph."""
    from cellseg_b200 import unet
    assert unet.ChannelsLastRunner.supports(unet.build_unet(3, seed=0))
    assert not unet.ChannelsLastRunner.supports(unet.UNet(2, 3, 4))                                  # 4 classes: head kernels are 32 -> 3
    assert not unet.ChannelsLastRunner.supports(unet.UNet(2, 1, 3))                                  # 1 input channel
    assert not unet.ChannelsLastRunner.supports(unet.UNet(2, 3, 3, channels=(8, 16, 32), strides=(2, 2)))
    m = unet.UNet(2, 3, 3, channels=(16, 32, 64), strides=(2, 2))                                    # shallower, same ends: supported
    assert unet.ChannelsLastRunner.supports(m)
    with torch.no_grad():
        assert m(torch.rand(1, 3, 32, 32)).shape == (1, 3, 32, 32)                                   # CPU input → module graph


def test_host_pipeline_schedule():
    from cellseg_b200.sliding_window import SegmentationEngine as E
    assert E._schedule(64, 16, True) == [4, 12, 16, 16, 12, 4]
    assert E._schedule(64, 16, False) == [16, 16, 16, 16]
    for n in range(1, 70):
        for sub in (1, 2, 3, 4, 7, 8, 16):
            s = E._schedule(n, min(sub, n), True)
            assert sum(s) == n and all(0 < v <= sub for v in s)
