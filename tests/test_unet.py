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
        unet.FUSED_ADN = True
        a = m(x)
        unet.FUSED_ADN = False
        b = m(x)
        unet.FUSED_ADN = True
    print("fused vs unfused UNet: max abs diff", (a - b).abs().max().item(), "scale", b.abs().mean().item())
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)
