"""Channels-last predictor glue (csrc/unet_glue.cu) on the fiber emulator against plain PyTorch fp32 on the CPU:
stem convolutions, NHWC InstanceNorm+PReLU(+residual) from two-level statistics, transposed-convolution head.
The reference lines are monai's UNet built at baseline/predict.py:45-52."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_lib  # noqa: E402
from emu_lib import ptr, check  # noqa: E402


def nhwc(t):
    """[N,C,H,W] torch → contiguous numpy [N,H*W,C]"""
    return np.ascontiguousarray(t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]).numpy())


@pytest.mark.parametrize("N,H,W", [(2, 32, 48), (1, 6, 10), (1, 64, 64)])
def test_stem_matches_conv2d(N, H, W):
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 3, H, W, generator=g)
    w0 = torch.randn(16, 3, 3, 3, generator=g) * 0.3
    wr = torch.randn(16, 3, 3, 3, generator=g) * 0.3
    OHW = (H // 2) * (W // 2)
    nch = (OHW + 255) // 256
    A = np.zeros((N, OHW, 16), np.float32); R = np.zeros_like(A)
    part = np.zeros((N, nch, 16, 2), np.float32)
    xn, w0n, wrn = x.numpy(), w0.numpy(), wr.numpy()
    check(L.cs_unet_stem(ptr(xn), N, H, W, ptr(w0n), ptr(wrn), ptr(A), ptr(R), ptr(part), None))
    np.testing.assert_allclose(A, nhwc(F.conv2d(x, w0, None, 2, 1)), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(R, nhwc(F.conv2d(x, wr, None, 2, 1)), rtol=1e-5, atol=1e-5)
    # the chunk statistics of A feed the norm kernel directly
    a = torch.tensor([0.2])
    y = np.zeros_like(A)
    an = a.numpy()
    check(L.cs_nhwc_norm_prelu(ptr(A), None, ptr(y), ptr(an), None, N, OHW, 16, ptr(part), 256, 1e-5, None))
    ref = F.prelu(F.instance_norm(F.conv2d(x, w0, None, 2, 1), eps=1e-5), a)
    np.testing.assert_allclose(y, nhwc(ref), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 24, 40), (1, 32, 16, 16), (2, 64, 9, 7), (1, 128, 16, 16), (1, 256, 8, 8), (1, 4, 50, 50), (1, 8, 3, 3)])
def test_nhwc_norm_matches_instance_norm(N, C, H, W):
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(N, C, H, W, generator=g) * 2.5 + torch.randn(1, C, 1, 1, generator=g) * 3
    res = torch.randn(N, C, H, W, generator=g)
    rb = torch.randn(C, generator=g)
    a = torch.tensor([0.25])
    HW = H * W
    nch = L.cs_nhwc_stat_chunks(HW, C)
    assert nch == (HW + 2048 // (C // 4) - 1) // (2048 // (C // 4))
    xn, rn, rbn, an = nhwc(x), nhwc(res), rb.numpy(), a.numpy()
    part = np.zeros((N, nch, C, 2), np.float32)
    check(L.cs_nhwc_stats(ptr(xn), N, HW, C, ptr(part), None))
    ref = F.prelu(F.instance_norm(x, eps=1e-5), a)
    y = np.zeros_like(xn)
    check(L.cs_nhwc_norm_prelu(ptr(xn), None, ptr(y), ptr(an), None, N, HW, C, ptr(part), 2048 // (C // 4), 1e-5, None))
    np.testing.assert_allclose(y, nhwc(ref), rtol=1e-4, atol=2e-5)
    check(L.cs_nhwc_norm_prelu(ptr(xn), ptr(rn), ptr(y), ptr(an), ptr(rbn), N, HW, C, ptr(part), 2048 // (C // 4), 1e-5, None))
    np.testing.assert_allclose(y, nhwc(ref + res + rb.view(1, -1, 1, 1)), rtol=1e-4, atol=3e-5)
    z = xn.copy()                                               # in place
    check(L.cs_nhwc_norm_prelu(ptr(z), None, ptr(z), ptr(an), None, N, HW, C, ptr(part), 2048 // (C // 4), 1e-5, None))
    np.testing.assert_allclose(z, nhwc(ref), rtol=1e-4, atol=2e-5)


def test_nhwc_rejects_unsupported_channels():
    L = emu_lib.lib()
    x = np.zeros((1, 4, 12), np.float32); part = np.zeros((1, 1, 12, 2), np.float32)
    assert L.cs_nhwc_stats(ptr(x), 1, 4, 12, ptr(part), None) != 0
    assert L.cs_nhwc_stat_chunks(4, 12) == 0


@pytest.mark.parametrize("N,Hin,Win", [(2, 16, 24), (1, 5, 3), (1, 32, 32), (1, 4, 70), (1, 9, 67)])
def test_head_matches_torch(N, Hin, Win):
    """ConvTranspose2d(32→3) on the never-materialised cat, then IN → PReLU → conv3x3 + bias + identity residual."""
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(Hin)
    xa = torch.randn(N, 16, Hin, Win, generator=g)
    xb = torch.randn(N, 16, Hin, Win, generator=g)
    wt = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    w3 = torch.randn(3, 3, 3, 3, generator=g) * 0.3
    b3 = torch.randn(3, generator=g)
    a = torch.tensor([0.3])
    zt = F.conv_transpose2d(torch.cat([xa, xb], 1), wt, None, 2, 1, 1)
    yt = F.prelu(F.instance_norm(zt, eps=1e-5), a)
    ref = F.conv2d(yt, w3, b3, 1, 1) + yt
    H, W = 2 * Hin, 2 * Win
    nch = L.cs_unet_head_partial_chunks(Hin, Win)
    assert nch == ((Hin + 7) // 8) * ((Win + 31) // 32)
    z = np.zeros((N, 3, H, W), np.float32)
    part = np.zeros((N, nch, 3, 2), np.float32)
    xan, xbn, wtn = nhwc(xa), nhwc(xb), wt.numpy()
    check(L.cs_unet_head_convt(ptr(xan), ptr(xbn), N, Hin, Win, ptr(wtn), ptr(z), ptr(part), None))
    np.testing.assert_allclose(z, zt.numpy(), rtol=1e-4, atol=2e-5)
    out = np.zeros_like(z)
    w3n, b3n, an = w3.numpy(), b3.numpy(), a.numpy()
    check(L.cs_unet_head_final(ptr(z), ptr(part), N, H, W, ptr(an), ptr(w3n), ptr(b3n), 1e-5, ptr(out), None))
    np.testing.assert_allclose(out, ref.numpy(), rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize("N,H,W,roi", [(2, 40, 52, 32), (1, 20, 70, 32), (1, 32, 32, 32)])
def test_stem_from_u8_equals_gather_then_stem(N, H, W, roi):
    """The fused form (uint8 image → scale table → window pixels inside the first convolution) must produce exactly what
    cs_gather_windows_u8 followed by cs_unet_stem produces, for every window incl. shifted-back and zero-padded ones."""
    from cellseg_b200 import sliding_window as sw
    L = emu_lib.lib()
    rng = np.random.default_rng(H)
    img = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    img[0, :3] = 0
    plan = sw.make_plan(H, W, roi, 0.25)
    lut = emu_lib.scale_lut(img)
    ys, xs, _, _ = emu_lib.plan_arrays(plan)
    total = N * plan.n_win
    g = torch.Generator().manual_seed(1)
    w0 = (torch.randn(16, 3, 3, 3, generator=g) * 0.3).numpy(); wr = (torch.randn(16, 3, 3, 3, generator=g) * 0.3).numpy()
    OHW = (roi // 2) ** 2
    nch = (OHW + 255) // 256
    begin, count = 1 if total > 2 else 0, max(1, total - 2)
    win = emu_lib.gather(img, plan, begin, count, lut=lut)
    A0 = np.zeros((count, OHW, 16), np.float32); R0 = np.zeros_like(A0); P0 = np.zeros((count, nch, 16, 2), np.float32)
    check(L.cs_unet_stem(ptr(win), count, roi, roi, ptr(w0), ptr(wr), ptr(A0), ptr(R0), ptr(P0), None))
    A1 = np.zeros_like(A0); R1 = np.zeros_like(A0); P1 = np.zeros_like(P0)
    check(L.cs_unet_stem_u8(ptr(img), ptr(lut), N, H, W, roi, plan.ny, plan.nx, plan.pad_y, plan.pad_x, ptr(ys), ptr(xs), begin, count,
                            ptr(w0), ptr(wr), ptr(A1), ptr(R1), ptr(P1), None))
    assert np.array_equal(A0, A1) and np.array_equal(R0, R1) and np.array_equal(P0, P1)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 12, 20), (1, 64, 9, 7)])
def test_norm_block_writes_concat_slot(N, C, H, W):
    """The one-call norm block writes its result densely and/or into a channel slot of a wider NHWC tensor (skip concatenation)."""
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(C)
    x = torch.randn(N, C, H, W, generator=g) * 2 + 1
    a = torch.tensor([0.25])
    HW = H * W
    xn, an = nhwc(x), a.numpy()
    ws = np.zeros((N, L.cs_nhwc_stat_chunks(HW, C), C, 2), np.float32)
    ref = nhwc(F.prelu(F.instance_norm(x, eps=1e-5), a))
    y = np.zeros_like(xn)
    both = np.full((N, HW, 3 * C), 7.0, np.float32)
    check(L.cs_instnorm_prelu_nhwc(ptr(xn), None, ptr(y), ptr(an), None, N, HW, C, 1e-5, ptr(ws), ptr(both), 3 * C, C, None))
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=2e-5)
    assert np.array_equal(both[..., C:2 * C], y) and (both[..., :C] == 7).all() and (both[..., 2 * C:] == 7).all()
    both2 = np.full_like(both, 7.0)
    check(L.cs_instnorm_prelu_nhwc(ptr(xn), None, None, ptr(an), None, N, HW, C, 1e-5, ptr(ws), ptr(both2), 3 * C, C, None))
    assert np.array_equal(both, both2)
    assert L.cs_instnorm_prelu_nhwc(ptr(xn), None, None, ptr(an), None, N, HW, C, 1e-5, ptr(ws), ptr(both2), C + 2, 2, None) != 0


def test_convt_as_conv2x2_weight_is_exact():
    from cellseg_b200.unet import convt_as_conv2x2_weight
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(5, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, None, 2, 1, 1)
    full = F.conv2d(x, convt_as_conv2x2_weight(w), None, 1, 1)[:, :, 1:, 1:]               # [2, 4*3, 6, 7]
    got = full.reshape(2, 2, 2, 3, 6, 7).permute(0, 3, 4, 1, 5, 2).reshape(2, 3, 12, 14)   # depth to space
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 8, 12), (1, 32, 5, 3), (1, 64, 4, 4), (1, 4, 9, 20)])
def test_shuffle_norm_block_matches_convtranspose_then_norm(N, C, H, W):
    """x = conv2d(input, W2, padding 1) in NHWC with the unused first row/column → prelu(instance_norm(conv_transpose2d(input)))."""
    from cellseg_b200.unet import convt_as_conv2x2_weight
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(C + W)
    cin = 8
    inp = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cin, C, 3, 3, generator=g) * 0.3
    a = torch.tensor([0.2])
    ref = F.prelu(F.instance_norm(F.conv_transpose2d(inp, w, None, 2, 1, 1), eps=1e-5), a)
    xv = np.ascontiguousarray(F.conv2d(inp, convt_as_conv2x2_weight(w), None, 1, 1).permute(0, 2, 3, 1).numpy())     # [N, H+1, W+1, 4C]
    an = a.numpy()
    part = np.zeros((N, L.cs_nhwc_stat_chunks(H * W, 4 * C), 4 * C, 2), np.float32)
    y = np.zeros((N, 2 * H * 2 * W, C), np.float32)
    check(L.cs_instnorm_prelu_nhwc_shuffle(ptr(xv), ptr(y), ptr(an), N, H, W, C, 1e-5, ptr(part), None))
    np.testing.assert_allclose(y, nhwc(ref), rtol=2e-4, atol=3e-5)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 12, 20), (1, 64, 9, 7), (1, 128, 4, 4)])
def test_slice_norm_block(N, C, H, W):
    """x and the residual are channel slices of wider tensors (one merged convolution fed both)."""
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(C)
    wide = torch.randn(N, 2 * C, H, W, generator=g) * 2 + 1
    x2 = torch.randn(N, C, H, W, generator=g)
    rb = torch.randn(C, generator=g)
    a = torch.tensor([0.25])
    HW = H * W
    wn, x2n, an, rbn = nhwc(wide), nhwc(x2), a.numpy(), rb.numpy()
    ws = np.zeros((N, L.cs_nhwc_stat_chunks(HW, C), C, 2), np.float32)
    y = np.zeros((N, HW, C), np.float32)
    check(L.cs_instnorm_prelu_nhwc_slices(ptr(wn), 2 * C, 0, None, 0, 0, ptr(y), ptr(an), None, N, HW, C, 1e-5, ptr(ws), None, 0, 0, None))
    np.testing.assert_allclose(y, nhwc(F.prelu(F.instance_norm(wide[:, :C], eps=1e-5), a)), rtol=1e-4, atol=2e-5)
    both = np.full((N, HW, 2 * C), 7.0, np.float32)
    check(L.cs_instnorm_prelu_nhwc_slices(ptr(x2n), C, 0, ptr(wn), 2 * C, C, ptr(y), ptr(an), ptr(rbn), N, HW, C, 1e-5, ptr(ws), ptr(both), 2 * C, C, None))
    ref = F.prelu(F.instance_norm(x2, eps=1e-5), a) + wide[:, C:] + rb.view(1, -1, 1, 1)
    np.testing.assert_allclose(y, nhwc(ref), rtol=1e-4, atol=3e-5)
    assert np.array_equal(both[..., C:], y) and (both[..., :C] == 7).all()


# ---- property tests (hypothesis): random shapes through the emulated kernels ----------------------------------------------
from hypothesis import given, settings, strategies as hst


@settings(max_examples=25, deadline=None)
@given(n=hst.integers(1, 3), lc=hst.integers(2, 8), h=hst.integers(1, 37), w=hst.integers(1, 37), seed=hst.integers(0, 10**6),
       with_res=hst.booleans())
def test_property_norm_block_any_shape(n, lc, h, w, seed, with_res):
    """cs_instnorm_prelu_nhwc == prelu(instance_norm(x)) (+ residual) for every channel count the kernels accept and any plane size
    (single pixels included: variance 0 → output 0 (+ residual))."""
    L = emu_lib.lib()
    c = 1 << lc
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g) * torch.rand(1, generator=g).item() * 4 + torch.randn(1, c, 1, 1, generator=g) * 2
    res = torch.randn(n, c, h, w, generator=g)
    a = torch.tensor([0.1])
    xn, rn, an = nhwc(x), nhwc(res), a.numpy()
    ws = np.zeros((n, L.cs_nhwc_stat_chunks(h * w, c), c, 2), np.float32)
    y = np.zeros_like(xn)
    check(L.cs_instnorm_prelu_nhwc(ptr(xn), ptr(rn) if with_res else None, ptr(y), ptr(an), None, n, h * w, c, 1e-5, ptr(ws), None, 0, 0, None))
    ref = F.prelu(F.instance_norm(x, eps=1e-5), a) if h * w > 1 else torch.zeros_like(x)
    if with_res:
        ref = ref + res
    np.testing.assert_allclose(y, nhwc(ref), rtol=2e-4, atol=5e-5)


@settings(max_examples=12, deadline=None)
@given(hin=hst.integers(1, 20), win=hst.integers(1, 40), seed=hst.integers(0, 10**6))
def test_property_head_any_shape(hin, win, seed):
    L = emu_lib.lib()
    g = torch.Generator().manual_seed(seed)
    xa, xb = torch.randn(1, 16, hin, win, generator=g), torch.randn(1, 16, hin, win, generator=g)
    wt = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    w3 = torch.randn(3, 3, 3, 3, generator=g) * 0.3
    b3 = torch.randn(3, generator=g)
    a = torch.tensor([0.3])
    zt = F.conv_transpose2d(torch.cat([xa, xb], 1), wt, None, 2, 1, 1)
    yt = F.prelu(F.instance_norm(zt, eps=1e-5), a)
    ref = F.conv2d(yt, w3, b3, 1, 1) + yt
    z = np.zeros((1, 3, 2 * hin, 2 * win), np.float32)
    part = np.zeros((1, L.cs_unet_head_partial_chunks(hin, win), 3, 2), np.float32)
    xan, xbn, wtn, w3n, b3n, an = nhwc(xa), nhwc(xb), wt.numpy(), w3.numpy(), b3.numpy(), a.numpy()
    check(L.cs_unet_head_convt(ptr(xan), ptr(xbn), 1, hin, win, ptr(wtn), ptr(z), ptr(part), None))
    np.testing.assert_allclose(z, zt.numpy(), rtol=1e-4, atol=3e-5)
    out = np.zeros_like(z)
    check(L.cs_unet_head_final(ptr(z), ptr(part), 1, 2 * hin, 2 * win, ptr(an), ptr(w3n), ptr(b3n), 1e-5, ptr(out), None))
    np.testing.assert_allclose(out, ref.numpy(), rtol=5e-4, atol=2e-4)
