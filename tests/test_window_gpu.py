"""GPU tier: sliding-window kernels and the fused engine against the oracle.

Tier 1 (bit-exact): blended logits and instance maps on identical window logits.
Tier 2 (north_star tolerance 1e-5 relative): class-1 probabilities against torch's CPU softmax.
"""
import numpy as np
import pytest
import torch

from oracle import morph
from oracle import sliding_window as osw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sw():
    assert torch.cuda.is_available()
    from cellseg_b200 import sliding_window
    return sliding_window


@pytest.fixture(scope="module")
def unet():
    from cellseg_b200.unet import build_unet
    torch.backends.cudnn.allow_tf32 = False          # parity runs compare CPU and GPU forwards
    torch.backends.cuda.matmul.allow_tf32 = False
    return build_unet(3, seed=2022).eval().cuda()


@pytest.mark.parametrize("maxv", [255, 180, 1, 0])
def test_scale_and_float_bit_exact(sw, maxv):
    rng = np.random.default_rng(maxv)
    img = rng.integers(0, maxv + 1, (3, 301, 407, 3), dtype=np.uint8)
    img[0, 5, 6, 2] = maxv
    out = sw.to_unit_float(torch.from_numpy(img).cuda()).cpu().numpy()
    for n in range(3):
        assert np.array_equal(out[n], osw.scale_to_unit(img[n]).numpy()[0], equal_nan=True)


@pytest.mark.parametrize("H,W", [(1024, 1024), (200, 200), (500, 500), (512, 700), (257, 256), (100, 900)])
def test_gather_matches_monai_slicing(sw, H, W):
    rng = np.random.default_rng(H + W)
    img = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    x = torch.cat([osw.scale_to_unit(img[n]) for n in range(2)])
    seen = []
    osw.sliding_window_inference(x, (256, 256), 4, lambda d: (seen.append(d.clone()), d)[1])
    ref = torch.cat(seen)
    plan = sw.make_plan(H, W, 256, 0.25)
    d = torch.from_numpy(img).cuda()
    got_u8 = sw.gather_windows(d, plan, 0, 2 * plan.n_win, lut=sw.scale_lut(d)).cpu()
    got_f32 = sw.gather_windows(x.cuda(), plan, 0, 2 * plan.n_win).cpu()
    assert torch.equal(got_u8, ref) and torch.equal(got_f32, ref)


@pytest.mark.parametrize("H,W,mode,variant", [(1024, 1024, "constant", "A"), (500, 500, "constant", "A"),
                                              (200, 200, "constant", "A"), (512, 700, "gaussian", "A"),
                                              (600, 520, "gaussian", "B"), (257, 1000, "constant", "A")])
def test_blend_bit_exact_on_identical_logits(sw, H, W, mode, variant):
    N = 2
    plan = sw.make_plan(H, W, 256, 0.25)
    g = torch.Generator().manual_seed(H * 7 + W)
    wl = torch.randn((N * plan.n_win, 3, 256, 256), generator=g) * 2.0
    ref = osw.blend_from_window_logits(wl, N, H, W, 256, 0.25, mode, gaussian_variant=variant)
    imap = sw.importance_map(256, mode, variant=variant)
    out = sw.blend(wl.cuda(), N, plan, imap, want=("blended", "p1", "mask"))
    assert torch.equal(out["blended"].cpu(), ref)                       # tier 1: bit-exact blend
    p_ref = torch.softmax(ref, dim=1)[:, 1]
    p = out["p1"].cpu()
    assert torch.allclose(p, p_ref, rtol=1e-5, atol=1e-7)               # tier 2
    p_gpu = torch.softmax(out["blended"], dim=1)[:, 1].cpu()            # torch's own CUDA softmax
    assert torch.allclose(p, p_gpu, rtol=1e-6, atol=1e-8)
    sure = (p_ref - 0.5).abs() > 1e-6
    assert torch.equal(out["mask"].cpu().bool()[sure], (p_ref > 0.5)[sure])


@pytest.mark.parametrize("mode", ["constant", "gaussian"])
def test_mask_only_kernel_equals_softmax_kernel(sw, mode):
    """The mask-only blend kernel decides p1 > 0.5 without the softmax division; it must agree
    bit-for-bit with the kernel that computes p1, including on ties and knife-edge pixels."""
    N = 2
    plan = sw.make_plan(1024, 768, 256, 0.25)
    g = torch.Generator().manual_seed(7)
    wl = torch.randn((N * plan.n_win, 3, 256, 256), generator=g) * 2.0
    # two classes neck and neck, the third far away: p1 within a few ulp of 0.5, exact ties included
    k = wl.shape[0] // 2
    delta = torch.randint(-3, 4, wl[:k, 1].shape, generator=g).float() * 2.0 ** -22
    wl[:k, 1] = wl[:k, 0] * (1.0 + delta)
    wl[:k, 2] = -30.0
    wl[: k // 2, 2] = wl[: k // 2, 0] - 0.5 ** 20          # three-way near ties
    imap = sw.importance_map(256, mode)
    d = wl.cuda()
    full = sw.blend(d, N, plan, imap, want=("mask", "p1"))
    fast = sw.blend(d, N, plan, imap, want=("mask",))
    assert torch.equal(full["mask"], fast["mask"])
    assert torch.equal(full["mask"].bool(), full["p1"] > 0.5)
    frac = full["mask"].float().mean().item()
    assert 0.05 < frac < 0.95


@pytest.mark.parametrize("mode", ["constant", "gaussian"])
@pytest.mark.parametrize("hw", [(1024, 768), (300, 420), (200, 200), (513, 1030), (64, 2000)])
def test_tma_staged_kernel_equals_direct_kernel(sw, mode, hw, monkeypatch):
    """The opt-in TMA-staged blend kernel (CELLSEG_BLEND_TMA=1: tensor-map boxes of window logits into shared memory, mbarrier
    byte counting) runs the direct kernel's arithmetic on the same values: identical masks, on geometries with padding,
    shifted last windows, partial tiles and widths that are not multiples of 4 (those tiles take its direct fallback)."""
    import os
    H, W = hw
    N = 3
    plan = sw.make_plan(H, W, 256, 0.25)
    g = torch.Generator().manual_seed(11)
    wl = (torch.randn((N * plan.n_win, 3, 256, 256), generator=g) * 1.5).cuda()
    imap = sw.importance_map(256, mode)
    monkeypatch.delenv("CELLSEG_BLEND_TMA", raising=False)
    direct = sw.blend(wl, N, plan, imap, want=("mask",))["mask"].clone()
    monkeypatch.setenv("CELLSEG_BLEND_TMA", "1")
    staged = sw.blend(wl, N, plan, imap, want=("mask",))["mask"].clone()
    assert torch.equal(direct, staged)
    assert 0.02 < direct.float().mean().item() < 0.98


def test_dropin_sliding_window_inference_matches_oracle(sw, unet):
    from cellseg_b200 import synth
    img = synth.modality_image(600, 700, seed=0)
    x = osw.scale_to_unit(img).cuda()          # the reference's permuted-HWC view (predict.py:108)
    seen = []
    with torch.no_grad():
        got = sw.sliding_window_inference(x, (256, 256), 4, lambda w: (seen.append(unet(w)), seen[-1])[1])
        # identical-logits tier: the oracle blends the very window logits the drop-in saw → bit-exact
        ref = osw.blend_from_window_logits(torch.cat(seen).cpu(), 1, 600, 700, 256)
        assert torch.equal(got.cpu(), ref)
        # whole-call tier: oracle drives the same GPU predictor itself (cuDNN may pick another algorithm per call)
        ref2 = osw.sliding_window_inference(x.cpu(), (256, 256), 4, lambda d: unet(d.contiguous().cuda()).cpu())
        assert torch.allclose(got.cpu(), ref2, rtol=1e-4, atol=1e-5)
    cpu_model = type(unet)().eval()
    cpu_model.load_state_dict(unet.state_dict())
    with torch.no_grad():
        ref_cpu = osw.sliding_window_inference(x.cpu(), (256, 256), 4, cpu_model)
    assert torch.allclose(got.cpu(), ref_cpu, rtol=1e-3, atol=1e-4)     # cuDNN vs CPU convolution rounding only


@pytest.mark.parametrize("shape,seed", [((1024, 1024), 0), ((512, 512), 1), ((200, 200), 2), ((700, 900), 3)])
def test_engine_instance_maps_bit_exact(sw, unet, shape, seed):
    """Fused engine vs the oracle chain fed the engine's own window logits (identical-logits tier)."""
    from cellseg_b200 import synth
    imgs = np.stack([np.atleast_3d(synth.modality_image(*shape, seed=seed + 4 * k, modality=(seed + k) % 2 * 3)) for k in range(2)])
    if imgs.shape[-1] == 1:
        imgs = np.repeat(imgs, 3, -1)
    eng = sw.SegmentationEngine(unet, chunk=16)
    d = torch.from_numpy(imgs).cuda()
    labels, counts = eng(d)
    labels = labels.cpu().numpy().copy()
    logits, b = eng.window_logits(d)
    plan = b["plan"]
    ref_blend = osw.blend_from_window_logits(logits.cpu(), 2, shape[0], shape[1], 256)
    out = sw.blend(logits, 2, plan, None, want=("mask", "p1"))
    p_ref = torch.softmax(ref_blend, dim=1)[:, 1]
    assert torch.allclose(out["p1"].cpu(), p_ref, rtol=1e-5, atol=1e-7)
    mask = out["mask"].cpu().numpy().astype(bool)
    sure = ((p_ref - 0.5).abs() > 1e-6).numpy()
    assert np.array_equal(mask[sure], (p_ref.numpy() > 0.5)[sure])
    for n in range(2):
        ref = morph.postprocess_mask(mask[n])
        assert np.array_equal(labels[n], ref)
        assert int(counts[n]) == ref.max()


def test_engine_matches_oracle_pipeline_end_to_end(sw, unet):
    """Whole reference pipeline on CPU-side oracle code (same GPU predictor) vs the engine."""
    from cellseg_b200 import synth
    img = synth.modality_image(640, 512, seed=5, modality=0)
    eng = sw.SegmentationEngine(unet, chunk=4)
    labels, _ = eng(torch.from_numpy(img[None]).cuda())
    with torch.no_grad():
        blended = osw.sliding_window_inference(osw.scale_to_unit(img), (256, 256), 4, lambda d: unet(d.contiguous().cuda()).cpu())
    p = torch.softmax(blended, dim=1)[0, 1].numpy()
    if np.abs(p - 0.5).min() > 1e-4:                 # no pixel near the knife edge (cuDNN algorithm choice moves logits by ~1e-6)
        assert np.array_equal(labels[0].cpu().numpy(), morph.postprocess_mask(p > 0.5))


def test_segment_host_pipelined_equals_device_call(sw, unet):
    """Host buffers in/out through three streams and sub-batches == one device-resident call."""
    from cellseg_b200 import synth
    torch.backends.cudnn.benchmark = False
    imgs = np.stack([synth.modality_image(512, 640, seed=s, modality=0) for s in range(7)])
    eng = sw.SegmentationEngine(unet, chunk=12)
    ref, cnt = eng(torch.from_numpy(imgs).cuda())
    ref, cnt = ref.cpu().clone(), cnt.cpu().clone()
    host = torch.from_numpy(imgs).pin_memory()
    for sub in (3, 7, 16):
        out, counts = eng.segment_host(host, sub_batch=sub)
        # sub-batches change the window batches handed to cuDNN: identical counts and all but a knife-edge pixel or two
        assert torch.equal(counts.cpu(), cnt) or (counts.cpu() - cnt).abs().max() <= 2
        assert ((out != ref).float().mean().item()) < 1e-3
    out, counts = eng.segment_host(host, sub_batch=7)
    assert torch.equal(out, ref) and torch.equal(counts.cpu(), cnt)


@pytest.mark.parametrize("graph", [False, True])
def test_segment_host_calls_pipelined_across_call_boundaries(sw, unet, graph):
    """wait=False: several calls — two geometries interleaved, the same geometry back to back, different host buffers — are queued
    without draining the streams in between (buffer reuse is ordered by per-slot events); after wait() every output equals what the
    same call returns on its own."""
    from cellseg_b200 import synth
    torch.backends.cudnn.benchmark = False
    a = np.stack([synth.modality_image(300, 520, seed=s, modality=0) for s in range(5)])
    b = np.stack([synth.modality_image(512, 384, seed=10 + s, modality=3) for s in range(4)])
    a2 = np.ascontiguousarray(a[::-1])
    eng = sw.SegmentationEngine(unet, chunk=12, cuda_graph=graph)
    hosts = [torch.from_numpy(x).pin_memory() for x in (a, b, a2, a, b)]
    want = []
    for h in hosts:                                           # each call on its own, drained
        out, cnt = eng.segment_host(h, sub_batch=2)
        want.append((out.clone(), cnt.cpu().clone()))
    outs = [eng.segment_host(h, sub_batch=2, wait=False) for h in hosts]
    eng.wait()
    for (out, cnt), (wo, wc) in zip(outs, want):
        assert torch.equal(out, wo) and torch.equal(cnt.cpu(), wc)
    # one sub-batch per call (sub_batch >= N): consecutive calls must land in alternating slots
    want1 = []
    for h in hosts[:3]:
        out, cnt = eng.segment_host(h, sub_batch=8)
        want1.append((out.clone(), cnt.cpu().clone()))
    outs = [eng.segment_host(h, sub_batch=8, wait=False) for h in (hosts[0], hosts[2], hosts[0], hosts[1], hosts[2])]
    eng.wait()
    for (out, cnt), i in zip(outs, (0, 2, 0, 1, 2)):
        assert torch.equal(out, want1[i][0]) and torch.equal(cnt.cpu(), want1[i][1])
    # and once more over the same buffers (events of the finished calls are stale but harmless)
    outs = [eng.segment_host(h, sub_batch=2, wait=False) for h in hosts[:3]]
    eng.wait()
    for (out, cnt), (wo, wc) in zip(outs, want[:3]):
        assert torch.equal(out, wo) and torch.equal(cnt.cpu(), wc)


def test_fused_stem_path_equals_gather_path(sw, unet):
    """window_logits through ChannelsLastRunner.from_images (windows read from the uint8 image inside the first convolutions,
    logits written in place) vs the explicit gather → predictor → copy path."""
    from cellseg_b200 import synth
    imgs = np.stack([synth.modality_image(300, 520, seed=s, modality=s % 4 if s % 4 != 2 else 0) for s in range(3)])
    d = torch.from_numpy(imgs).cuda()
    eng = sw.SegmentationEngine(unet, chunk=7)
    assert eng._fused_runner(d.device) is not None
    fused = eng.window_logits(d)[0].clone()
    plan = sw.make_plan(300, 520, 256, 0.25)
    lut = sw.scale_lut(d)
    with torch.no_grad():
        ref = torch.cat([unet(sw.gather_windows(d, plan, g, min(7, 3 * plan.n_win - g), lut=lut)) for g in range(0, 3 * plan.n_win, 7)])
    assert fused.shape == ref.shape
    assert torch.allclose(fused, ref, rtol=1e-4, atol=1e-5)


def test_whole_call_graph_equals_eager(sw, unet):
    """cuda_graph=True replays the complete call (scale table → predictor forwards → blend → post-processing) as one graph."""
    from cellseg_b200 import synth, capi
    imgs = np.stack([synth.modality_image(512, 384, seed=s, modality=0) for s in range(3)])
    d = torch.from_numpy(imgs).cuda()
    eager = sw.SegmentationEngine(unet, chunk=10)
    ref, cnt = eager(d)
    ref, cnt = ref.clone(), cnt.clone()
    eng = sw.SegmentationEngine(unet, chunk=10, cuda_graph=True)
    lib = capi.load()
    for it in range(3):
        l0 = lib.cs_kernel_launch_count()
        g0 = eng.graph_kernel_launches
        out, c = eng(d)
        torch.cuda.synchronize()
        if it > 0:                           # replays launch nothing from the host and account for their kernels
            assert lib.cs_kernel_launch_count() == l0 and eng.graph_kernel_launches - g0 > 50
        assert (out != ref).float().mean().item() < 1e-4 and (c - cnt).abs().max().item() <= 1
    first = out.clone()
    d.copy_(torch.from_numpy(np.ascontiguousarray(imgs[::-1])).cuda())       # same buffer, new content → same graph
    out2, _ = eng(d)
    ref2, _ = eager(d)
    assert (out2 != ref2).float().mean().item() < 1e-4 and not torch.equal(out2, first)


def test_graph_mode_survives_other_geometries(sw, unet):
    """A captured call has its buffer addresses baked in: calls with other (larger) geometries in between must not disturb it."""
    from cellseg_b200 import synth
    small = torch.from_numpy(np.stack([synth.modality_image(300, 260, seed=s, modality=0) for s in range(2)])).cuda()
    big = torch.from_numpy(np.stack([synth.modality_image(700, 900, seed=s, modality=0) for s in range(3)])).cuda()
    eager = sw.SegmentationEngine(unet, chunk=16)
    ref_small, ref_big = eager(small)[0].clone(), eager(big)[0].clone()
    eng = sw.SegmentationEngine(unet, chunk=16, cuda_graph=True)
    for it in range(3):
        a = eng(small)[0].clone()
        b = eng(big)[0].clone()
        assert (a != ref_small).float().mean().item() < 1e-4 and (b != ref_big).float().mean().item() < 1e-4, it
    host = torch.from_numpy(np.stack([synth.modality_image(300, 260, seed=s, modality=0) for s in range(9)])).pin_memory()
    ref9 = eager(host.cuda())[0].cpu()
    for _ in range(2):                                   # ramped sub-batches (1, 3, 4, 1 images): four graphs, two slots
        out, _ = eng.segment_host(host, sub_batch=4)
        assert (out != ref9).float().mean().item() < 1e-4
