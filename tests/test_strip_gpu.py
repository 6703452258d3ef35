"""GPU tier for the strip-sharded whole-slide path (BASELINE config 3): all strips of an image driven in lockstep on
one GPU must reproduce the single-image result bit for bit (the ranks of a multi-GPU job run exactly these phases,
with the rows moved by NCCL instead of in-process — tools/wsi_check.py does that under torchrun)."""
import numpy as np
import pytest
import torch

from oracle import morph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wsi():
    assert torch.cuda.is_available()
    from cellseg_b200 import wsi
    return wsi


def _lockstep(wsi, mask, n, **kw):
    H, W = mask.shape
    rows = wsi.strip_rows(H, n)
    strips = [wsi.StripPost(r1 - r0, W, "cuda", **kw) for r0, r1 in rows]
    d = torch.from_numpy(np.ascontiguousarray(mask)).cuda()
    labels, total = wsi.run_lockstep(strips, [d[r0:r1] for r0, r1 in rows])
    return torch.cat(labels), total


@pytest.mark.parametrize("n", [2, 3, 8])
def test_strips_equal_oracle(wsi, n):
    from cellseg_b200 import synth
    for mask in (synth.blob_mask(1500, 2100, 3), synth.noise_mask(1100, 900, 0.5, 4), synth.blob_mask(900, 1300, 5, sigma=2.0),
                 ~synth.blob_mask(1000, 1000, 6, sigma=3.0), synth.spiral_mask(700, 520), synth.comb_mask(640, 300, True)):
        ref = morph.postprocess_mask(mask)
        got, total = _lockstep(wsi, mask, n)
        assert np.array_equal(got.cpu().numpy(), ref)
        assert total == ref.max()


def test_full_size_wsi_8_strips_equals_single_gpu(wsi):
    from cellseg_b200 import postprocess, synth
    mask = synth.blob_mask(10000, 10000, 0)
    single, n = postprocess.postprocess_device(torch.from_numpy(mask).cuda())
    got, total = _lockstep(wsi, mask, 8)
    assert total == int(n[0])
    assert bool((got == single).all())


@pytest.mark.parametrize("H,W,n,mode", [(1500, 1100, 4, "constant"), (1000, 700, 3, "gaussian"), (2048, 640, 8, "constant")])
def test_strip_segmenter_equals_engine(wsi, H, W, n, mode):
    """u8 image → instance map: strips (global window grid, per-strip blend rows, merged labels) == whole image."""
    from cellseg_b200 import sliding_window as sw
    from cellseg_b200 import synth
    from cellseg_b200.unet import build_unet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    model = build_unet(3, seed=2022).eval().cuda()
    img = torch.from_numpy(synth.modality_image(H, W, seed=1, modality=0)).cuda()
    eng = sw.SegmentationEngine(model, chunk=7, mode=mode)
    ref, cnt = eng(img[None])
    ref = ref[0].clone()
    # windows must be predicted in the same batches for bit-identical logits → chunk = one window row per forward
    segs = [wsi.StripSegmenter(model, H, W, r, n, chunk=1, mode=mode, cuda_graph=False) for r in range(n)]
    eng1 = sw.SegmentationEngine(model, chunk=1, mode=mode)
    ref1, cnt1 = eng1(img[None])
    got, total = wsi.segment_lockstep(segs, img)
    assert total == int(cnt1[0])
    assert bool((got == ref1[0]).all())
    # and against the chunked engine wherever cuDNN gave identical logits (it normally does for equal batch sizes)
    assert (got != ref).float().mean().item() < 1e-3


@pytest.mark.parametrize("peak", [255, 200])
def test_segment_host_streams_the_upload_and_equals_the_plain_call(wsi, peak):
    """StripSegmenter.segment_host uploads the rows in pieces behind the predictor, guessing that the image maximum is 255 (what the
    percentile normalisation ahead of predict.py:107 produces); the guess is verified, a wrong one redoes the step — either way the
    labels equal those of the plain device-resident call."""
    from cellseg_b200 import synth
    from cellseg_b200.unet import build_unet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    H, W = 1300, 900
    model = build_unet(3, seed=2022).eval().cuda()
    img = synth.modality_image(H, W, seed=3, modality=0).copy()
    img = (img.astype(np.float32) * (peak / max(1, int(img.max())))).astype(np.uint8)
    img[5, 7, 1] = peak
    assert int(img.max()) == peak
    seg = wsi.StripSegmenter(model, H, W, 0, 1, chunk=7, cuda_graph=False)
    s0, s1 = seg.source_rows()
    host = torch.from_numpy(np.ascontiguousarray(img[s0:s1])).pin_memory()
    ref, total_ref = seg(host.cuda())
    ref = ref.clone()
    for chunks in (1, 3, 8):
        out, total = seg.segment_host(host, copy_chunks=chunks)
        assert total == total_ref and bool((out.cuda() == ref).all())
