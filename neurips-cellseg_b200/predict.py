"""Drop-in for the reference's baseline/predict.py: same CLI (-i/-o/--model_path/--show_overlay/--model_name/
--num_class/--input_size, predict.py:24-35), same per-image pipeline and outputs (`<name>_label.tiff`, int32,
deflate; predict.py:87-122) — with predict.py:107-113 (scale → sliding window → softmax → threshold → hole
filling → small-object removal → labelling) executed by libcellseg_sm100.so around the PyTorch/cuDNN UNet.

Additive flags (reference defaults unchanged): --blend {constant,gaussian}, --chunk (windows per forward;
the reference's sw_batch_size=4 is a memory knob — per-sample InstanceNorm makes window results independent
of batch composition), --random_init SEED (no checkpoints are available offline).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np

join = os.path.join


def normalize_channel(img, lower=1, upper=99):
    """predict.py:14-21 with skimage.exposure.rescale_intensity(in_range=(p_lo, p_hi), out_range='uint8')
    written out: clip to the percentile range, scale to [0, 255] in float64, truncate to uint8."""
    non_zero_vals = img[np.nonzero(img)]
    percentiles = np.percentile(non_zero_vals, [lower, upper])
    if percentiles[1] - percentiles[0] > 0.001:
        lo, hi = float(percentiles[0]), float(percentiles[1])
        x = np.clip(img, lo, hi)
        x = (x - lo) / (hi - lo)
        img_norm = np.asarray(x * 255.0 + 0.0)
    else:
        img_norm = img
    return img_norm.astype(np.uint8)


def preprocess_image(img_data: np.ndarray):
    """predict.py:94-104: channel fix-up and per-channel percentile normalisation → (img_data, u8[H,W,3])."""
    if len(img_data.shape) == 2:
        img_data = np.repeat(np.expand_dims(img_data, axis=-1), 3, axis=-1)
    elif len(img_data.shape) == 3 and img_data.shape[-1] > 3:
        img_data = img_data[:, :, :3]
    pre_img_data = np.zeros(img_data.shape, dtype=np.uint8)
    for i in range(3):
        img_channel_i = img_data[:, :, i]
        if len(img_channel_i[np.nonzero(img_channel_i)]) > 0:
            pre_img_data[:, :, i] = normalize_channel(img_channel_i, lower=1, upper=99)
    return img_data, pre_img_data


def percentiles_from_histogram(counts: np.ndarray, qs=(1, 99)):
    """np.percentile(values[values != 0], qs) (numpy's default "linear" method, float64) from a histogram of the
    values: counts[v] = number of pixels equal to v.  None when every pixel is zero."""
    cnt = counts.astype(np.int64).copy()
    cnt[0] = 0
    n = int(cnt.sum())
    if n == 0:
        return None
    cum = np.cumsum(cnt)
    out = []
    for q in qs:
        quantile = np.true_divide(q, 100)
        virtual = (n - 1) * quantile
        prev = np.floor(virtual)
        gamma = virtual - prev
        prev_i = int(prev)
        next_i = min(prev_i + 1, n - 1)
        a = float(np.searchsorted(cum, prev_i, side="right"))      # value of the prev_i-th order statistic
        b = float(np.searchsorted(cum, next_i, side="right"))
        diff = b - a
        r = a + diff * gamma                                        # numpy's _lerp
        if gamma >= 0.5:
            r = b - diff * (1 - gamma)
        out.append(r)
    return np.array(out, dtype=np.float64)


def prenorm_lut(hist: np.ndarray, dtype) -> np.ndarray:
    """Value → uint8 tables [3][bins] reproducing normalize_channel (predict.py:14-21) for every possible input value,
    given the per-channel histograms."""
    bins = hist.shape[1]
    values = np.arange(bins, dtype=np.int64).astype(dtype)
    lut = np.zeros((3, bins), dtype=np.uint8)
    for c in range(3):
        p = percentiles_from_histogram(hist[c])
        if p is None:
            continue                                                # channel without a non-zero pixel stays zero (:103)
        if p[1] - p[0] > 0.001:
            lo, hi = float(p[0]), float(p[1])
            x = np.clip(values, lo, hi)
            x = (x - lo) / (hi - lo)
            lut[c] = np.asarray(x * 255.0 + 0.0).astype(np.uint8)
        else:
            lut[c] = values.astype(np.uint8)
    return lut


def preprocess_image_device(img_data: np.ndarray, device="cuda"):
    """predict.py:94-104 with the per-pixel passes on the GPU (histogram + LUT kernels of libcellseg): raw image as read
    from disk → u8 [H,W,3] CUDA tensor.  uint8 and uint16 images; anything else takes the numpy path."""
    import ctypes as C
    import torch
    from . import capi
    if img_data.dtype not in (np.uint8, np.uint16) or img_data.ndim not in (2, 3):
        return torch.from_numpy(np.ascontiguousarray(preprocess_image(img_data)[1])).to(device)
    lib = capi.load()
    H, W = img_data.shape[:2]
    ch = 1 if img_data.ndim == 2 else img_data.shape[2]
    if img_data.ndim == 3 and ch < 3:
        raise IndexError(f"index 2 is out of bounds for axis 2 with size {ch}")     # what predict.py:101 raises
    eb = img_data.dtype.itemsize
    bins = 256 if eb == 1 else 65536
    raw = np.ascontiguousarray(img_data)
    src = torch.from_numpy(raw.view(np.uint8)).to(device, non_blocking=True)
    st = C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
    hist = torch.empty((3, bins), dtype=torch.int32, device=src.device)
    capi.check(lib, lib.cs_prenorm_hist(src.data_ptr(), eb, H, W, ch, hist.data_ptr(), st))
    h = hist.cpu().numpy().view(np.uint32)
    if ch == 1:
        h[1] = h[0]; h[2] = h[0]                                     # the replicated channels share the statistics
    lut = torch.from_numpy(prenorm_lut(h, img_data.dtype)).to(src.device)
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=src.device)
    capi.check(lib, lib.cs_prenorm_apply(src.data_ptr(), eb, H, W, ch, lut.data_ptr(), out.data_ptr(), st))
    return out


def build_model(model_name: str, num_class: int, input_size: int):
    """predict.py:44-78.  `unet` is rebuilt monai-free (cellseg_b200.unet); the transformer predictors are dense
    models outside the accelerated path and are constructed through monai when it is installed."""
    name = model_name.lower()
    if name == "unet":
        from .unet import UNet
        return UNet(spatial_dims=2, in_channels=3, out_channels=num_class, channels=(16, 32, 64, 128, 256),
                    strides=(2, 2, 2, 2), num_res_units=2)
    if name in ("unetr", "swinunetr"):
        try:
            import monai
        except ImportError as e:
            raise RuntimeError(f"--model_name {model_name} needs monai (not installed); --model_name unet is built in") from e
        if name == "swinunetr":
            return monai.networks.nets.SwinUNETR(img_size=(input_size, input_size), in_channels=3, out_channels=num_class,
                                                 feature_size=24, spatial_dims=2)
        try:
            from baseline.models.unetr2d import UNETR2D         # the reference's own module (baseline/models/unetr2d.py)
        except ImportError as e:
            raise RuntimeError("--model_name unetr needs the reference checkout on PYTHONPATH (baseline.models.unetr2d, a ViT "
                               "built from monai blocks) — dense models are outside this package; --model_name unet is built in") from e
        return UNETR2D(in_channels=3, out_channels=num_class, img_size=(input_size, input_size), feature_size=16,
                       hidden_size=768, mlp_dim=3072, num_heads=12, pos_embed="perceptron", norm_name="instance",
                       res_block=True, dropout_rate=0.0)
    raise ValueError(f"unknown --model_name {model_name!r} (unet, unetr, swinunetr)")


def overlay_boundaries(img_data: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """predict.py:118-122 (visualisation only, host side): inner boundaries at connectivity 1, dilated by disk(2)."""
    import scipy.ndimage as ndi
    lab = labels
    pad = np.pad(lab, 1, mode="edge")
    differs = np.zeros(lab.shape, dtype=bool)
    for dy, dx in ((0, 1), (2, 1), (1, 0), (1, 2)):
        differs |= pad[dy:dy + lab.shape[0], dx:dx + lab.shape[1]] != lab
    boundary = differs & (lab != 0)
    yy, xx = np.mgrid[-2:3, -2:3]
    boundary = ndi.binary_dilation(boundary, structure=(yy * yy + xx * xx) <= 4)
    img_data[boundary, :] = 255
    return img_data


def main(argv=None):
    parser = argparse.ArgumentParser('Baseline for Microscopy image segmentation', add_help=False)
    parser.add_argument('-i', '--input_path', default='./inputs', type=str, help='training data path; subfolders: images, labels')
    parser.add_argument("-o", '--output_path', default='./outputs', type=str, help='output path')
    parser.add_argument('--model_path', default='./work_dir/swinunetr_3class', help='path where to save models and segmentation results')
    parser.add_argument('--show_overlay', required=False, default=False, action="store_true", help='save segmentation overlay')
    parser.add_argument('--model_name', default='swinunetr', help='select mode: unet (built in), unetr, swinunetr (both need monai; unetr also the reference checkout)')
    parser.add_argument('--num_class', default=3, type=int, help='segmentation classes')
    parser.add_argument('--input_size', default=256, type=int, help='segmentation classes')
    # additive
    parser.add_argument('--blend', default='constant', choices=('constant', 'gaussian'), help='importance map (MONAI mode)')
    parser.add_argument('--chunk', default=64, type=int, help='windows per predictor forward')
    parser.add_argument('--random_init', default=None, type=int, help='seed: use random-init weights instead of a checkpoint')
    args = parser.parse_args(argv)

    import torch
    from . import capi, tiffio
    from .sliding_window import SegmentationEngine

    input_path, output_path = args.input_path, args.output_path
    os.makedirs(output_path, exist_ok=True)
    img_names = sorted(os.listdir(join(input_path)))

    if not torch.cuda.is_available():
        raise capi.CellSegError("cellseg_b200 predict needs a CUDA device (sm_100a); there is no CPU path")
    capi.load()
    # under torchrun the input files are dealt out over the ranks, one GPU each: images are independent units, so the
    # multi-GPU form of this CLI needs no collective at all
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    device = torch.device("cuda")
    if world > 1:
        device = torch.device("cuda", local % torch.cuda.device_count())
        torch.cuda.set_device(device)
    img_names = img_names[rank::world]
    if args.random_init is not None:
        torch.manual_seed(args.random_init)
    model = build_model(args.model_name, args.num_class, args.input_size).to(device)
    if args.random_init is None:
        checkpoint = torch.load(join(args.model_path, 'best_Dice_model.pth'), map_location=device)
        model.load_state_dict(checkpoint['model_state_dict'])
    model.eval()
    engine = SegmentationEngine(model, roi=args.input_size, mode=args.blend, chunk=args.chunk)
    # File I/O runs beside the GPU: the next image is decoded while the current one is segmented, and label TIFFs are deflated and
    # written by a small pool (the reference does all three in sequence and counts the write in its per-image time, predict.py:106-116).
    from concurrent.futures import ThreadPoolExecutor
    reader, writer = ThreadPoolExecutor(1), ThreadPoolExecutor(2)
    pending = []
    nxt = reader.submit(tiffio.imread, join(input_path, img_names[0])) if img_names else None
    try:
        with torch.no_grad():
            for i, img_name in enumerate(img_names):
                img_data = nxt.result()
                nxt = reader.submit(tiffio.imread, join(input_path, img_names[i + 1])) if i + 1 < len(img_names) else None
                pre_img_dev = preprocess_image_device(img_data, device)        # predict.py:94-104 on the GPU
                if img_data.ndim == 2:
                    img_data = np.repeat(np.expand_dims(img_data, axis=-1), 3, axis=-1)
                elif img_data.ndim == 3 and img_data.shape[-1] > 3:
                    img_data = img_data[:, :, :3]

                t0 = time.time()
                labels, _ = engine(pre_img_dev[None])
                test_pred_mask = labels[0].cpu().numpy()                        # a copy: the engine reuses its label buffer
                pending.append(writer.submit(tiffio.imwrite_labels, join(output_path, img_name.split('.')[0] + '_label.tiff'), test_pred_mask))
                t1 = time.time()
                print(f'Prediction finished: {img_name}; img size = {tuple(pre_img_dev.shape)}; costing: {t1-t0:.2f}s')

                if args.show_overlay:
                    pending.append(writer.submit(lambda p, a, b: tiffio.imsave(p, overlay_boundaries(a, b)),
                                                 join(output_path, 'overlay_' + img_name), img_data, test_pred_mask))
                while len(pending) > 8:                                         # bound the label maps waiting in host memory
                    pending.pop(0).result()
        for f in pending:
            f.result()                                                          # surfaces write errors
    finally:
        reader.shutdown(wait=False, cancel_futures=True)
        writer.shutdown(wait=True)


if __name__ == "__main__":
    main()
