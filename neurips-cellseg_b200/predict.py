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
        from baseline.models.unetr2d import UNETR2D
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
    parser.add_argument('--model_name', default='swinunetr', help='select mode: unet, unetr, swinunetr')
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
    device = torch.device("cuda")
    if args.random_init is not None:
        torch.manual_seed(args.random_init)
    model = build_model(args.model_name, args.num_class, args.input_size).to(device)
    if args.random_init is None:
        checkpoint = torch.load(join(args.model_path, 'best_Dice_model.pth'), map_location=device)
        model.load_state_dict(checkpoint['model_state_dict'])
    model.eval()
    engine = SegmentationEngine(model, roi=args.input_size, mode=args.blend, chunk=args.chunk)
    with torch.no_grad():
        for img_name in img_names:
            img_data = tiffio.imread(join(input_path, img_name))
            img_data, pre_img_data = preprocess_image(img_data)

            t0 = time.time()
            labels, _ = engine(torch.from_numpy(np.ascontiguousarray(pre_img_data[None])).to(device))
            test_pred_mask = labels[0].cpu().numpy()
            tiffio.imwrite_labels(join(output_path, img_name.split('.')[0] + '_label.tiff'), test_pred_mask)
            t1 = time.time()
            print(f'Prediction finished: {img_name}; img size = {pre_img_data.shape}; costing: {t1-t0:.2f}s')

            if args.show_overlay:
                tiffio.imsave(join(output_path, 'overlay_' + img_name), overlay_boundaries(img_data, test_pred_mask))


if __name__ == "__main__":
    main()
