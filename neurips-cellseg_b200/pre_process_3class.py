"""Drop-in for the reference's data/pre_process_3class.py (training-data preparation): same CLI (-i/-o,
sub-folders images/ and labels/, `<name>_label.tiff` ground truth, PNG outputs), same functions.

  normalize_channel / channel fix-up (:22-29, :83-94)   → predict.preprocess_image_device (histogram + LUT kernels, numpy-exact percentiles)
  create_interior_map (:31-56)                          → cs_interior_boundary (inner boundaries dilated by disk(1), interior candidates),
                                                          cs_remove_small (16 pixels, 4-connected), cs_interior_combine → {0, 1, 2}

There is no CPU path: the functions need the CUDA library and a device."""
from __future__ import annotations

import argparse
import ctypes as C
import os
from os.path import join

import numpy as np

from . import capi


def create_interior_map(inst_map):
    """pre_process_3class.py:31-56.  inst_map: (H, W) integer instance map as numpy array or CUDA tensor (the reference passes
    `gt.astype(np.int16)`; ids are compared as int16 here as well).  Returns uint8 (H, W): 0 background, 1 interior, 2 boundary —
    numpy in → numpy out, CUDA tensor in → CUDA tensor out."""
    import torch
    from . import postprocess
    lib = capi.load()
    if not torch.cuda.is_available():
        raise capi.CellSegError("create_interior_map needs a CUDA device; cellseg_b200 has no CPU path")
    as_numpy = not isinstance(inst_map, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(inst_map).astype(np.int32)).cuda() if as_numpy else inst_map.to(torch.int32).contiguous()
    if t.dim() != 2:
        raise ValueError("instance map must be 2-D")
    H, W = t.shape
    st = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    boundary = torch.empty((H, W), dtype=torch.uint8, device=t.device)
    interior = torch.empty((H, W), dtype=torch.uint8, device=t.device)
    capi.check(lib, lib.cs_interior_boundary(t.data_ptr(), 1, H, W, boundary.data_ptr(), interior.data_ptr(), st))
    kept = postprocess.remove_small_objects(interior.view(torch.bool), min_size=16)
    out = torch.empty((H, W), dtype=torch.uint8, device=t.device)
    capi.check(lib, lib.cs_interior_combine(boundary.data_ptr(), kept.view(torch.uint8).data_ptr(), H * W, out.data_ptr(), st))
    return out.cpu().numpy() if as_numpy else out


def main(argv=None):
    parser = argparse.ArgumentParser('Preprocessing for microscopy image segmentation', add_help=False)
    parser.add_argument('-i', '--input_path', default='./data/Train_Labeled', type=str, help='training data path; subfolders: images, labels')
    parser.add_argument("-o", '--output_path', default='./data/Train_Pre_3class', type=str, help='preprocessing data path')
    args = parser.parse_args(argv)

    import torch
    from PIL import Image
    from . import tiffio
    from .predict import preprocess_image_device
    img_path, gt_path = join(args.input_path, 'images'), join(args.input_path, 'labels')
    img_names = sorted(os.listdir(img_path))
    gt_names = [img_name.split('.')[0] + '_label.tiff' for img_name in img_names]
    os.makedirs(join(args.output_path, 'images'), exist_ok=True)
    os.makedirs(join(args.output_path, 'labels'), exist_ok=True)
    device = torch.device("cuda")
    for img_name, gt_name in zip(img_names, gt_names):
        img_data = tiffio.imread(join(img_path, img_name))
        gt_data = tiffio.imread(join(gt_path, gt_name))
        pre_img = preprocess_image_device(img_data, device).cpu().numpy()                 # :83-94
        interior_map = create_interior_map(gt_data.astype(np.int16))                      # :97-98
        Image.fromarray(pre_img.astype(np.uint8)).save(join(args.output_path, 'images', img_name.split('.')[0] + '.png'))
        Image.fromarray(interior_map.astype(np.uint8)).save(join(args.output_path, 'labels', gt_name.split('.')[0] + '.png'))


if __name__ == "__main__":
    main()
