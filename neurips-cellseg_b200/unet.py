"""monai-free constructor of the predictor the reference builds at baseline/predict.py:45-52:

    monai.networks.nets.UNet(spatial_dims=2, in_channels=3, out_channels=num_class,
                             channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2), num_res_units=2)

The convolutions are the only dense contraction on the path and stay on PyTorch/cuDNN (out of scope
for hand-written kernels, BASELINE.json north_star).  This module only rebuilds the same topology with
MONAI-compatible `state_dict` keys so a reference checkpoint
(`torch.load(.../best_Dice_model.pth)['model_state_dict']`, predict.py:80-81) loads unchanged:
MONAI defaults kernel 3, up-kernel 3, PReLU, InstanceNorm(affine=False), no dropout, bias, ordering "NDA".
"""
from __future__ import annotations

import torch
from torch import nn


class _ADN(nn.Sequential):
    """Norm → (Dropout: absent at p=0) → Activation, MONAI key names "N" / "A"."""

    def __init__(self, channels: int):
        super().__init__()
        self.add_module("N", nn.InstanceNorm2d(channels))
        self.add_module("A", nn.PReLU())


class _Convolution(nn.Sequential):
    """MONAI `Convolution`: child "conv" (+ child "adn" unless conv_only)."""

    def __init__(self, cin: int, cout: int, stride: int = 1, kernel: int = 3, transposed: bool = False,
                 conv_only: bool = False):
        super().__init__()
        pad = (kernel - 1) // 2
        if transposed:
            conv = nn.ConvTranspose2d(cin, cout, kernel, stride, pad, output_padding=stride - 1, bias=True)
        else:
            conv = nn.Conv2d(cin, cout, kernel, stride, pad, bias=True)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", _ADN(cout))


class _ResidualUnit(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int = 1, subunits: int = 2, last_conv_only: bool = False):
        super().__init__()
        self.conv = nn.Sequential()
        sc, ss = cin, stride
        for su in range(max(1, subunits)):
            only = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}", _Convolution(sc, cout, ss, conv_only=only))
            sc, ss = cout, 1
        if stride != 1 or cin != cout:
            k, p = (3, 1) if stride != 1 else (1, 0)
            self.residual = nn.Conv2d(cin, cout, k, stride, p, bias=True)
        else:
            self.residual = nn.Identity()

    def forward(self, x):
        return self.conv(x) + self.residual(x)


class _SkipConnection(nn.Module):
    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.submodule = submodule

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], dim=1)


class UNet(nn.Module):
    """Residual UNet with the reference's hyper-parameters as defaults."""

    def __init__(self, spatial_dims: int = 2, in_channels: int = 3, out_channels: int = 3,
                 channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2), num_res_units: int = 2):
        super().__init__()
        if spatial_dims != 2:
            raise ValueError("the reference path is 2-D")
        if num_res_units < 1:
            raise ValueError("only the residual variant used by the reference is provided")
        self.num_res_units = num_res_units

        def block(inc, outc, ch, st, is_top):
            c, s = ch[0], st[0]
            if len(ch) > 2:
                sub = block(c, c, ch[1:], st[1:], False)
                upc = c * 2
            else:
                sub = _ResidualUnit(c, ch[1], 1, num_res_units)
                upc = c + ch[1]
            down = _ResidualUnit(inc, c, s, num_res_units)
            up = nn.Sequential(_Convolution(upc, outc, s, transposed=True),
                               _ResidualUnit(outc, outc, 1, 1, last_conv_only=is_top))
            return nn.Sequential(down, _SkipConnection(sub), up)

        self.model = block(in_channels, out_channels, tuple(channels), tuple(strides), True)

    def forward(self, x):
        return self.model(x)


def build_unet(num_class: int = 3, seed: int | None = None) -> UNet:
    """The reference's `--model_name unet` predictor; `seed` gives reproducible random-init weights
    (no checkpoints are available offline)."""
    if seed is not None:
        torch.manual_seed(seed)
    return UNet(2, 3, num_class)
