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


# Fused InstanceNorm → PReLU (→ + residual) through libcellseg (csrc/norm.cu) for CUDA float32 inference tensors;
# the plain PyTorch modules otherwise (CPU: the oracle's predictor; training; other dtypes).  Same state_dict either way.
FUSED_ADN = True


def _fused_ok(x: torch.Tensor, adn) -> bool:
    return (FUSED_ADN and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and x.dim() == 4
            and (x.shape[2] * x.shape[3]) % 4 == 0 and adn.A.weight.numel() == 1 and not adn.N.affine
            and not adn.N.track_running_stats)


class _ADN(nn.Sequential):
    """Norm → (Dropout: absent at p=0) → Activation, MONAI key names "N" / "A"."""

    def __init__(self, channels: int):
        super().__init__()
        self.add_module("N", nn.InstanceNorm2d(channels))
        self.add_module("A", nn.PReLU())

    def forward(self, x, residual=None, residual_bias=None):
        if _fused_ok(x, self):
            import ctypes as C
            from . import capi
            lib = capi.load()
            x = x.contiguous()
            res = None if residual is None else residual.contiguous()
            n, c, h, w = x.shape
            capi.check(lib, lib.cs_instnorm_prelu(x.data_ptr(), None if res is None else res.data_ptr(), x.data_ptr(),
                                                  self.A.weight.data_ptr(), None if residual_bias is None else residual_bias.data_ptr(), c,
                                                  n * c, h * w, float(self.N.eps),
                                                  C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
            return x                       # in place: the convolution output is not used elsewhere
        y = self.A(self.N(x))
        if residual is None:
            return y
        return y + residual if residual_bias is None else y + (residual + residual_bias.view(1, -1, 1, 1))


def _conv_nobias(conv, x):
    """The convolution without its bias term.  Used (fused mode only) where an InstanceNorm without affine follows —
    a per-channel constant is removed by the mean subtraction — and for the residual convolution, whose bias is added
    inside the fused kernel instead of by a separate broadcast pass."""
    if isinstance(conv, nn.ConvTranspose2d):
        return torch.nn.functional.conv_transpose2d(x, conv.weight, None, conv.stride, conv.padding, conv.output_padding,
                                                    conv.groups, conv.dilation)
    return torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)


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
        self.conv_only = conv_only
        if not conv_only:
            self.add_module("adn", _ADN(cout))

    def forward(self, x, residual=None, residual_bias=None):
        if self.conv_only:
            x = self.conv(x)
            if residual is None:
                return x
            return x + residual if residual_bias is None else x + (residual + residual_bias.view(1, -1, 1, 1))
        x = _conv_nobias(self.conv, x) if _fused_ok(x, self.adn) else self.conv(x)
        return self.adn(x, residual, residual_bias)


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
        units = list(self.conv)
        last = units[-1]
        rb = None
        if isinstance(self.residual, nn.Conv2d) and not last.conv_only and _fused_ok(x, last.adn):
            res, rb = _conv_nobias(self.residual, x), self.residual.bias     # bias added inside the fused kernel
        else:
            res = self.residual(x)
        cx = x
        for u in units[:-1]:
            cx = u(cx)
        return last(cx, res, rb)           # the residual add rides on the last unit's norm/activation pass


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
