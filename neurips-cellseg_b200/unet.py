"""monai-free constructor of the predictor the reference builds at baseline/predict.py:45-52:

    monai.networks.nets.UNet(spatial_dims=2, in_channels=3, out_channels=num_class,
                             channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2), num_res_units=2)

The dense convolutions are the only contraction on the path and stay on PyTorch/cuDNN (out of scope for hand-written
kernels, BASELINE.json north_star).  This module rebuilds the same topology with MONAI-compatible `state_dict` keys so a
reference checkpoint (`torch.load(.../best_Dice_model.pth)['model_state_dict']`, predict.py:80-81) loads unchanged
(MONAI defaults kernel 3, up-kernel 3, PReLU, InstanceNorm(affine=False), no dropout, bias, ordering "NDA"), and runs it for
CUDA float32 inference through `ChannelsLastRunner`: cuDNN convolutions on NHWC tensors, libcellseg kernels (csrc/unet_glue.cu)
for the 3-channel stem/head convolutions, every norm block, the skip concatenations and the window gather.
On the CPU, in training, or for other topologies the plain module graph runs.
"""
from __future__ import annotations

import torch
from torch import nn


# Fused InstanceNorm → PReLU (→ + residual) through libcellseg (csrc/norm.cu) for CUDA float32 inference tensors;
# the plain PyTorch modules otherwise (CPU: the oracle's predictor; training; other dtypes).  Same state_dict either way.
FUSED_ADN = True
# Channels-last inference forward on CUDA (ChannelsLastRunner below): cuDNN convolutions on NHWC tensors, libcellseg kernels
# for the 3-channel stem/head convolutions and every norm block.  False → the module graph (NCHW) runs instead.
FAST_NHWC = True


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
        self._runner = None                # channels-last inference runner, built on first CUDA inference call

    def runner(self):
        """The channels-last CUDA inference runner of this network, or None when FAST_NHWC is off / the topology is not the
        reference's (then the module graph runs)."""
        if not FAST_NHWC:
            return None
        if self._runner is None:
            self._runner = ChannelsLastRunner(self) if ChannelsLastRunner.supports(self) else False
        return self._runner or None

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and x.dim() == 4:
            r = self.runner()
            if r is not None and x.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
                return r(x)
        return self.model(x)


def build_unet(num_class: int = 3, seed: int | None = None) -> UNet:
    """The reference's `--model_name unet` predictor; `seed` gives reproducible random-init weights
    (no checkpoints are available offline)."""
    if seed is not None:
        torch.manual_seed(seed)
    return UNet(2, 3, num_class)


def convt_as_conv2x2_weight(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k3, s2, p1, output_padding 1) weight [Cin, Cout, 3, 3] → weight [4*Cout, Cin, 2, 2] of the regular
    convolution (cross-correlation, symmetric padding 1) that computes the same thing with one output group per output parity:
    y[2i+dy, 2j+dx, co] = conv2d(x, W2, padding=1)[i+1, j+1, (dy*2+dx)*Cout + co].  Derivation: output row 2i takes x[i]·w[ky=1],
    row 2i+1 takes x[i]·w[ky=2] + x[i+1]·w[ky=0]; likewise along columns."""
    cin, cout = w.shape[0], w.shape[1]
    wt = w.transpose(0, 1)                                  # [Cout, Cin, ky, kx]
    w2 = torch.zeros((4, cout, cin, 2, 2), dtype=w.dtype, device=w.device)
    taps = {0: ((0, 1),), 1: ((0, 2), (1, 0))}              # parity → ((input offset, kernel index), ...)
    for dy in (0, 1):
        for dx in (0, 1):
            for oy, ky in taps[dy]:
                for ox, kx in taps[dx]:
                    w2[dy * 2 + dx, :, :, oy, ox] = wt[:, :, ky, kx]
    return w2.reshape(4 * cout, cin, 2, 2)


class ChannelsLastRunner:
    """Inference forward of `UNet` (reference hyper-parameters) on CUDA float32 with every activation channels-last:

        windows NCHW ── cs_unet_stem (both stride-2 3→16 convolutions of model.0 + statistics) ── cs_nhwc_norm_prelu ──
        cuDNN convolutions on NHWC tensors, each followed by cs_nhwc_stats + cs_nhwc_norm_prelu (InstanceNorm → PReLU
        → + residual + residual-conv bias in one pass) ── cs_unet_head_convt (ConvTranspose2d 32→3 straight from the two
        halves of the top skip connection) ── cs_unet_head_final (norm → PReLU → Conv2d 3→3 + identity residual) ── logits NCHW

    Same parameters / state_dict as the module graph; biases of convolutions that feed a non-affine InstanceNorm are
    dropped (the mean subtraction removes them exactly).  The dense convolutions stay on cuDNN (north star)."""

    @staticmethod
    def supports(net: "UNet") -> bool:
        try:
            d0, skip0, up0 = net.model[0], net.model[1], net.model[2]
            c0, cr = d0.conv.unit0.conv, d0.residual
            head_t, head_c = up0[0].conv, up0[1].conv.unit0
            ok = (isinstance(cr, nn.Conv2d) and tuple(c0.weight.shape) == (16, 3, 3, 3) and tuple(cr.weight.shape) == (16, 3, 3, 3)
                  and c0.stride == (2, 2) and cr.stride == (2, 2) and len(d0.conv) == 2
                  and isinstance(head_t, nn.ConvTranspose2d) and tuple(head_t.weight.shape) == (32, 3, 3, 3) and head_t.stride == (2, 2)
                  and head_c.conv_only and tuple(head_c.conv.weight.shape) == (3, 3, 3, 3) and len(up0[1].conv) == 1
                  and isinstance(up0[1].residual, nn.Identity) and isinstance(skip0, _SkipConnection))
            for m in net.modules():
                if isinstance(m, _ADN):
                    ok = ok and m.A.weight.numel() == 1 and not m.N.affine and not m.N.track_running_stats
                    c = m.N.num_features
                    ok = ok and (c == 3 or (4 <= c <= 256 and c & (c - 1) == 0))
            return bool(ok)
        except (AttributeError, IndexError, TypeError):
            return False

    def __init__(self, net: "UNet"):
        self.net = net
        self._wcache = {}

    # -- helpers ----------------------------------------------------------------------------------------------------
    def _cl_weight(self, p: torch.Tensor) -> torch.Tensor:
        key = id(p)
        hit = self._wcache.get(key)
        if hit is None or hit[0] != (p._version, p.data_ptr()):
            hit = ((p._version, p.data_ptr()), p.detach().contiguous(memory_format=torch.channels_last))
            self._wcache[key] = hit
        return hit[1]

    def _host_weight(self, p: torch.Tensor):
        """Host copy (numpy, contiguous float32) of a small weight that travels as kernel parameters.  The copy is made on
        first use — outside any CUDA-graph capture, since it synchronises — and refreshed when the parameter changes."""
        key = ("host", id(p))
        hit = self._wcache.get(key)
        if hit is None or hit[0] != (p._version, p.data_ptr()):
            hit = ((p._version, p.data_ptr()), p.detach().to("cpu", torch.float32).contiguous().numpy())
            self._wcache[key] = hit
        return hit[1].ctypes.data

    def _conv(self, conv, x):
        w = self._cl_weight(conv.weight)
        if isinstance(conv, nn.ConvTranspose2d):
            y = torch.nn.functional.conv_transpose2d(x, w, None, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
        else:
            y = torch.nn.functional.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if not y.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous(memory_format=torch.channels_last)
        return y

    def _norm(self, x, adn, res=None, res_bias=None, partials=None, stat_px=None, cat=None, dense=True):
        """prelu(instance_norm(x)) (+ res + res_bias), in place on x (a convolution output nobody else reads) when `dense`;
        `cat` = (concat buffer [n, c_total, h, w] channels-last, channel offset) additionally receives the result in that slot,
        so the skip connection's torch.cat never runs.  Returns x (None when dense is False)."""
        import ctypes as C
        from . import capi
        lib = capi.load()
        n, c, h, w = x.shape
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        resp = None if res is None else res.data_ptr()
        rbp = None if res_bias is None else res_bias.data_ptr()
        if partials is None:
            ws = torch.empty((n, lib.cs_nhwc_stat_chunks(h * w, c), c, 2), dtype=torch.float32, device=x.device)
            y2, c2, off = (cat[0].data_ptr(), cat[0].shape[1], cat[1]) if cat is not None else (None, 0, 0)
            capi.check(lib, lib.cs_instnorm_prelu_nhwc(x.data_ptr(), resp, x.data_ptr() if dense else None, adn.A.weight.data_ptr(), rbp,
                                                       n, h * w, c, float(adn.N.eps), ws.data_ptr(), y2, c2, off, st))
        else:
            assert cat is None and dense
            capi.check(lib, lib.cs_nhwc_norm_prelu(x.data_ptr(), resp, x.data_ptr(), adn.A.weight.data_ptr(), rbp, n, h * w, c,
                                                   partials.data_ptr(), stat_px, float(adn.N.eps), st))
        return x if dense else None

    def _derived_weight(self, tag: str, params, make):
        """A weight tensor derived from parameters (merged / re-laid-out), cached until one of them changes."""
        key = (tag,) + tuple(id(p) for p in params)
        ver = tuple((p._version, p.data_ptr()) for p in params)
        hit = self._wcache.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, make().contiguous(memory_format=torch.channels_last))
            self._wcache[key] = hit
        return hit[1]

    def _norm_slices(self, x, x_off, c, adn, res=None, res_off=0, res_bias=None, cat=None, dense=True):
        """The norm block on channel slices: input = channels [x_off, x_off + c) of the channels-last tensor x, residual likewise.
        Returns a new dense channels-last tensor (None when dense is False)."""
        import ctypes as C
        from . import capi
        lib = capi.load()
        n, cx, h, w = x.shape
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last) if dense else None
        ws = torch.empty((n, lib.cs_nhwc_stat_chunks(h * w, c), c, 2), dtype=torch.float32, device=x.device)
        y2, c2, off = (cat[0].data_ptr(), cat[0].shape[1], cat[1]) if cat is not None else (None, 0, 0)
        capi.check(lib, lib.cs_instnorm_prelu_nhwc_slices(x.data_ptr(), cx, x_off, None if res is None else res.data_ptr(),
                                                          0 if res is None else res.shape[1], res_off, None if y is None else y.data_ptr(),
                                                          adn.A.weight.data_ptr(), None if res_bias is None else res_bias.data_ptr(),
                                                          n, h * w, c, float(adn.N.eps), ws.data_ptr(), y2, c2, off, st))
        return y

    def _res_unit(self, ru, x, cat=None, dense=True):
        units = list(ru.conv)
        first, last = units[0], units[-1]
        rc = ru.residual
        mergeable = (isinstance(rc, nn.Conv2d) and len(units) >= 2 and rc.kernel_size == first.conv.kernel_size and rc.stride == first.conv.stride
                     and rc.padding == first.conv.padding and rc.in_channels == first.conv.in_channels and rc.out_channels == first.conv.out_channels
                     and rc.out_channels % 4 == 0)
        if mergeable:
            # unit0 and the residual branch read the same input with the same geometry: ONE convolution with 2C outputs;
            # unit0's norm reads channels [0, C), the last norm takes its residual from channels [C, 2C)
            c = rc.out_channels
            wm = self._derived_weight("merge", (first.conv.weight, rc.weight), lambda: torch.cat([first.conv.weight.detach(), rc.weight.detach()], 0))
            m = torch.nn.functional.conv2d(x, wm, None, rc.stride, rc.padding)
            if not m.is_contiguous(memory_format=torch.channels_last):
                m = m.contiguous(memory_format=torch.channels_last)
            cx = self._norm_slices(m, 0, c, first.adn)
            for u in units[1:-1]:
                cx = self._norm(self._conv(u.conv, cx), u.adn)
            return self._norm_slices(self._conv(last.conv, cx), 0, c, last.adn, res=m, res_off=c, res_bias=rc.bias, cat=cat, dense=dense)
        if isinstance(rc, nn.Conv2d):
            res, rb = self._conv(rc, x), rc.bias
        else:
            res, rb = x, None
        cx = x
        for u in units[:-1]:
            cx = self._norm(self._conv(u.conv, cx), u.adn)
        return self._norm(self._conv(last.conv, cx), last.adn, res, rb, cat=cat, dense=dense)

    def _up_conv(self, conv, adn, x):
        """ConvTranspose2d → norm block.  k3/s2/p1/op1 with ≤ 64 outputs runs as one regular 2x2 convolution with 4C outputs
        (cuDNN's transposed kernels are 2.4x slower at 64 → 16 channels) whose depth-to-space shuffle happens inside the norm kernel."""
        import ctypes as C
        from . import capi
        c = conv.out_channels
        fast = (isinstance(conv, nn.ConvTranspose2d) and conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.padding == (1, 1)
                and conv.output_padding == (1, 1) and conv.groups == 1 and conv.dilation == (1, 1) and c <= 32 and c % 4 == 0 and c & (c - 1) == 0)
        if not fast:
            return self._norm(self._conv(conv, x), adn)
        lib = capi.load()
        w2 = self._derived_weight("convt2x2", (conv.weight,), lambda: convt_as_conv2x2_weight(conv.weight.detach()))
        m = torch.nn.functional.conv2d(x, w2, None, 1, 1)                       # [n, 4c, h+1, w+1]
        if not m.is_contiguous(memory_format=torch.channels_last):
            m = m.contiguous(memory_format=torch.channels_last)
        n, _, h, w = x.shape
        y = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        ws = torch.empty((n, lib.cs_nhwc_stat_chunks(h * w, 4 * c), 4 * c, 2), dtype=torch.float32, device=x.device)
        capi.check(lib, lib.cs_instnorm_prelu_nhwc_shuffle(m.data_ptr(), y.data_ptr(), adn.A.weight.data_ptr(), n, h, w, c, float(adn.N.eps), ws.data_ptr(),
                                                           C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return y

    def _block(self, seq, x, cat=None):
        """One UNet level below the top: down → skip(sub) → up.  Both halves of the skip concatenation are written into one
        buffer by the norm kernels that produce them.  `cat`: the enclosing level's concat slot for this block's result."""
        down, skip, up = seq[0], seq[1], seq[2]
        n, _, h, w = x.shape
        c1 = list(down.conv)[-1].conv.out_channels
        stride = list(down.conv)[0].conv.stride[0]
        sub = skip.submodule
        inner = isinstance(sub, nn.Sequential) and len(sub) == 3 and isinstance(sub[1], _SkipConnection)
        c2 = (list(sub[2][1].conv)[-1] if inner else list(sub.conv)[-1]).conv.out_channels
        both = torch.empty((n, c1 + c2, h // stride, w // stride), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        x1 = self._res_unit(down, x, cat=(both, 0))
        if inner:
            self._block(sub, x1, cat=(both, c1))
        else:
            self._res_unit(sub, x1, cat=(both, c1), dense=False)
        t = self._up_conv(up[0].conv, up[0].adn, both)
        return self._res_unit(up[1], t, cat=cat, dense=cat is None)

    # -- forward ----------------------------------------------------------------------------------------------------
    def _run(self, stem, k: int, H: int, W: int, dev, out):
        import ctypes as C
        from . import capi
        lib = capi.load()
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d0, skip0, up0 = self.net.model[0], self.net.model[1], self.net.model[2]
        oh, ow = H // 2, W // 2
        cl = dict(dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        a = torch.empty((k, 16, oh, ow), **cl)
        r = torch.empty((k, 16, oh, ow), **cl)
        pa = torch.empty((k, (oh * ow + 255) // 256, 16, 2), dtype=torch.float32, device=dev)
        stem(lib, a, r, pa, self._host_weight(d0.conv.unit0.conv.weight), self._host_weight(d0.residual.weight), st)
        u0 = self._norm(a, d0.conv.unit0.adn, partials=pa, stat_px=256)
        x0 = self._norm(self._conv(d0.conv.unit1.conv, u0), d0.conv.unit1.adn, r, d0.residual.bias)
        sub = skip0.submodule
        inner = isinstance(sub, nn.Sequential) and len(sub) == 3 and isinstance(sub[1], _SkipConnection)
        s = self._block(sub, x0) if inner else self._res_unit(sub, x0)
        z = torch.empty((k, 3, H, W), dtype=torch.float32, device=dev)
        pz = torch.empty((k, lib.cs_unet_head_partial_chunks(oh, ow), 3, 2), dtype=torch.float32, device=dev)
        capi.check(lib, lib.cs_unet_head_convt(x0.data_ptr(), s.data_ptr(), k, oh, ow, self._host_weight(up0[0].conv.weight), z.data_ptr(),
                                               pz.data_ptr(), st))
        if out is None:
            out = torch.empty((k, 3, H, W), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (k, 3, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
            raise capi.CellSegError("output buffer must be contiguous float32 [k, 3, H, W]")
        hc = up0[1].conv.unit0.conv
        capi.check(lib, lib.cs_unet_head_final(z.data_ptr(), pz.data_ptr(), k, H, W, up0[0].adn.A.weight.data_ptr(), self._host_weight(hc.weight),
                                               None if hc.bias is None else self._host_weight(hc.bias), float(up0[0].adn.N.eps), out.data_ptr(), st))
        return out

    @torch.no_grad()
    def __call__(self, win: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """win: float32 [k,3,H,W] window batch (NCHW) → logits [k,3,H,W] (written into `out` when given)."""
        from . import capi
        win = win.contiguous()
        k, _, H, W = win.shape

        def stem(lib, a, r, pa, w0, wr, st):
            capi.check(lib, lib.cs_unet_stem(win.data_ptr(), k, H, W, w0, wr, a.data_ptr(), r.data_ptr(), pa.data_ptr(), st))
        return self._run(stem, k, H, W, win.device, out)

    @torch.no_grad()
    def from_images(self, img_u8: torch.Tensor, lut: torch.Tensor, plan, begin: int, count: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Windows [begin, begin+count) of the sliding-window plan over uint8 images [N,H,W,3] → logits [count,3,roi,roi].
        The windows are never materialised: the first convolutions read the image bytes through the scale table `lut`."""
        ys, xs, _, _ = plan.tensors(img_u8.device)
        return self.from_window_grid(img_u8, lut, img_u8.shape[0], plan.H, plan.W, plan.roi, plan.ny, plan.nx, plan.pad_y, plan.pad_x,
                                     ys, xs, begin, count, out)

    @torch.no_grad()
    def from_window_grid(self, img_u8, lut, n_img: int, H: int, W: int, roi: int, ny: int, nx: int, pad_y: int, pad_x: int, ys, xs,
                         begin: int, count: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """from_images with the window grid spelled out (arguments of cs_gather_windows_u8): `img_u8` holds n_img images of H x W
        pixels, padded coordinate (Y, X) is image pixel (Y - pad_y, X - pad_x), pixels outside the image read as zero.  A row
        strip of a larger image is addressed with H = rows held and pad_y = first row held (wsi.py)."""
        from . import capi
        if img_u8.dtype != torch.uint8 or not img_u8.is_contiguous() or roi % 2:
            raise capi.CellSegError("expects contiguous uint8 HWC images and an even ROI")

        def stem(lib, a, r, pa, w0, wr, st):
            capi.check(lib, lib.cs_unet_stem_u8(img_u8.data_ptr(), lut.data_ptr(), n_img, H, W, roi, ny, nx, pad_y, pad_x,
                                                ys.data_ptr(), xs.data_ptr(), begin, count, w0, wr, a.data_ptr(), r.data_ptr(), pa.data_ptr(), st))
        return self._run(stem, count, roi, roi, img_u8.device, out)
