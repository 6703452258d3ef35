"""Host side of the sliding-window path (baseline/predict.py:107-113).

`sliding_window_inference` keeps the signature of `monai.inferers.sliding_window_inference` as the
reference calls it (predict.py:109; model_training_3class.py:291) and returns the blended predictor
output; `SegmentationEngine` is the fused image → instance-map path (u8 image in, int32 labels out) in
which window logits are read once and no float plane is written back to HBM.  Window-grid arithmetic
(padding, scan interval, shifted last window) follows MONAI's published algorithm;
all per-pixel work is in libcellseg_sm100.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import capi


# ---------------------------------------------------------------------------
# window grid (host integers)
# ---------------------------------------------------------------------------
def _axis_starts(size: int, roi: int, overlap: float) -> list[int]:
    """Window start offsets along one axis of the padded image (size >= roi)."""
    if size == roi:
        interval = roi
    else:
        interval = int(roi * (1 - overlap))
        interval = interval if interval > 0 else 1
    upper = int(math.ceil(float(size) / interval))
    count = 1
    for d in range(upper):
        if d * interval + roi >= size:
            count = d + 1
            break
    out = []
    for i in range(count):
        s = i * interval
        out.append(s - max(s + roi - size, 0))     # the last window is shifted back inside, not padded
    return out


def _cover_table(starts: list[int], roi: int, size: int) -> np.ndarray:
    """For every padded coordinate: (first covering window index << 8) | number of covering windows."""
    s = np.asarray(starts, dtype=np.int64)
    c = np.arange(size, dtype=np.int64)
    lo = np.searchsorted(s + roi, c, side="right")       # first window with start + roi > c
    hi = np.searchsorted(s, c, side="right")             # windows with start <= c
    cnt = hi - lo
    if cnt.min() < 1 or cnt.max() > 255:
        raise capi.CellSegError("window grid does not cover the image")
    return ((lo << 8) | cnt).astype(np.int32)


@dataclass
class WindowPlan:
    H: int
    W: int
    roi: int
    overlap: float
    pad_y: int = 0
    pad_x: int = 0
    ys: list = field(default_factory=list)
    xs: list = field(default_factory=list)
    dev: dict = field(default_factory=dict)

    @property
    def ny(self): return len(self.ys)

    @property
    def nx(self): return len(self.xs)

    @property
    def n_win(self): return len(self.ys) * len(self.xs)

    def tensors(self, device):
        key = str(device)
        if key not in self.dev:
            Hp, Wp = max(self.H, self.roi), max(self.W, self.roi)
            self.dev[key] = tuple(torch.from_numpy(np.asarray(a, dtype=np.int32)).to(device) for a in (
                self.ys, self.xs, _cover_table(self.ys, self.roi, Hp), _cover_table(self.xs, self.roi, Wp)))
        return self.dev[key]


_plans: dict = {}


def make_plan(H: int, W: int, roi: int = 256, overlap: float = 0.25) -> WindowPlan:
    key = (H, W, roi, overlap)
    if key not in _plans:
        p = WindowPlan(H, W, roi, overlap)
        dy, dx = max(roi - H, 0), max(roi - W, 0)
        p.pad_y, p.pad_x = dy // 2, dx // 2           # low side gets the floor, as MONAI pads
        p.ys = _axis_starts(max(H, roi), roi, overlap)
        p.xs = _axis_starts(max(W, roi), roi, overlap)
        _plans[key] = p
    return _plans[key]


# ---------------------------------------------------------------------------
# importance map (computed once on the host, passed to the kernel as data)
# ---------------------------------------------------------------------------
def importance_map(roi: int, mode: str = "constant", sigma_scale: float = 0.125, variant: str = "A"):
    """None for the constant map (the kernel then multiplies by 1.0f exactly like MONAI's ones map);
    for "gaussian" the float32 [roi, roi] map after MONAI's clamp.  Variant "A" (monai 0.9–1.1, the
    range the reference's API usage pins): unit impulse at roi//2 blurred by a separable erf-Gaussian
    and divided by its maximum.  Variant "B" (monai ≥ 1.2): analytic exp(-d²/2σ²) about (roi-1)/2."""
    if mode == "constant":
        return None
    if mode != "gaussian":
        raise ValueError(f"unknown blend mode {mode!r}")
    sigma = sigma_scale * roi
    if variant == "A":
        tail = int(max(float(sigma) * 4.0, 0.5) + 0.5)
        x = torch.arange(-tail, tail + 1, dtype=torch.float32)
        t = 0.70710678 / abs(sigma)
        k = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)
        imp = torch.zeros((1, 1, roi, roi), dtype=torch.float32)
        imp[0, 0, roi // 2, roi // 2] = 1.0
        imp = torch.nn.functional.conv2d(imp, k.view(1, 1, -1, 1), padding=(k.numel() // 2, 0))
        imp = torch.nn.functional.conv2d(imp, k.view(1, 1, 1, -1), padding=(0, k.numel() // 2))
        m = imp[0, 0] / imp.max()
    elif variant == "B":
        x = torch.arange(-(roi - 1) / 2.0, (roi - 1) / 2.0 + 1.0, dtype=torch.float32)
        g = torch.exp(x ** 2 / (-2.0 * sigma ** 2))
        m = g[:, None] * g[None, :]
    else:
        raise ValueError("gaussian variant must be 'A' or 'B'")
    floor = max(m[m != 0].min().item(), 1e-3)
    return torch.clamp(m.to(torch.float32), min=floor).contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise capi.CellSegError(f"{what} must be a CUDA tensor; cellseg_b200 has no CPU path")


# ---------------------------------------------------------------------------
# thin kernel wrappers
# ---------------------------------------------------------------------------
def scale_lut(img_u8: torch.Tensor) -> torch.Tensor:
    """[N,H,W,3] u8 → float32 [N,256] LUT with lut[n][v] = float32(float64(v) / float64(max_n))."""
    _need_cuda(img_u8, "image")
    lib = capi.load()
    N = img_u8.shape[0]
    scratch = torch.empty(N, dtype=torch.int32, device=img_u8.device)
    lut = torch.empty((N, 256), dtype=torch.float32, device=img_u8.device)
    capi.check(lib, lib.cs_scale_lut_u8(img_u8.data_ptr(), N, img_u8[0].numel(), scratch.data_ptr(), lut.data_ptr(), _stream()))
    return lut


def to_unit_float(img_u8: torch.Tensor, lut: torch.Tensor | None = None) -> torch.Tensor:
    """predict.py:107-108 on device: [N,H,W,3] u8 → [N,3,H,W] float32 in [0,1]."""
    _need_cuda(img_u8, "image")
    lib = capi.load()
    img_u8 = img_u8.contiguous()
    N, H, W, ch = img_u8.shape
    assert ch == 3 and img_u8.dtype == torch.uint8
    lut = scale_lut(img_u8) if lut is None else lut
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=img_u8.device)
    capi.check(lib, lib.cs_u8hwc_to_f32chw(img_u8.data_ptr(), N, H, W, lut.data_ptr(), out.data_ptr(), _stream()))
    return out


def gather_windows(src: torch.Tensor, plan: WindowPlan, begin: int, count: int, out: torch.Tensor | None = None,
                   lut: torch.Tensor | None = None) -> torch.Tensor:
    """Windows [begin, begin+count) of the batch (MONAI order: image-major, then y, then x) →
    [count,3,roi,roi] float32.  `src` is float32 [N,3,H,W], or u8 [N,H,W,3] together with `lut`."""
    _need_cuda(src, "image")
    lib = capi.load()
    ys, xs, _, _ = plan.tensors(src.device)
    R = plan.roi
    if out is None:
        out = torch.empty((count, 3, R, R), dtype=torch.float32, device=src.device)
    if src.dtype == torch.uint8:
        N = src.shape[0]
        capi.check(lib, lib.cs_gather_windows_u8(src.data_ptr(), lut.data_ptr(), N, plan.H, plan.W, R, plan.ny, plan.nx, plan.pad_y,
                                                 plan.pad_x, ys.data_ptr(), xs.data_ptr(), begin, count, out.data_ptr(), _stream()))
    else:
        N, ch = src.shape[:2]
        if ch != 3 or src.dtype != torch.float32:
            raise capi.CellSegError("gather_windows expects a float32 [N,3,H,W] image (the reference feeds 3 channels)")
        src = src.contiguous()      # the reference's tensor is a permuted HWC view (predict.py:108)
        capi.check(lib, lib.cs_gather_windows(src.data_ptr(), N, plan.H, plan.W, R, plan.ny, plan.nx, plan.pad_y, plan.pad_x,
                                              ys.data_ptr(), xs.data_ptr(), begin, count, out.data_ptr(), _stream()))
    return out


def blend(win_logits: torch.Tensor, N: int, plan: WindowPlan, imap: torch.Tensor | None = None, threshold: float = 0.5,
          want=("blended",), packed_ptr: int = 0):
    """Fused blend → /count → softmax → p1 > threshold over all windows of N images.
    `want` ⊆ {"blended", "p1", "mask"}; `packed_ptr` (device address inside a post-processing
    workspace) additionally receives the bit-packed mask.  Returns a dict of the requested tensors."""
    _need_cuda(win_logits, "window logits")
    lib = capi.load()
    dev = win_logits.device
    Cc = win_logits.shape[1]
    if win_logits.shape[0] != N * plan.n_win or win_logits.dtype != torch.float32 or not win_logits.is_contiguous():
        raise capi.CellSegError("window logits must be contiguous float32 [N*n_win, C, roi, roi]")
    ys, xs, cy, cx = plan.tensors(dev)
    out = {}
    if "blended" in want: out["blended"] = torch.empty((N, Cc, plan.H, plan.W), dtype=torch.float32, device=dev)
    if "p1" in want: out["p1"] = torch.empty((N, plan.H, plan.W), dtype=torch.float32, device=dev)
    if "mask" in want: out["mask"] = torch.empty((N, plan.H, plan.W), dtype=torch.uint8, device=dev)
    ptr = lambda k: out[k].data_ptr() if k in out else None
    if imap is not None:
        imap = imap.to(dev, torch.float32).contiguous()
    capi.check(lib, lib.cs_blend_softmax_threshold(
        win_logits.data_ptr(), Cc, N, plan.H, plan.W, plan.roi, plan.ny, plan.nx, plan.pad_y, plan.pad_x,
        ys.data_ptr(), xs.data_ptr(), cy.data_ptr(), cx.data_ptr(), imap.data_ptr() if imap is not None else None,
        float(threshold), packed_ptr or None, ptr("mask"), ptr("p1"), ptr("blended"), _stream()))
    return out


# ---------------------------------------------------------------------------
# MONAI-signature drop-in
# ---------------------------------------------------------------------------
def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, sw_device=None, device=None,
                             gaussian_variant="A", *args, **kwargs):
    """Drop-in for `monai.inferers.sliding_window_inference(inputs, roi_size, sw_batch_size, predictor)`
    (baseline/predict.py:109).  inputs: CUDA float32 [B,3,H,W]; returns the blended predictor output
    [B,C,H,W].  Windows are produced, predicted `sw_batch_size` at a time and blended in MONAI's order;
    the arithmetic is identical (`out += w*pred; cnt += w; out/cnt`), the gather and blend are single kernels."""
    roi = int(roi_size if isinstance(roi_size, int) else roi_size[0])
    if not isinstance(roi_size, int) and len(set(int(r) for r in roi_size)) != 1:
        raise NotImplementedError("square ROI only (the reference uses 256x256)")
    if padding_mode != "constant" or cval != 0:
        raise NotImplementedError("only MONAI's default zero padding is on the reference path")
    _need_cuda(inputs, "inputs")
    x = inputs.to(torch.float32).contiguous()
    B, _, H, W = x.shape
    plan = make_plan(H, W, roi, float(overlap))
    total = B * plan.n_win
    logits = None
    for g in range(0, total, sw_batch_size):
        k = min(sw_batch_size, total - g)
        pred = predictor(gather_windows(x, plan, g, k), *args, **kwargs)
        if logits is None:
            logits = torch.empty((total,) + tuple(pred.shape[1:]), dtype=torch.float32, device=x.device)
        logits[g:g + k] = pred
    imap = importance_map(roi, mode, sigma_scale, gaussian_variant)
    return blend(logits, B, plan, imap, want=("blended",))["blended"]


# ---------------------------------------------------------------------------
# fused image → instance map engine
# ---------------------------------------------------------------------------
class SegmentationEngine:
    """u8 images [N,H,W,3] → int32 instance maps [N,H,W]  (predict.py:107-113 end to end on device).

    K0 (max + scale table) → predictor over `chunk` windows per forward → K2 (blend + softmax + threshold → packed mask) → hole
    fill → small-object removal → raster-order labelling.  When the predictor is `cellseg_b200.unet.UNet` (the reference's
    network) its channels-last runner reads the windows straight from the u8 image inside the first convolutions and writes
    the logits in place (no gather kernel, no window batch, no copies); any other predictor gets K1 (gather from the u8 image)
    and a copy of its output.  `cuda_graph=True` replays the whole call as one CUDA graph.  Buffers are cached per geometry;
    everything runs on torch's current stream.
    """

    def __init__(self, predictor, roi: int = 256, overlap: float = 0.25, mode: str = "constant", chunk: int = 64,
                 hole_area: int = 64, min_size: int = 16, threshold: float = 0.5, sigma_scale: float = 0.125,
                 gaussian_variant: str = "A", channels_last: bool = False, cuda_graph: bool = False):
        self.predictor = predictor
        self.roi, self.overlap, self.mode, self.chunk = roi, overlap, mode, chunk
        self.hole_area, self.min_size, self.threshold = hole_area, min_size, threshold
        self.imap = importance_map(roi, mode, sigma_scale, gaussian_variant)
        self.channels_last = channels_last
        self.cuda_graph = cuda_graph
        self._bufs = {}
        self._graphs = {}
        self._call_graphs = {}
        self._graph_pool = None
        self.graph_kernel_launches = 0     # libcellseg kernels executed by CUDA-graph replays (not seen by the launch counter)

    def _buffers(self, N, H, W, device, slot: int = 0):
        """Per-geometry buffers.  `slot` selects an independent set of OUTPUT buffers (labels, counts) so that one
        sub-batch can still be copied to the host while the next one is computed; logits and scratch are shared
        (compute is serialised on one stream)."""
        key = (N, H, W, str(device))
        if key not in self._bufs:
            plan = make_plan(H, W, self.roi, self.overlap)
            lib = capi.load()
            wsb = lib.cs_postprocess_workspace_bytes(N, H, W)
            while len(self._bufs) >= 8:
                self._bufs.pop(next(iter(self._bufs)))   # oldest geometry; captured graphs and segment_host's in-flight list
                                                         # keep what they still use alive (see __call__, segment_host)
            # every geometry owns its buffers, the post-processing workspace included: a captured CUDA graph has their addresses
            # baked in, so nothing a graph touches may ever be reallocated or shared with a geometry of another size
            self._bufs[key] = dict(
                plan=plan, logits=None, out={},
                lut=torch.empty((N, 256), dtype=torch.float32, device=device),
                lut_scratch=torch.empty(N, dtype=torch.int32, device=device),
                wsb=wsb, ws=torch.empty(max(wsb, 1), dtype=torch.uint8, device=device))
        b = self._bufs[key]
        if slot not in b["out"]:
            b["out"][slot] = (torch.empty((N, H, W), dtype=torch.int32, device=device), torch.empty((N,), dtype=torch.int32, device=device))
        b["labels"], b["counts"] = b["out"][slot]
        return b

    def _forward(self, k: int, device, graph: bool | None = None):
        """(static input [k,3,R,R], run() → output [k,C,R,R]) for a forward over k windows."""
        graph = self.cuda_graph if graph is None else graph
        key = (k, str(device), graph, self._weights_signature() if graph else None)
        if key not in self._graphs:
            win = torch.zeros((k, 3, self.roi, self.roi), dtype=torch.float32, device=device)
            if self.channels_last:
                win = win.contiguous(memory_format=torch.channels_last)
            if graph:
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream(device=device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    for _ in range(3):
                        self.predictor(win)          # cuDNN autotuning and lazy initialisation happen outside the capture
                cur.wait_stream(side)
                lib = capi.load()
                l0 = lib.cs_kernel_launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self.predictor(win)
                per_replay = lib.cs_kernel_launch_count() - l0     # libcellseg kernels inside the graph (fused norm blocks)

                def replay():
                    graph.replay()
                    self.graph_kernel_launches += per_replay
                    return out
                self._graphs[key] = (win, replay)
            else:
                self._graphs[key] = (win, lambda: self.predictor(win))
        return self._graphs[key]

    def _weights_signature(self):
        params = getattr(self.predictor, "parameters", None)
        if params is None:
            return None
        return hash(tuple((p._version, p.data_ptr()) for p in params()))

    def _fused_runner(self, device):
        """The predictor's channels-last runner when it can read the windows straight from the uint8 images."""
        from .unet import UNet
        p = self.predictor
        if isinstance(p, UNet) and self.roi % 2 == 0 and not self.channels_last and next(p.parameters()).device == device:
            return p.runner()
        return None

    @torch.no_grad()
    def window_logits(self, img_u8: torch.Tensor, slot: int = 0):
        """Stage 1: all window logits of the batch, [N*n_win, C, roi, roi] (kept for parity tests)."""
        _need_cuda(img_u8, "image")
        img_u8 = img_u8.contiguous()
        N, H, W, _ = img_u8.shape
        lib = capi.load()
        b = self._buffers(N, H, W, img_u8.device, slot)
        plan = b["plan"]
        capi.check(lib, lib.cs_scale_lut_u8(img_u8.data_ptr(), N, img_u8[0].numel(), b["lut_scratch"].data_ptr(),
                                            b["lut"].data_ptr(), _stream()))
        total = N * plan.n_win
        chunk = min(self.chunk, total)
        runner = self._fused_runner(img_u8.device)
        if runner is not None:
            # gather + scaling fused into the first convolutions, logits written in place: no window batch, no copies
            if b["logits"] is None or b["logits"].shape[1] != 3:
                b["logits"] = torch.empty((total, 3, self.roi, self.roi), dtype=torch.float32, device=img_u8.device)
            for g in range(0, total, chunk):
                k = min(chunk, total - g)
                runner.from_images(img_u8, b["lut"], plan, g, k, out=b["logits"][g:g + k])
            return b["logits"], b
        per_chunk_graph = self.cuda_graph and not torch.cuda.is_current_stream_capturing()
        for g in range(0, total, chunk):
            k = min(chunk, total - g)
            win, run = self._forward(chunk if per_chunk_graph else k, img_u8.device, per_chunk_graph)
            if self.channels_last:
                win[:k].copy_(gather_windows(img_u8, plan, g, k, lut=b["lut"]))   # gather NCHW, restride to NHWC
            else:
                gather_windows(img_u8, plan, g, k, out=win[:k], lut=b["lut"])
            pred = run()
            if b["logits"] is None or b["logits"].shape[1] != pred.shape[1]:
                b["logits"] = torch.empty((total, pred.shape[1], self.roi, self.roi), dtype=torch.float32, device=img_u8.device)
            b["logits"][g:g + k] = pred[:k]
        return b["logits"], b

    @torch.no_grad()
    def _call_eager(self, img_u8: torch.Tensor, slot: int = 0):
        logits, b = self.window_logits(img_u8, slot)
        N, H, W, _ = img_u8.shape
        lib = capi.load()
        ws = b["ws"]
        slot = lib.cs_postprocess_packed_slot(ws.data_ptr(), ws.numel(), N, H, W)
        if not slot:
            raise capi.CellSegError("post-processing workspace too small")
        if self.imap is not None and self.imap.device != img_u8.device:
            self.imap = self.imap.to(img_u8.device, torch.float32).contiguous()
        blend(logits, N, b["plan"], self.imap, self.threshold, want=(), packed_ptr=slot)
        capi.check(lib, lib.cs_postprocess_packed(N, H, W, self.hole_area, self.min_size, 8, b["labels"].data_ptr(),
                                                  b["counts"].data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        return b["labels"], b["counts"]

    @torch.no_grad()
    def __call__(self, img_u8: torch.Tensor, slot: int = 0):
        """→ (labels int32 [N,H,W], counts int32 [N]); the tensors are reused by the next call on the same slot.

        With `cuda_graph=True` the whole call — scale table, every predictor forward, blend, post-processing — is captured
        once per (input buffer, geometry, slot) and replayed as ONE CUDA graph: the step is ~900 short kernels, and replay
        removes their launch gaps and all Python/ctypes overhead.  The input must then live in a buffer that is reused
        (a staging buffer, as `segment_host` does); a new buffer address triggers a new capture."""
        if not self.cuda_graph:
            return self._call_eager(img_u8, slot)
        _need_cuda(img_u8, "image")
        if not img_u8.is_contiguous():
            raise capi.CellSegError("graph mode needs a contiguous uint8 image batch")
        # a captured graph has the predictor's weights baked in (kernel parameters, derived channels-last copies): the weights'
        # (version, address) signature is part of the key, so load_state_dict / an in-place update leads to a fresh capture
        key = (img_u8.data_ptr(), tuple(img_u8.shape), slot, str(img_u8.device), self._weights_signature())
        ent = self._call_graphs.get(key)
        if ent is None:
            lib = capi.load()
            cur = torch.cuda.current_stream(img_u8.device)
            side = torch.cuda.Stream(device=img_u8.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):          # cuDNN autotuning, lazy initialisation, buffer and workspace allocation
                    self._call_eager(img_u8, slot)
            cur.wait_stream(side)
            torch.cuda.synchronize(img_u8.device)
            l0 = lib.cs_kernel_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._graph_pool):
                out = self._call_eager(img_u8, slot)
            if self._graph_pool is None:
                self._graph_pool = graph.pool()
            while len(self._call_graphs) >= 16:
                self._call_graphs.pop(next(iter(self._call_graphs)))      # oldest capture
            # everything whose address the graph has baked in stays alive with it: the input tensor, the geometry's buffer set
            # (logits, scale table, workspace, label slots) and the runner's derived weights as of the capture
            runner = self._fused_runner(img_u8.device)
            keep = (img_u8, self._buffers(*img_u8.shape[:3], img_u8.device, slot), list(runner._wcache.values()) if runner is not None else None)
            ent = (graph, out, lib.cs_kernel_launch_count() - l0, keep)
            self._call_graphs[key] = ent
        ent[0].replay()
        self.graph_kernel_launches += ent[2]
        return ent[1]

    @staticmethod
    def _schedule(N: int, sub: int, ramp: bool) -> list[int]:
        """Sub-batch sizes for the host pipeline.  With `ramp` the first and last sub-batches are a quarter of `sub`: the first
        host→device copy and the last device→host copy are the only ones nothing overlaps, so they should be short."""
        q = sub // 4
        if not ramp or q < 1 or N < 2 * sub:
            return [min(sub, N - s) for s in range(0, N, sub)]
        sizes = [q, sub - q]
        left = N - 2 * sub
        sizes += [sub] * (left // sub)
        if left % sub:
            sizes.append(left % sub)
        return sizes + [sub - q, q]

    @torch.no_grad()
    def segment_host(self, img_u8_host: torch.Tensor, labels_host: torch.Tensor | None = None, device=None, sub_batch: int = 16, ramp: bool = True,
                     wait: bool = True):
        """Host buffers in → host buffers out, the call a user of the reference's predict.py:107-113 makes:
        u8 [N,H,W,3] (pinned for asynchronous copies) → int32 [N,H,W] instance maps in `labels_host` (allocated
        pinned when None).  The batch is cut into sub-batches of up to `sub_batch` images (shorter at both ends, see
        `_schedule`) that flow through three streams — host→device copy, compute (torch's current stream), device→host
        copy — so that both PCIe directions overlap the computation.

        `wait=True` returns after everything has landed in `labels_host`.  `wait=False` returns as soon as the work is queued: the
        next call's upload then overlaps this call's computation and download (a stream of batches keeps all three streams busy
        across call boundaries); `labels_host` and the returned counts may be read after `engine.wait()`.  Buffer reuse across calls
        is ordered by events per staging slot and per label slot, not by draining the streams."""
        device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        N, H, W, _ = img_u8_host.shape
        sub = max(1, min(int(sub_batch), N))
        if labels_host is None:
            labels_host = torch.empty((N, H, W), dtype=torch.int32, pin_memory=True)
        # staging buffers are cached per geometry (not replaced when another size comes by): in graph mode the captured calls
        # are keyed by the staging buffer's address, so a fresh buffer would mean a fresh capture
        key = (sub, H, W, str(device))
        if not hasattr(self, "_stages"):
            self._stages, self._copy_streams = {}, {}
            self._stage_busy, self._label_busy, self._pending, self._next_slot = {}, {}, [], {}
        if key not in self._stages:
            self._stages[key] = [torch.empty((sub, H, W, 3), dtype=torch.uint8, device=device) for _ in range(2)]
        if str(device) not in self._copy_streams:
            self._copy_streams[str(device)] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
        self._stage = self._stages[key]
        self._cin, self._cout = self._copy_streams[str(device)]
        main = torch.cuda.current_stream(device)
        counts_all = torch.empty((N,), dtype=torch.int32, device=device)
        s0 = 0
        for n in self._schedule(N, sub, ramp):
            # the two staging / label slots alternate across sub-batches AND across calls (a call with a single sub-batch must not
            # land in the slot the previous call still computes on)
            slot = self._next_slot.get(key, 0)
            self._next_slot[key] = slot ^ 1
            with torch.cuda.stream(self._cin):
                busy = self._stage_busy.get((key, slot))
                if busy is not None:
                    self._cin.wait_event(busy)                       # the computation that read this staging slot last (this call or an earlier one)
                self._stage[slot][:n].copy_(img_u8_host[s0:s0 + n], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._cin)
            main.wait_event(ev_in)
            lkey = (n, H, W, str(device), slot)
            busy = self._label_busy.get(lkey)
            if busy is not None:
                main.wait_event(busy)                                # the download that read this label slot last
            labels, counts = self(self._stage[slot][:n], slot)
            counts_all[s0:s0 + n] = counts
            ev_done = torch.cuda.Event()
            ev_done.record(main)
            self._stage_busy[(key, slot)] = ev_done
            # Label tensors whose device→host copy may still be queued on the copy-out stream.  The geometry / graph caches may evict
            # their buffer set while the pipeline runs (a third distinct sub-batch size in a row does it); holding the tensors until
            # wait() — and telling the allocator about the second stream — keeps the blocks from being handed to the next allocation
            # under a pending copy.
            self._pending.append((labels, counts, counts_all))
            labels.record_stream(self._cout)
            with torch.cuda.stream(self._cout):
                self._cout.wait_event(ev_done)
                labels_host[s0:s0 + n].copy_(labels, non_blocking=True)
                ev_out = torch.cuda.Event()
                ev_out.record(self._cout)
            self._label_busy[lkey] = ev_out
            s0 += n
        if wait:
            self.wait(device)
        return labels_host, counts_all

    def wait(self, device=None):
        """Block until every `segment_host` call queued so far has delivered its labels to host memory."""
        if not hasattr(self, "_copy_streams"):
            return
        for name, (cin, cout) in self._copy_streams.items():
            if device is not None and name != str(torch.device(device)):
                continue
            main = torch.cuda.current_stream(torch.device(name))
            main.wait_stream(cout)
            main.synchronize()
        self._pending.clear()
