"""Strip-sharded whole-slide path (BASELINE config 3): one image, row strips over
the GPUs of a box, one process per GPU.  The result is bit-identical to the single-GPU path.

Sliding window: the window grid is global (monai's grid for the whole image, predict.py:109).  A rank runs every
window row that intersects its strip and blends only its own rows, visiting the covering windows in ascending global
index — no halo exchange; the only collective is an all-reduce(MAX) for the `/ max` scaling (predict.py:107).

Post-processing (predict.py:113): each strip runs the tile chain of libcellseg on its own rows; between phases the
ranks all-gather their strips' first / last rows (≤ 4·W ints per rank and phase):
  hole / object filters   pieces smaller than the threshold that touch the cut are merged with the facing pieces of
                          the neighbouring strip inside the strip's node union-find (sizes add up, contact with a
                          large piece makes the whole piece large)
  labelling               every rank resolves the same small union-find over strip-local label ranks (raster order
                          is preserved by offsets[strip] + rank) and gets a LUT rank → global id, applied in k_final.
The library itself never communicates; this module moves the rows with torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  `run_lockstep` drives several strips in one process — the single-GPU and CPU test harness.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi

TILE_ROWS = 64


def strip_rows(H: int, world: int) -> list[tuple[int, int]]:
    """Row ranges of the strips: whole 64-row tile rows, as even as possible; empty strips are not produced, so the
    list may be shorter than `world` for very small images."""
    nty = -(-H // TILE_ROWS)
    n = min(world, nty)
    cuts = [(nty * i) // n for i in range(n + 1)]
    return [(cuts[i] * TILE_ROWS, min(H, cuts[i + 1] * TILE_ROWS)) for i in range(n)]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_of(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream) if t.is_cuda else None


class StripPost:
    """The post-processing phases of one strip (thin wrapper over the cs_strip_* entry points)."""

    def __init__(self, Hs: int, W: int, device, hole_area: int = 64, min_size: int = 16, connectivity: int = 2):
        if connectivity not in (1, 2):
            raise ValueError("connectivity must be 1 or 2")
        self.lib = capi.load()
        self.Hs, self.W, self.device = int(Hs), int(W), torch.device(device)
        self.hole, self.min, self.conn = int(hole_area), int(min_size), 4 if connectivity == 1 else 8
        self.wsb = self.lib.cs_strip_workspace_bytes(self.Hs, self.W)
        if self.wsb == 0:
            capi.check(self.lib, -1)
        self.ws = torch.empty(self.wsb, dtype=torch.uint8, device=self.device)
        self.labels = torch.empty((self.Hs, self.W), dtype=torch.int32, device=self.device)
        self.count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lut = None
        self.total = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _args(self):
        return self.Hs, self.W, self.hole, self.min, self.conn, _ptr(self.labels)

    def _tail(self):
        return _ptr(self.ws), self.wsb, _stream_of(self.ws)

    def packed_slot(self) -> int:
        slot = self.lib.cs_strip_packed_slot(_ptr(self.ws), self.wsb, self.Hs, self.W)
        if not slot:
            raise capi.CellSegError("strip workspace too small")
        return slot

    def begin(self, mask: torch.Tensor | None = None):
        """mask: u8/bool [Hs, W] on the strip's device, or None when the packed mask was written into packed_slot()."""
        if mask is not None:
            mask = (mask != 0).view(torch.uint8) if mask.dtype not in (torch.uint8, torch.bool) else mask
            mask = mask.contiguous().view(torch.uint8)
            assert tuple(mask.shape) == (self.Hs, self.W)
        capi.check(self.lib, self.lib.cs_strip_begin(_ptr(mask), *self._args(), *self._tail()))
        self._keep = mask

    def filter_export(self, stage: int) -> torch.Tensor:
        rows = torch.empty((2, 2, self.W), dtype=torch.int32, device=self.device)
        capi.check(self.lib, self.lib.cs_strip_filter_export(stage, *self._args(), _ptr(rows), *self._tail()))
        return rows

    def filter_resume(self, stage: int, above_last: torch.Tensor | None, below_first: torch.Tensor | None):
        a = None if above_last is None else above_last.contiguous()
        b = None if below_first is None else below_first.contiguous()
        capi.check(self.lib, self.lib.cs_strip_filter_resume(stage, *self._args(), _ptr(a), _ptr(b), *self._tail()))

    def label_export(self) -> torch.Tensor:
        """→ int32 [2*W + 1]: first-row ranks, last-row ranks, component count of the strip."""
        out = torch.empty(2 * self.W + 1, dtype=torch.int32, device=self.device)
        capi.check(self.lib, self.lib.cs_strip_label_export(self.Hs, self.W, self.conn, _ptr(self.labels), _ptr(out),
                                                            C.c_void_p(out.data_ptr() + 8 * self.W), *self._tail()))
        return out

    def label_resolve(self, gathered: torch.Tensor, my_index: int):
        """gathered: int32 [n_strips, 2*W + 1] (every strip's label_export).  Reads the counts on the host (one
        synchronisation) to size the global union-find, then builds this strip's rank → global-id LUT."""
        n = gathered.shape[0]
        counts = [int(c) for c in gathered[:, -1].tolist()]
        offsets = np.zeros(n + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(counts)
        total = int(offsets[-1])
        rows_all = gathered[:, : 2 * self.W].contiguous()
        off_dev = torch.from_numpy(offsets).to(self.device)
        ws2b = self.lib.cs_strip_resolve_workspace_bytes(total)
        ws2 = torch.empty(max(ws2b, 1), dtype=torch.uint8, device=self.device)
        self.lut = torch.empty(counts[my_index] + 1, dtype=torch.int32, device=self.device)
        capi.check(self.lib, self.lib.cs_strip_label_resolve(_ptr(rows_all), _ptr(off_dev), n, self.W, self.conn, my_index,
                                                             int(offsets[my_index]), counts[my_index], total, _ptr(self.lut), _ptr(self.total),
                                                             _ptr(ws2), ws2b, _stream_of(self.ws)))
        self._keep2 = (rows_all, off_dev, ws2)

    def final(self) -> torch.Tensor:
        capi.check(self.lib, self.lib.cs_strip_final(self.Hs, self.W, _ptr(self.lut), _ptr(self.labels), *self._tail()))
        return self.labels


# ---------------------------------------------------------------------------
# drivers
# ---------------------------------------------------------------------------
def run_lockstep(strips: list[StripPost], masks: list | None):
    """All strips of one image in one process, phase by phase (what the ranks do concurrently).  → (labels list, total)."""
    n = len(strips)
    for i, s in enumerate(strips):
        s.begin(None if masks is None else masks[i])
    for stage in (1, 2):
        rows = [s.filter_export(stage) for s in strips]
        for i, s in enumerate(strips):
            s.filter_resume(stage, rows[i - 1][1] if i > 0 else None, rows[i + 1][0] if i < n - 1 else None)
    gathered = torch.stack([s.label_export() for s in strips])
    for i, s in enumerate(strips):
        s.label_resolve(gathered, i)
    labels = [s.final() for s in strips]
    return labels, int(strips[0].total.item())


def run_distributed(strip: StripPost, mask, rank: int, world: int, group=None):
    """This rank's strip of a `world`-strip image; collectives through torch.distributed.  → (labels, total)."""
    import torch.distributed as dist

    def gather(t):
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous().view(-1), group=group)
        return out.view((world,) + tuple(t.shape))

    strip.begin(mask)
    for stage in (1, 2):
        rows = gather(strip.filter_export(stage))
        strip.filter_resume(stage, rows[rank - 1, 1] if rank > 0 else None, rows[rank + 1, 0] if rank < world - 1 else None)
    gathered = gather(strip.label_export())
    strip.label_resolve(gathered, rank)
    labels = strip.final()
    return labels, int(strip.total.item())


# ---------------------------------------------------------------------------
# image strip → instance-map strip (sliding window + UNet + post-processing)
# ---------------------------------------------------------------------------
def window_plan(H: int, W: int, world: int, roi: int = 256, overlap: float = 0.25):
    """Who predicts and who needs which windows of the global grid (flat, row-major window indices).  Every window is predicted
    exactly once: rank r owns the contiguous range [r·n/world, (r+1)·n/world) of the n = ny·nx windows — equal shares to within
    one window, whatever the number of window rows — and *needs* all windows of the window rows that intersect its pixel strip.
    → (strips, owned[(g0, g1)] per rank, needed[(g0, g1)] per rank, transfers[(owner, needer, g0, g1)] of contiguous ranges)."""
    from .sliding_window import make_plan
    plan = make_plan(H, W, roi, overlap)
    strips = strip_rows(H, world)
    if len(strips) != world:
        raise ValueError(f"image of {H} rows is too small for {world} strips")
    n = plan.ny * plan.nx
    owned = [(r * n // world, (r + 1) * n // world) for r in range(world)]
    needed, transfers = [], []
    for q, (r0, r1) in enumerate(strips):
        lo = min(i for i, y in enumerate(plan.ys) if y + roi > r0)
        hi = max(i for i, y in enumerate(plan.ys) if y < r1) + 1
        needed.append((lo * plan.nx, hi * plan.nx))
        for o, (g0, g1) in enumerate(owned):
            a, b = max(g0, lo * plan.nx), min(g1, hi * plan.nx)
            if o != q and a < b:
                transfers.append((o, q, a, b))
    return strips, owned, needed, transfers


class StripSegmenter:
    """One rank's share of predict.py:107-113 on a whole-slide image of H x W pixels split into `world` row strips.

    Every window of the global grid is predicted by exactly one rank (equal contiguous shares of the flat window list, to
    within one window); the windows a strip needs but does not own — one or two window rows next to each cut — arrive from
    the neighbouring ranks as logits over NVLink (NCCL send/recv, ≈ 40 MB per window row at W = 10 000), far cheaper than
    predicting them twice."""

    def __init__(self, predictor, H: int, W: int, rank: int, world: int, roi: int = 256, overlap: float = 0.25, mode: str = "constant",
                 chunk: int = 100, hole_area: int = 64, min_size: int = 16, threshold: float = 0.5, cuda_graph: bool = True,
                 device=None):
        from .sliding_window import SegmentationEngine, importance_map, make_plan
        strips, owned, needed, transfers = window_plan(H, W, world, roi, overlap)
        self.H, self.W, self.rank, self.world = H, W, rank, world
        self.r0, self.r1 = strips[rank]
        self.plan = make_plan(H, W, roi, overlap)
        if self.plan.pad_y or self.plan.pad_x:
            raise ValueError("strip sharding is for images larger than the ROI")
        ys, nx = self.plan.ys, self.plan.nx
        self.own_lo, self.own_hi = owned[rank]                       # flat window indices this rank predicts
        self.need_lo, self.need_hi = needed[rank]                    # flat window indices its strip blends
        self.buf_lo = min(self.own_lo, self.need_lo) if self.own_hi > self.own_lo else self.need_lo
        self.buf_hi = max(self.own_hi, self.need_hi)
        self.transfers = transfers
        self.sends = [(q, a, b) for o, q, a, b in transfers if o == rank]
        self.recvs = [(o, a, b) for o, q, a, b in transfers if q == rank]
        if self.own_hi > self.own_lo:
            self.src0, self.src1 = ys[self.own_lo // nx], ys[(self.own_hi - 1) // nx] + roi     # image rows this rank's own windows read
        else:
            self.src0, self.src1 = self.r0, self.r0 + 1
        self.n_win = self.own_hi - self.own_lo                                              # windows this rank predicts
        self.roi, self.threshold = roi, threshold
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.engine = SegmentationEngine(predictor, roi=roi, overlap=overlap, mode=mode, chunk=chunk, cuda_graph=cuda_graph)
        self.imap = importance_map(roi, mode)
        if self.imap is not None:
            self.imap = self.imap.to(self.device)
        self.post = StripPost(self.r1 - self.r0, W, self.device, hole_area, min_size, 2)
        self.lib = capi.load()
        self.maxv = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lut = torch.empty((1, 256), dtype=torch.float32, device=self.device)
        self.logits = None
        self.phase_ms = None               # set to {} to collect per-phase device times in __call__

    def source_rows(self) -> tuple[int, int]:
        """Image rows [src0, src1) this rank's windows read."""
        return self.src0, self.src1

    def windows(self, g0: int, g1: int) -> torch.Tensor:
        """Logits of the windows [g0, g1) of the global grid, [g1-g0, C, R, R] (a view into this rank's buffer)."""
        return self.logits[g0 - self.buf_lo:g1 - self.buf_lo]

    @torch.no_grad()
    def local_max(self, src_rows_u8: torch.Tensor) -> torch.Tensor:
        """Max over this rank's rows (device int32[1]); the global `/ max` of predict.py:107 is the max over ranks."""
        img = src_rows_u8.contiguous()
        capi.check(self.lib, self.lib.cs_u8_max(_ptr(img), 1, img.numel(), _ptr(self.maxv), _stream_of(img)))
        return self.maxv

    @torch.no_grad()
    def predict_own_windows(self, src_rows_u8: torch.Tensor, global_max: torch.Tensor, before_forward=None):
        """Gather + UNet for the window rows this rank owns, into its logits buffer.  `before_forward(g, k)` is called ahead of the
        forward over this rank's windows [g, g + k) (segment_host waits there for the image rows that forward reads)."""
        lib, plan, R = self.lib, self.plan, self.roi
        img = src_rows_u8.contiguous()
        assert tuple(img.shape) == (self.src1 - self.src0, self.W, 3) and img.dtype == torch.uint8
        st = _stream_of(img)
        ys_d, xs_d, _, _ = plan.tensors(self.device)
        capi.check(lib, lib.cs_scale_lut_from_max(_ptr(global_max), 1, _ptr(self.lut), st))
        begin = self.own_lo
        chunk = max(1, min(self.engine.chunk, self.n_win))
        runner = self.engine._fused_runner(self.device)
        if runner is not None:
            # the first convolutions read the strip's rows directly (gather and /max scaling fused), logits land in place
            if self.logits is None:
                self.logits = torch.empty((self.buf_hi - self.buf_lo, 3, R, R), dtype=torch.float32, device=self.device)
            for g in range(0, self.n_win, chunk):
                k = min(chunk, self.n_win - g)
                o = self.own_lo - self.buf_lo + g
                if before_forward is not None:
                    before_forward(g, k)
                runner.from_window_grid(img, self.lut, 1, img.shape[0], self.W, R, plan.ny, plan.nx, self.src0, 0, ys_d, xs_d, begin + g, k,
                                        out=self.logits[o:o + k])
            return
        for g in range(0, self.n_win, chunk):
            k = min(chunk, self.n_win - g)
            if before_forward is not None:
                before_forward(g, k)
            win, run = self.engine._forward(chunk if self.engine.cuda_graph else k, self.device)
            capi.check(lib, lib.cs_gather_windows_u8(_ptr(img), _ptr(self.lut), 1, img.shape[0], self.W, R, plan.ny, plan.nx, self.src0, 0,
                                                     _ptr(ys_d), _ptr(xs_d), begin + g, k, _ptr(win), st))
            pred = run()
            if self.logits is None:
                self.logits = torch.empty((self.buf_hi - self.buf_lo, pred.shape[1], R, R), dtype=torch.float32, device=self.device)
            o = self.own_lo - self.buf_lo + g
            self.logits[o:o + k] = pred[:k]

    def exchange_window_rows(self):
        """Send the owned window rows the neighbours need, receive the needed rows owned elsewhere (NCCL P2P)."""
        import torch.distributed as dist
        if not self.sends and not self.recvs:
            return
        ops = [dist.P2POp(dist.isend, self.windows(a, b), q) for q, a, b in self.sends]
        ops += [dist.P2POp(dist.irecv, self.windows(a, b), o) for o, a, b in self.recvs]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    @torch.no_grad()
    def blend_strip(self):
        """Blend this strip's rows from the logits of the window rows it needs → packed mask in the post-processing slot."""
        lib, plan, R = self.lib, self.plan, self.roi
        ys_d, xs_d, cy_d, cx_d = plan.tensors(self.device)
        Cc = self.logits.shape[1]
        # window g of the global grid lives at logits[g - buf_lo]
        virt = self.logits.data_ptr() - self.buf_lo * Cc * R * R * 4
        capi.check(lib, lib.cs_blend_softmax_threshold(C.c_void_p(virt), Cc, 1, self.r1 - self.r0, self.W, R, plan.ny, plan.nx, self.r0, 0,
                                                       _ptr(ys_d), _ptr(xs_d), _ptr(cy_d), _ptr(cx_d), _ptr(self.imap), float(self.threshold),
                                                       C.c_void_p(self.post.packed_slot()), None, None, None, _stream_of(self.logits)))

    @torch.no_grad()
    def __call__(self, src_rows_u8: torch.Tensor):
        """src_rows_u8: u8 [src1-src0, W, 3] on the device — image rows [src0, src1).  → (labels [r1-r0, W], total).
        With `self.phase_ms` set to a dict, per-phase device times (CUDA events on the current stream) are accumulated into it."""
        import torch.distributed as dist
        marks = []

        def mark(name):
            if self.phase_ms is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        mx = self.local_max(src_rows_u8)
        if self.world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mark("max + all-reduce")
        self.predict_own_windows(src_rows_u8, mx)
        mark("predictor")
        if self.world > 1:
            self.exchange_window_rows()
        mark("logit exchange")
        self.blend_strip()
        mark("blend")
        if self.world > 1:
            out = run_distributed(self.post, None, self.rank, self.world)
        else:
            labels, total = run_lockstep([self.post], None)
            out = (labels[0], total)
        mark("post-processing + label merge")
        if self.phase_ms is not None:
            torch.cuda.synchronize()
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                self.phase_ms[name] = self.phase_ms.get(name, 0.0) + a.elapsed_time(b)
        return out


    def _after_predictor(self):
        """Logit exchange → blend → post-processing with the label merge (the part of __call__ behind the predictor)."""
        if self.world > 1:
            self.exchange_window_rows()
        self.blend_strip()
        if self.world > 1:
            return run_distributed(self.post, None, self.rank, self.world)
        labels, total = run_lockstep([self.post], None)
        return labels[0], total

    @torch.no_grad()
    def segment_host(self, host_rows_u8: torch.Tensor, out_host: torch.Tensor | None = None, speculative_max: int = 255, copy_chunks: int = 8):
        """Host rows in → host labels out, with the upload hidden behind the predictor.

        `host_rows_u8`: u8 [src1-src0, W, 3] in (pinned) host memory — image rows [src0, src1); → (int32 [r1-r0, W] labels in
        `out_host`, total).  predict.py:107 scales by the maximum of the WHOLE image, so nothing can be scaled before the last row has
        been seen; but what reaches that line is the output of the percentile normalisation to uint8 (predict.py:94-104), whose maximum
        is 255 unless a channel is degenerate.  So the rows are uploaded in `copy_chunks` pieces on a copy stream, the predictor starts
        on the first piece with the table for `speculative_max`, every forward waits only for the rows it reads, and the true maximum
        (k_u8_max over the uploaded rows + the all-reduce over the strips) is checked at the end: if the guess was wrong the step is
        redone the plain way, so the result never depends on the guess."""
        import torch.distributed as dist
        dev = self.device
        n_rows = self.src1 - self.src0
        assert tuple(host_rows_u8.shape) == (n_rows, self.W, 3) and host_rows_u8.dtype == torch.uint8
        if getattr(self, "_stage", None) is None:
            self._stage = torch.empty((n_rows, self.W, 3), dtype=torch.uint8, device=dev)
            self._cin = torch.cuda.Stream(device=dev)
            self._guess = torch.zeros(1, dtype=torch.int32, device=dev)
        if out_host is None:
            out_host = torch.empty((self.r1 - self.r0, self.W), dtype=torch.int32).pin_memory()
        stage, cin = self._stage, self._cin
        main = torch.cuda.current_stream(dev)
        self._guess.fill_(int(speculative_max))
        cin.wait_stream(main)                                   # whoever read the staging rows last has finished
        cuts = [n_rows * i // copy_chunks for i in range(copy_chunks + 1)]
        events = []
        with torch.cuda.stream(cin):
            for a, b in zip(cuts[:-1], cuts[1:]):
                if b > a:
                    stage[a:b].copy_(host_rows_u8[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cin)
                    events.append((b, ev))
        waited = 0

        def before_forward(g, k):
            nonlocal waited
            last = self.own_lo + g + k - 1                      # flat index of the forward's last window
            need = min(self.plan.ys[last // self.plan.nx] + self.roi - self.src0, n_rows)
            while waited < len(events) and (waited == 0 or events[waited - 1][0] < need):
                main.wait_event(events[waited][1])
                waited += 1

        self.predict_own_windows(stage, self._guess, before_forward)
        for _, ev in events[waited:]:
            main.wait_event(ev)
        mx = self.local_max(stage)
        if self.world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        labels, total = self._after_predictor()
        out_host.copy_(labels, non_blocking=True)
        main.synchronize()
        if int(mx.item()) != int(speculative_max):              # same value on every rank: all of them redo, or none
            labels, total = self(stage)
            out_host.copy_(labels, non_blocking=True)
            main.synchronize()
        return out_host, total


def segment_lockstep(segmenters: list, image_u8: torch.Tensor):
    """All strips of one image in one process / on one GPU (test harness for the sliding-window strip logic)."""
    srcs = [image_u8[s.src0:s.src1] for s in segmenters]
    mx = torch.stack([s.local_max(x).clone() for s, x in zip(segmenters, srcs)]).max(dim=0).values
    for s, x in zip(segmenters, srcs):
        s.predict_own_windows(x, mx)
    for o, q, a, b in segmenters[0].transfers:                   # what exchange_window_rows does over NCCL
        segmenters[q].windows(a, b).copy_(segmenters[o].windows(a, b))
    for s in segmenters:
        s.blend_strip()
    labels, total = run_lockstep([s.post for s in segmenters], None)
    return torch.cat(labels), total


# ---------------------------------------------------------------------------
# bench.py --workload wsi
# ---------------------------------------------------------------------------
def synthetic_slide(H: int, W: int, seed: int = 0) -> np.ndarray:
    """Brightfield-like u8 [H,W,3] slide assembled from 1000-pixel synthetic fields, pre-normalised like
    predict.py:94-104 (outside any timed region, as in the reference's own timer)."""
    from . import synth
    from .predict import preprocess_image
    img = np.empty((H, W, 3), dtype=np.uint8)
    k = 0
    for y in range(0, H, 1000):
        for x in range(0, W, 1000):
            h, w = min(1000, H - y), min(1000, W - x)
            img[y:y + h, x:x + w] = synth.modality_image(1000, 1000, seed=seed * 1000 + k, modality=0)[:h, :w]
            k += 1
    return preprocess_image(img)[1]


def bench(args, rank, world, local, dev, metric, unit, peaks, ClockSampler):
    """One 10000 x 10000 slide over `world` GPUs (strong scaling); returns the JSON line (rank 0) or None."""
    import torch.distributed as dist
    from .unet import build_unet
    H = W = int(getattr(args, "wsi_size", 10000))
    model = build_unet(3, seed=2022).eval().to(dev)
    torch.backends.cudnn.benchmark = True
    seg = StripSegmenter(model, H, W, rank, world, chunk=args.chunk, cuda_graph=not args.no_graph, device=dev)
    s0, s1 = seg.source_rows()
    host = torch.from_numpy(np.ascontiguousarray(synthetic_slide(H, W)[s0:s1])).pin_memory()
    resident = host.to(dev)
    out_host = torch.empty((seg.r1 - seg.r0, W), dtype=torch.int32).pin_memory()
    lib = capi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    def e2e_step():
        seg.segment_host(host, out_host)                         # chunked upload behind the predictor, labels back to pinned memory

    total = 0
    for _ in range(max(args.warmup, 3)):
        _, total = seg(resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.cs_kernel_launch_count() + seg.engine.graph_kernel_launches
    ms_dev = timed(lambda: seg(resident), args.steps)
    launches = lib.cs_kernel_launch_count() + seg.engine.graph_kernel_launches - l0
    capi.profile_enable(True)                      # per-kernel durations: the same steps again with an event pair around every launch
    for _ in range(args.steps):
        seg(resident)
    torch.cuda.synchronize()
    prof = capi.profile_read()
    capi.profile_enable(False)
    seg.phase_ms = {}                              # per-phase device time of the same steps (events + one sync per step: not the timed run)
    for _ in range(args.steps):
        seg(resident)
    phases = {k: round(v / args.steps, 4) for k, v in seg.phase_ms.items()}
    seg.phase_ms = None
    barrier()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    px = H * W
    peak, peak_src = peaks
    kernels = {name: {"launches_per_step": n / args.steps, "avg_ms": round(ms / n, 5), "ms_per_step": round(ms / args.steps, 4)}
               for name, (ms, n) in prof.items()}
    strip_px = (seg.r1 - seg.r0) * W
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    win_px = seg.n_win * 65536.0
    alg = {"k_blend": (12.0 * seg.n_win * 65536 / strip_px + 0.125) * strip_px, "k_final": 4.125 * strip_px, "k_tile_fresh": 0.25 * strip_px,
           "k_u8_max": 3.0 * (s1 - s0) * W, "k_gather": 15.0 * win_px, "k_stem": 35.0 * win_px, "k_head_convt": 44.0 * win_px,
           "k_head_final": 24.0 * win_px,
           # norm blocks: MiB per 256² window over the sites each kernel serves (bench.py algorithmic_bytes) = 16 B per window pixel per MiB
           "k_nhwc_stats_view": 3.25 * 16 * win_px, "k_nhwc_norm_prelu_slices": 5.25 * 16 * win_px, "k_nhwc_norm_prelu_shuffle": 3.0 * 16 * win_px,
           "k_nhwc_stats": 3.5 * 16 * win_px, "k_nhwc_norm_prelu": 12.0 * 16 * win_px}
    ab = next((v for k, v in alg.items() if dom.startswith(k)), None)
    if ab is not None and (dom.startswith(("k_gather", "k_stem", "k_head", "k_nhwc"))):
        ab = ab / kernels[dom]["launches_per_step"]             # per launch: the per-step total is spread over the launches
    ach = round(ab / kernels[dom]["avg_ms"] / 1e6, 1) if ab else None
    return {
        "metric": metric, "value": round(px * args.steps / ms_dev / 1e3, 2), "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single {H}x{W} whole-slide synthetic image (BASELINE configs[2]), {world} row strips, NCCL boundary merge",
                   "roi": 256, "overlap": 0.25, "blend": "constant", "windows_per_forward": args.chunk, "weights": "random-init UNet seed 2022",
                   "strip_rows_rank0": [seg.r0, seg.r1], "windows_rank0": seg.n_win, "windows_exchanged": sum(b - a for _, _, a, b in seg.transfers), "instances": total,
                   "l2": "inputs larger than L2 (slide rows + window logits per rank ≫ 126 MB)"},
        "e2e": {"value": round(px * args.steps / ms_e2e / 1e3, 2), "unit": unit, "h2d_bytes_per_step": int(host.numel()) * world,
                "d2h_bytes_per_step": int(px * 4), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4) if ach else None,
                     "traffic": None, "peak_source": peak_src, "avg_launch_ms": kernels[dom]["avg_ms"], "note": "rank 0's strip"},
        "phases_ms_rank0": phases,
        "kernels": kernels,
    }
