#!/usr/bin/env python
"""bench.py — Mpixel/s, image → instance map (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload batch|wsi]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path (baseline/predict.py:107-113: scale → sliding-window gather → UNet
(PyTorch/cuDNN, out of scope) → blend + softmax + threshold → hole filling → small-object removal → labelling)
over one batch of synthetic images.  Workload `batch` (default) = BASELINE configs[1]: 64 synthetic 1024x1024x3
u8 images per GPU (image-sharded, no data-path collective → weak scaling); `wsi` = configs[2]: one 10000x10000
image, row strips over the ranks with the NCCL boundary-label merge (strong scaling).

  value        whole-job Mpixel/s with the u8 images already resident in HBM (CUDA events, max over ranks)
  e2e          same through SegmentationEngine.segment_host: pinned host u8 in, int32 labels to pinned host out,
               both copies of every step inside the timed region; the K calls are queued back to back (wait=False)
               and drained once before the timer stops, so upload / compute / download overlap across calls
  roofline     the dominant libcellseg kernel: algorithmic bytes / its live CUDA-event duration vs the measured
               HBM peak of MEASURED_PEAKS.json;  `kernels` lists every libcellseg kernel the same way
  cpu_baseline the oracle pipeline (oracle/pipeline.py: restated MONAI sliding window + torch-CPU UNet +
               scipy.ndimage morphology) on a bounded sample of the same images, rank 0, N=1
  --impl reference   times that CPU pipeline alone (the reference's own implementation cannot run here: monai /
               skimage are not installable offline), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "image_to_instance_map_throughput"
UNIT = "Mpixel/s"
BATCH_IMAGES, BATCH_HW = 64, 1024
# sub-batch schedule of the end-to-end call: with calls pipelined across their boundaries the short first / last sub-batches of the
# ramp are not needed (the fill and the drain happen once per timed region, not once per step)
E2E_RAMP = os.environ.get("CS_E2E_RAMP", "0") == "1"
E2E_SUB = int(os.environ.get("CS_E2E_SUB", "32"))          # images per sub-batch of the end-to-end call (measured: gap to the device-resident step 1.0 ms at 16, 0.65 ms at 32 and 64)
WSI_HW = 10000


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synthetic_batch(n_images: int, hw: int, first_seed: int) -> np.ndarray:
    """Pre-normalised u8 [n,hw,hw,3] images: the four modality generators, then the reference's
    channel fix-up + percentile normalisation (predict.py:94-104) — outside the timed region like the reference's
    own timer (predict.py:106)."""
    from cellseg_b200 import synth
    from cellseg_b200.predict import preprocess_image
    out = np.empty((n_images, hw, hw, 3), dtype=np.uint8)
    for i in range(n_images):
        img = synth.modality_image(hw, hw, seed=first_seed + i)
        out[i] = preprocess_image(img)[1]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def chain_alone(masks_u8, steps: int = 10) -> dict:
    """Device time of the mask → instance-map chain alone (hole filling, object filter, labelling: SURVEY.md §8 a-5..a-7) on a
    resident batch of masks: one call packs the masks into the workspace, then cs_postprocess_packed is timed back to back with
    CUDA events on the launching stream.  Fractions against the measured HBM peak at SURVEY.md §8d's algorithmic 9 B/px and at
    the bytes the bit-packed chain really has to move (1/8 B in, 4 B out)."""
    import ctypes as C
    from cellseg_b200 import capi
    lib = capi.load()
    N, H, W = masks_u8.shape
    px = N * H * W
    wsb = lib.cs_postprocess_workspace_bytes(N, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=masks_u8.device)
    labels = torch.empty((N, H, W), dtype=torch.int32, device=masks_u8.device)
    counts = torch.empty((N,), dtype=torch.int32, device=masks_u8.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(lib, lib.cs_postprocess(masks_u8.data_ptr(), N, H, W, 64, 16, 8, labels.data_ptr(), counts.data_ptr(), ws.data_ptr(), wsb, st))

    def run():
        capi.check(lib, lib.cs_postprocess_packed(N, H, W, 64, 16, 8, labels.data_ptr(), counts.data_ptr(), ws.data_ptr(), wsb, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(); run(); b.record()
    torch.cuda.synchronize()
    ms = statistics.median(a.elapsed_time(b) for a, b in ev)
    peak, _ = measured_peaks()
    return {"ms": round(ms, 4), "instances": int(counts.sum().item()), "gbs_at_9Bpx": round(9 * px / ms / 1e6, 1), "frac_at_9Bpx": round(9 * px / ms / 1e6 / peak, 4),
            "gbs_at_4.125Bpx": round(4.125 * px / ms / 1e6, 1), "frac_at_4.125Bpx": round(4.125 * px / ms / 1e6 / peak, 4)}


def algorithmic_bytes(kernel: str, launches: int, px: int, rho: float, win_px_total: int, n_steps: int) -> float | None:
    """Algorithmic bytes PER LAUNCH of a libcellseg kernel.  px = image
    pixels per step, win_px_total = window pixels per step, launches = launches over n_steps steps."""
    per_step = launches / n_steps
    k = kernel.split("<")[0]
    for prefix in ("k_blend", "k_tile_fresh", "k_tile_resume"):
        if k.startswith(prefix):
            k = prefix
    table = {
        "k_u8_max": 3.0 * px,                          # read the u8 image once
        "k_gather": 15.0 * win_px_total,               # 3 B u8 read + 12 B f32 write per window pixel
        "k_blend": (12.0 * rho + 0.125) * px,          # every window logit once + 1 bit / pixel
        "k_tile_fresh": 0.25 * px,                     # packed mask in + out (every tile)
        "k_final": 4.125 * px,                         # packed mask in, int32 labels out
        "k_pack_u8": 1.125 * px,
        # fused InstanceNorm+PReLU(+residual) of the UNet's ADN blocks: x read + y write (+ residual read), summed over
        # the 17 blocks of one window forward: 20 316 160 B in register-resident planes, 1 572 864 B in the
        # 3 x 256 x 256 head that streams through L2
        "k_instnorm_prelu_reg": 20316160.0 * (win_px_total / 65536),
        "k_instnorm_prelu_stream": 1572864.0 * (win_px_total / 65536),
        # channels-last predictor glue, per window pixel: stem = 3 B of image bytes in + two 16-channel
        # half-resolution tensors out; head = 32 channels at half resolution in, 3 planes out; final = 3 planes in + out.
        "k_stem": (3.0 + 32.0) * win_px_total,
        "k_head_convt": (32.0 + 12.0) * win_px_total,
        "k_head_final": 24.0 * win_px_total,
    }
    # Norm blocks: MiB of activations read + written per 256² window, summed over the sites each kernel serves (fp32 NHWC sizes
    # per window: 128²x16 = 1 MiB, 64²x32 = 0.5, 32²x64 = 0.25, 16²x128 = 0.125, 16²x256 = 0.25; DESIGN.md §3.3 lists the sites):
    #   statistics read x once; the apply pass reads x (+ the residual) and writes y (dense and/or the concatenation slot).
    norm_mib = {"k_nhwc_stats": 3.5, "k_nhwc_norm_prelu": 12.0, "k_nhwc_stats_view": 3.25, "k_nhwc_norm_prelu_slices": 5.25,
                "k_nhwc_norm_prelu_shuffle": 3.0}
    for name, mib in norm_mib.items():
        table[name] = mib * 1048576 * (win_px_total / 65536)
    if k not in table:
        return None
    return table[k] / per_step


# ---------------------------------------------------------------------------
# CPU pipeline (oracle) — cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------
def cpu_pipeline_rate(images_u8: np.ndarray, n_sample: int, steps: int, warmup: int, seed: int = 2022):
    """Mpixel/s of oracle/pipeline.segment over `n_sample` images per step on all host threads."""
    from cellseg_b200.unet import build_unet
    from oracle import pipeline as opl
    model = build_unet(3, seed=seed).eval()
    sample = images_u8[:n_sample]
    px = sample.shape[0] * sample.shape[1] * sample.shape[2]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for img in sample:
            opl.segment(img, model)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return px / 1e6 / statistics.mean(times), statistics.mean(times) * 1e3, torch.get_num_threads()


def batch_config(args) -> dict:
    """`config` of the batch workload — the same object on both arms (the reference arm runs a bounded sample of it, described
    in its `cpu_baseline.sample`)."""
    return {"workload": f"batch of {args.images} synthetic 1024x1024x3 u8 images per GPU (BASELINE configs[1]), image-sharded",
            "roi": 256, "overlap": 0.25, "blend": "constant", "weights": "random-init UNet seed 2022",
            "l2": "inputs larger than L2 (192 MB u8 images, 1.26 GB window logits per step)"}


def run_reference(args, rank):
    """The reference's CPU implementation of the path (oracle port: monai / skimage are not installable offline) on all host
    threads: exactly --warmup untimed and --steps timed steps, each step a bounded sample (2 images) of the workload."""
    if rank != 0:
        return
    n_sample = 2
    imgs = synthetic_batch(n_sample, BATCH_HW, 0)
    rate, ms, cores = cpu_pipeline_rate(imgs, n_sample, args.steps, args.warmup)
    sample = (f"{n_sample} of the {args.images} 1024x1024 images per step (seeds 0..{n_sample - 1}); the CPU pipeline handles the images "
              f"one at a time with sw_batch_size 4 like baseline/predict.py:83-109")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": batch_config(args),
        "cpu_baseline": {"value": round(rate, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(rate, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of predict.py:107-113 on host cores (monai/skimage not installable offline; see DESIGN.md)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------
def run_ours(args, rank, world, local):
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from cellseg_b200 import hostutil
        hostutil.bind_to_gpu_numa(local)           # pinned staging buffers on the GPU's own socket (before they are allocated)
    from cellseg_b200 import capi
    from cellseg_b200.sliding_window import SegmentationEngine, make_plan
    from cellseg_b200.unet import build_unet
    lib = capi.load()

    if args.workload == "wsi":
        return run_wsi(args, rank, world, local, dev)
    if args.workload == "sweep":
        return run_sweep(args, rank, world, local, dev)

    n_img, hw = args.images, BATCH_HW
    host = torch.from_numpy(synthetic_batch(n_img, hw, rank * n_img)).pin_memory()
    model = build_unet(3, seed=2022).eval().to(dev)
    torch.backends.cudnn.benchmark = True
    engine = SegmentationEngine(model, chunk=args.chunk, channels_last=args.channels_last, cuda_graph=not args.no_graph)
    plan = make_plan(hw, hw, 256, 0.25)
    px = n_img * hw * hw
    rho = plan.n_win * 65536 / (hw * hw)
    resident = host.to(dev)
    labels_host = torch.empty((n_img, hw, hw), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
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
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(max(args.warmup, 3)):           # timing rules: at least 3 warm-up steps whatever was asked
        engine(resident)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # -- device-resident
    l0 = lib.cs_kernel_launch_count() + engine.graph_kernel_launches
    if os.environ.get("CS_NCU_RANGE"):             # `ncu --profile-from-start off`: capture the timed region only
        torch.cuda.profiler.start()
    ms_dev = timed(lambda: engine(resident), args.steps)
    if os.environ.get("CS_NCU_RANGE"):
        torch.cuda.profiler.stop()
    launches = lib.cs_kernel_launch_count() + engine.graph_kernel_launches - l0
    # -- per-kernel durations: the same step launched eagerly (a graph replay has no per-launch hooks) with a CUDA event
    #    pair around every libcellseg launch, on the launching stream
    engine._call_eager(resident)
    torch.cuda.synchronize()
    capi.profile_enable(True)
    for _ in range(args.steps):
        engine._call_eager(resident)
    torch.cuda.synchronize()
    prof = capi.profile_read()
    capi.profile_enable(False)
    # -- end to end with host buffers
    # The K timed calls are queued back to back (wait=False) and drained once, inside the timed region: a stream of batches keeps the
    # upload, compute and download streams busy across call boundaries.  Two output buffers alternate, as a consumer of the results would
    # need; every step's host→device and device→host copies are inside the timed region.
    labels_host2 = torch.empty_like(labels_host).pin_memory()
    outs = (labels_host, labels_host2)
    for i in range(2):
        engine.segment_host(host, outs[i % 2])

    def e2e_steps():
        for i in range(args.steps):
            engine.segment_host(host, outs[(i + 1) % 2], sub_batch=E2E_SUB, ramp=E2E_RAMP, wait=False)
        engine.wait()                                          # the last download has landed

    for _ in range(1):
        e2e_steps()
    ms_e2e = timed(e2e_steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    # self-check: the instance maps the timed end-to-end call left in host memory against an eager device-resident call
    # (sub-batches hand cuDNN other batch sizes, so a knife-edge pixel may differ; anything more is a bug)
    got = outs[args.steps % 2].to(dev)                        # what the last timed call delivered
    ref_labels = engine._call_eager(resident)[0]
    mismatch = float((got != ref_labels).float().mean().item())
    # … and against eager calls on the very sub-batches the pipeline used (same shapes → same cuDNN algorithms): must be identical
    same, s0 = 0, 0
    for n in engine._schedule(n_img, min(E2E_SUB, n_img), E2E_RAMP):
        same += int((got[s0:s0 + n] != engine._call_eager(resident[s0:s0 + n])[0]).sum().item())
        s0 += n

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    win_px_total = n_img * plan.n_win * 65536
    kernels = {}
    for name, (ms, n) in prof.items():
        ab = algorithmic_bytes(name, n, px, rho, win_px_total, args.steps)
        avg_ms = ms / n
        kernels[name] = {"launches_per_step": n / args.steps, "avg_ms": round(avg_ms, 5), "ms_per_step": round(ms / args.steps, 4),
                         "alg_bytes_per_launch": ab, "achieved_gbs": round(ab / avg_ms / 1e6, 1) if ab else None}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    ours_ms = sum(v["ms_per_step"] for v in kernels.values())
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(dom.split("<")[0])
    except Exception:
        pass
    d = kernels[dom]
    glue = ("k_stem", "k_nhwc", "k_head", "k_instnorm")
    chain = [k for k in kernels if not k.startswith(("k_u8_max", "k_scale_lut", "k_gather", "k_blend") + glue)]
    chain_ms = sum(kernels[k]["ms_per_step"] for k in chain)
    line = {
        "metric": METRIC, "value": round(world * px * args.steps / ms_dev / 1e3, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": batch_config(args),
        "implementation": {"windows_per_forward": args.chunk,
                           "unet": "cuDNN fp32 (TF32, torch default) convolutions on NHWC tensors; 3-channel stem/head convolutions, gather and "
                                   "all norm blocks in libcellseg" + ("; the whole step replayed as one CUDA graph" if not args.no_graph else "")},
        "e2e": {"value": round(world * px * args.steps / ms_e2e / 1e3, 2), "unit": UNIT, "h2d_bytes_per_step": int(host.numel()),
                "d2h_bytes_per_step": int(labels_host.numel() * 4), "ms_per_step": round(ms_e2e / args.steps, 3),
                "how": "K segment_host(wait=False) calls on pinned host buffers queued back to back, one engine.wait() before the timer stops"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": round(d["achieved_gbs"] / peak, 4) if d["achieved_gbs"] else None, "traffic": traffic,
                     "peak_source": peak_src, "avg_launch_ms": d["avg_ms"]},
        "checks": {"e2e_vs_eager_same_subbatches_mismatch_pixels": same, "e2e_vs_eager_whole_batch_mismatch_fraction": mismatch},
        "kernels": kernels,
        "stages": {"libcellseg_ms_per_step": round(ours_ms, 4), "cudnn_convs_and_cat_ms_per_step": round(ms_dev / args.steps - ours_ms, 3),
                   "predictor_glue_ms_per_step": round(sum(v["ms_per_step"] for k, v in kernels.items() if k.startswith(glue)), 4),
                   # algorithmic bytes of the mask -> instance-map chain: SURVEY.md §8d's 9 B/px (holes 1+1, objects 1+1, label 1+4 on byte
                   # masks) is the figure the fraction is quoted on; the bit-packed chain itself moves 4.125 B/px, reported next to it
                   "postprocess_chain": {"ms_per_step": round(chain_ms, 4), "alg_bytes_per_px": 9.0,
                                         "achieved_gbs": round(9 * px / chain_ms / 1e6, 1) if chain_ms else None,
                                         "frac": round(9 * px / chain_ms / 1e6 / peak, 4) if chain_ms else None,
                                         "achieved_gbs_at_packed_4.125Bpx": round(4.125 * px / chain_ms / 1e6, 1) if chain_ms else None,
                                         "frac_at_packed_4.125Bpx": round(4.125 * px / chain_ms / 1e6 / peak, 4) if chain_ms else None,
                                         "masks": "the step's own predictor masks (random-init UNet: salt-like noise, the labelling worst case; "
                                                  "cell-like masks under postprocess_chain_alone)"}},
    }
    if world == 1:
        # the chain alone, timed as one unit (no per-kernel events): on the masks this step's predictor produced and on cell-like
        # masks — a random-init predictor yields salt-like masks (~1000 specks per 64x128 tile), the labelling worst case
        from cellseg_b200 import synth
        pm = torch.empty((n_img, hw, hw), dtype=torch.uint8, device=dev)
        from cellseg_b200.sliding_window import blend as _blend
        logits, bufs = engine.window_logits(resident)
        pm.copy_(_blend(logits, n_img, bufs["plan"], None, want=("mask",))["mask"])
        line["stages"]["postprocess_chain_alone"] = {
            "note": "mask -> instance map chain timed as one unit, 64 x 1024^2, packed mask resident; 9 B/px = SURVEY 8d's algorithmic bytes",
            "predictor_masks": chain_alone(pm),
            "sparse_cell_field": chain_alone(torch.from_numpy(np.stack([synth.lattice_cells(hw, hw, seed=s_, pitch=200, radius=14) > 0
                                                                        for s_ in range(n_img)]).view(np.uint8)).to(dev)),
            "dense_blobs": chain_alone(torch.from_numpy(np.stack([synth.blob_mask(hw, hw, s_) for s_ in range(n_img)]).view(np.uint8)).to(dev)),
        }
    if world == 1 and not args.no_cpu_baseline:
        n_sample, passes = min(8, n_img), 8                 # ~10 s of CPU work on the box's host cores (a bounded sample)
        rate, ms, cores = cpu_pipeline_rate(host[:n_sample].numpy(), n_sample, passes, 1)
        line["cpu_baseline"] = {"value": round(rate, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_sample} of the {n_img} images, {passes} timed passes after 1 warm-up ({ms / 1e3:.1f} s per pass)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_sweep(args, rank, world, local, dev):
    """BASELINE configs[4]: mixed-modality size sweep (512² … 4096², four modality generators), Gaussian blend,
    image-sharded (every rank runs the whole sweep on its own seeds → weak scaling)."""
    from cellseg_b200 import capi
    from cellseg_b200.sliding_window import SegmentationEngine
    from cellseg_b200.unet import build_unet
    lib = capi.load()
    sizes = [512, 1024, 2048, 4096]
    model = build_unet(3, seed=2022).eval().to(dev)
    torch.backends.cudnn.benchmark = True
    engine = SegmentationEngine(model, chunk=args.chunk, mode="gaussian", cuda_graph=not args.no_graph)
    hosts = [torch.from_numpy(synthetic_batch(4, hw, rank * 100 + 10 * i)).pin_memory() for i, hw in enumerate(sizes)]
    residents = [h.to(dev) for h in hosts]
    outs = [torch.empty(h.shape[:3], dtype=torch.int32).pin_memory() for h in hosts]
    px = sum(h.shape[0] * h.shape[1] * h.shape[2] for h in hosts)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
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
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    def dev_step():
        for r in residents:
            engine(r)

    def host_step():
        for h, o in zip(hosts, outs):
            engine.segment_host(h, o, sub_batch=2, wait=False)      # the four size groups pipeline into each other
        engine.wait()

    for _ in range(max(args.warmup, 3)):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.cs_kernel_launch_count() + engine.graph_kernel_launches
    ms_dev = timed(dev_step, args.steps)
    launches = lib.cs_kernel_launch_count() + engine.graph_kernel_launches - l0
    # -- per-kernel durations: the same step launched eagerly (a graph replay has no per-launch hooks) with a CUDA event
    #    pair around every libcellseg launch, on the launching stream
    capi.profile_enable(True)
    for _ in range(args.steps):
        for r in residents:
            engine._call_eager(r)
    torch.cuda.synchronize()
    prof = capi.profile_read()
    capi.profile_enable(False)
    host_step()
    ms_e2e = timed(host_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peaks()
        kernels = {name: {"launches_per_step": n / args.steps, "avg_ms": round(ms / n, 5), "ms_per_step": round(ms / args.steps, 4)}
                   for name, (ms, n) in prof.items()}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        win_bytes = sum(h.shape[0] * len(__import__("cellseg_b200.sliding_window", fromlist=["make_plan"]).make_plan(h.shape[1], h.shape[2]).ys) ** 2
                        for h in hosts) * 65536 * 12.0
        ach = round((win_bytes + px / 8) / kernels[dom]["ms_per_step"] / 1e6, 1) if dom.startswith("k_blend") else None
        print(json.dumps({
            "metric": METRIC, "value": round(world * px * args.steps / ms_dev / 1e3, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "mixed-modality size sweep 512²/1024²/2048²/4096² x 4 modalities per GPU (BASELINE configs[4]), Gaussian blend, image-sharded",
                       "roi": 256, "overlap": 0.25, "blend": "gaussian (monai 0.9-1.1 map)", "windows_per_forward": args.chunk,
                       "weights": "random-init UNet seed 2022", "l2": "inputs larger than L2 (window logits 12 MB … 1.4 GB per size group)"},
            "e2e": {"value": round(world * px * args.steps / ms_e2e / 1e3, 2), "unit": UNIT, "h2d_bytes_per_step": int(sum(h.numel() for h in hosts)),
                    "d2h_bytes_per_step": int(px * 4), "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4) if ach else None,
                         "traffic": None, "peak_source": peak_src, "note": "summed over the four size groups of a step"},
            "kernels": kernels}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_wsi(args, rank, world, local, dev):
    from cellseg_b200 import wsi
    line = wsi.bench(args, rank, world, local, dev, METRIC, UNIT, measured_peaks(), ClockSampler)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="batch", choices=("batch", "wsi", "sweep"))
    ap.add_argument("--images", type=int, default=BATCH_IMAGES, help="images per GPU per step (batch workload)")
    ap.add_argument("--chunk", type=int, default=400, help="windows per UNet forward (400 = one forward per 16-image sub-batch)")
    ap.add_argument("--channels_last", action="store_true")
    ap.add_argument("--no_graph", action="store_true", help="do not capture the UNet forward in a CUDA graph")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--wsi_size", type=int, default=WSI_HW, help="edge of the square whole-slide image (wsi workload)")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        return run_reference(args, rank)
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
