"""F1-metric path on BASELINE config 4 (synthetic GT / prediction instance maps, ~20k cells per 4096 x 4096): device
timing of evaluate_pair's kernels and of the whole call, next to the oracle restatement of the reference's CPU code.
Usage: metric_bench.py [size] [--cpu]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cellseg_b200
from cellseg_b200 import capi, compute_metric as cm, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
gt, seg = synth.lattice_cells(size, size, seed=0), synth.lattice_cells(size, size, seed=1)
g, s = torch.from_numpy(gt).cuda(), torch.from_numpy(seg).cuda()
r = cm.evaluate_pair(g, s, 0.5)
torch.cuda.synchronize()
capi.profile_enable(True)
t0 = time.perf_counter()
for _ in range(5):
    r = cm.evaluate_pair(g, s, 0.5)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
prof = capi.profile_read(); capi.profile_enable(False)
t0 = time.perf_counter()
for _ in range(3):
    r2 = cm.evaluate_pair(gt, seg, 0.5)          # host arrays in (H2D inside)
dt_host = (time.perf_counter() - t0) / 3
px = size * size
out = {"workload": f"compute_metric on {size}x{size} GT/pred maps, {r['true_num']}/{r['pred_num']} cells after boundary removal", "result": r,
       "device_resident_ms": round(dt * 1e3, 3), "host_arrays_ms": round(dt_host * 1e3, 3), "Mpx_per_s_device": round(px / dt / 1e6, 1),
       "kernels_ms_per_call": {k: round(ms / 5, 4) for k, (ms, n) in prof.items()},
       "kernel_gbs": {}}
alg = {"k_label_flags": 4.0 * px, "k_remap": 8.0 * px, "k_overlap_sparse": 8.0 * px, "k_label_max": 4.0 * px, "k_count3": 2.0 * px}
for k, (ms, n) in prof.items():
    if k in alg:
        out["kernel_gbs"][k] = round(alg[k] / (ms / n) / 1e6, 1)
if "--cpu" in sys.argv:
    from oracle import metric as om
    t0 = time.perf_counter(); o = om.evaluate_pair(gt.copy(), seg.copy(), 0.5); dtc = time.perf_counter() - t0
    out["cpu_oracle_s"] = round(dtc, 2); out["cpu_Mpx_per_s"] = round(px / dtc / 1e6, 3)
    out["parity"] = [o[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")] == [r[k] for k in ("true_num", "pred_num", "tp", "fp", "fn")]
print(json.dumps(out))
