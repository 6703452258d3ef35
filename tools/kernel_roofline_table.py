"""One table per round: every libcellseg kernel of the bench step with its live CUDA-event time, algorithmic bytes, achieved GB/s and
fraction of the measured HBM peak (bench line), next to the DRAM bytes and DRAM utilisation ncu measured for the same kernel
(profiles/traffic.json, ncu summaries).  Usage: kernel_roofline_table.py profiles/r02_bench_batch_1gpu.json profiles/traffic.json
profiles/r02_ncu_step_glue.txt profiles/r02_ncu_step_post.txt > profiles/r02_kernel_roofline.md"""
import json, re, sys
from collections import defaultdict

bench = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
traffic = json.load(open(sys.argv[2]))
peak = bench["roofline"]["peak"]
util = defaultdict(list)
dur = defaultdict(list)
for path in sys.argv[3:]:
    cur = None
    for line in open(path):
        if line.startswith("----"):
            cur = line[5:].strip()
            continue
        m = re.match(r"\s+gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed\s+([\d.]+)", line)
        if m and cur:
            util[cur].append(float(m.group(1)))
        m = re.match(r"\s+gpu__time_duration\.sum\s+([\d.]+)\s+(\w+)", line)
        if m and cur:
            dur[cur].append(float(m.group(1)) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(m.group(2), 1.0))

def ncu_for(name):
    pats = {"k_nhwc_norm_prelu_shuffle": r"k_nhwc_norm_prelu_view<\(?bool\)?1>|norm_prelu_view<1>", "k_nhwc_norm_prelu_slices": r"norm_prelu_view<\(?bool\)?0>|norm_prelu_view<0>",
            "k_nhwc_norm_prelu": r"k_nhwc_norm_prelu\(", "k_nhwc_stats_view": r"k_nhwc_stats_view", "k_nhwc_stats": r"k_nhwc_stats\(",
            "k_stem<true>": r"k_stem<", "k_blend_mask": r"k_blend", "k_tile_fresh_label8": r"k_tile<", "k_tile_resume1_label8": r"k_tile_resume<1",
            "k_tile_resume2_label8": r"k_tile_resume<2"}
    rx = pats.get(name, re.escape(name.split("<")[0]) + r"[(<]")
    u = [x for k, v in util.items() if re.search(rx, k) for x in v]
    d = [x for k, v in dur.items() if re.search(rx, k) for x in v]
    return (sum(u) / len(u) if u else None, sum(d) / len(d) if d else None)

print(f"# Kernel roofline table — {bench['config']['workload']}\n")
print(f"Bench line: {bench['value']} {bench['unit']} device-resident, {bench['ms_per_step']} ms per step, e2e {bench['e2e']['value']} {bench['unit']}; "
      f"peak = {peak} GB/s (MEASURED_PEAKS.json hbm_gbs).  `ms/step`, `avg µs` = CUDA events around every launch of an eager replica of the step "
      f"(bench.py); `alg MB` = algorithmic bytes per launch (DESIGN.md §3); `DRAM MB`, `DRAM util`, `ncu µs` = ncu --set full captures of the same "
      f"kernels (cold caches, serialised), averaged over the captured launches.\n")
print("| kernel | launches/step | ms/step | avg µs | alg MB/launch | achieved GB/s | % of peak | DRAM MB/launch (ncu) | DRAM util (ncu) | ncu µs |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, k in sorted(bench["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    ab, ach = k.get("alg_bytes_per_launch"), k.get("achieved_gbs")
    tkey = name.replace("<true>", "").replace("<false>", "")
    tr = traffic.get(tkey) or traffic.get(name)
    u, d = ncu_for(name)
    print(f"| `{name}` | {k['launches_per_step']:.0f} | {k['ms_per_step']:.4f} | {k['avg_ms'] * 1e3:.1f} | "
          f"{ab / 1e6:.1f} | " if ab else f"| `{name}` | {k['launches_per_step']:.0f} | {k['ms_per_step']:.4f} | {k['avg_ms'] * 1e3:.1f} | — | ", end="")
    print(f"{ach:.0f} | {100 * ach / peak:.0f} % | " if ach else "— | — | ", end="")
    print(f"{tr / 1e6:.1f} | " if tr else "— | ", end="")
    print(f"{u:.0f} % | " if u is not None else "— | ", end="")
    print(f"{d:.1f} |" if d is not None else "— |")
st = bench["stages"]
print(f"\nSums per step: libcellseg kernels {st['libcellseg_ms_per_step']} ms (predictor glue {st['predictor_glue_ms_per_step']} ms), cuDNN convolutions "
      f"{st['cudnn_convs_and_cat_ms_per_step']} ms; mask → instance-map chain on the step's own masks {st['postprocess_chain']['ms_per_step']} ms "
      f"({100 * st['postprocess_chain']['frac']:.0f} % of the peak at SURVEY §8d's 9 B/px).")
