"""profiles/traffic.json from the ncu summaries (tools/ncu_summary.py output): dram__bytes_read.sum + dram__bytes_write.sum per launch,
averaged over the captured launches of each kernel, under the names bench.py's `kernels` object uses.
Usage: traffic_from_ncu.py profiles/r02_ncu_step_glue.txt profiles/r02_ncu_step_post.txt > profiles/traffic.json"""
import json, re, sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = [  # (regex on the demangled kernel line, bench.py name)
    (r"k_nhwc_norm_prelu_view<\(?bool\)?1>|k_nhwc_norm_prelu_view<1>", "k_nhwc_norm_prelu_shuffle"),
    (r"k_nhwc_norm_prelu_view<\(?bool\)?0>|k_nhwc_norm_prelu_view<0>", "k_nhwc_norm_prelu_slices"),
    (r"k_nhwc_norm_prelu\(", "k_nhwc_norm_prelu"), (r"k_nhwc_stats_view", "k_nhwc_stats_view"), (r"k_nhwc_stats\(", "k_nhwc_stats"),
    (r"k_stem<", "k_stem"), (r"k_head_convt", "k_head_convt"), (r"k_head_final", "k_head_final"), (r"k_blend", "k_blend_mask"),
    (r"k_certify", "k_certify"), (r"k_tile_resume<1", "k_tile_resume1_label8"), (r"k_tile_resume<2", "k_tile_resume2_label8"),
    (r"k_tile<", "k_tile_fresh_label8"), (r"k_filter_merge", "k_filter_merge"), (r"k_filter_accum", "k_filter_accum"),
    (r"k_label_merge", "k_label_merge"), (r"k_label_flatten", "k_label_flatten"), (r"k_scan_lookback", "k_scan_lookback"),
    (r"k_final", "k_final"), (r"k_u8_max", "k_u8_max"),
]
acc = defaultdict(list)
for path in sys.argv[1:]:
    cur, rd = None, None
    for line in open(path):
        if line.startswith("----"):
            cur = next((n for rx, n in NAMES if re.search(rx, line)), None)
            rd = None
            continue
        m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([\d.]+)\s+(\w+)", line)
        if m and cur:
            v = float(m.group(2)) * UNIT[m.group(3)]
            if m.group(1) == "read":
                rd = v
            elif rd is not None:
                acc[cur].append(rd + v)
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the captured launches of one eager step, 64 x 1024^2 "
                "images, 400 windows per forward) from the ncu --set full summaries " + ", ".join(p.split("/")[-1] for p in sys.argv[1:])}
for k in sorted(acc):
    out[k] = int(sum(acc[k]) / len(acc[k]))
print(json.dumps(out, indent=1))
