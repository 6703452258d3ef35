"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
Usage: ncu_lines.py file.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = {}
fname = ""
hdr = None
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue            # SASS rows carry an address; summary rows per source line have "-"
    ii, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        lines[(fname, int(r[0]))] = (r[1], int(r[ii]), int(r[ss]))
    except ValueError:
        pass
ti = sum(v[1] for v in lines.values()) or 1
ts = sum(v[2] for v in lines.values()) or 1
print(f"total warp-instructions {ti}, samples {ts}")
for (f, ln), (src, i, s) in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*i/ti:5.1f}% inst {100*s/ts:5.1f}% smp  {f}:{ln:<4d} {src.strip()[:110]}")
