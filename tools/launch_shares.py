"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time share per kernel, ours vs libraries.
Usage: launch_shares.py launches.csv [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, n = collections.OrderedDict(), 0
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '')[:80]
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] in ('ns', 'nsecond') else v * 1000 if r[ui] in ('ms', 'msecond') else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; n += 1
tot = sum(a[1] for a in agg.values())
ours = [k for k in agg if k.startswith(('k_', 'cs::'))]
print(f"launches {n}, summed kernel time {tot / 1000:.2f} ms; libcellseg kernels {100 * sum(agg[k][1] for k in ours) / tot:.1f} % of it "
      f"({sum(agg[k][0] for k in ours)} launches)")
print("| share | launches | avg us | kernel |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| {100 * t / tot:5.1f} % | {c} | {t / c:.1f} | {'**' + k + '**' if k in ours else k} |")
