"""Print the key metrics of an .ncu-rep (raw page) per kernel. Usage: ncu_summary.py file.ncu-rep [more metric substrings]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct'] + sys.argv[2:]
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
ki = hdr.index('Kernel Name')
for r in data:
    print('----', r[ki][:110])
    for w in want:
        for i, h in enumerate(hdr):
            if h == w or (w in sys.argv[2:] and w in h):
                print(f"  {h:75s} {r[i]:>18s} {units[i]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
    print("  top stalls:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in st))
