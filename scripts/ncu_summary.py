"""Summarise an .ncu-rep (raw page) into a small CSV/markdown: python scripts/ncu_summary.py rep [out.csv]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','smsp__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum']
stalls = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
lines = []
for k in keep + stalls:
    if k in hdr:
        i = hdr.index(k)
        lines.append([k, units[i]] + [r[i] for r in rows[2:]])
w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(rows) - 2)])
for l in lines:
    if l[0] in stalls:
        try:
            if max(float(x) for x in l[2:]) < 0.05: continue
        except ValueError: pass
    w.writerow(l)
