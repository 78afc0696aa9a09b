"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: launch_list_summary.py X [out.csv]"""
import csv, sys, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
tot = sum(u for _, u in rows); agg = collections.OrderedDict()
for k, u in rows:
    a = agg.setdefault(k[:70], [0, 0.0]); a[0] += 1; a[1] += u
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
w = csv.writer(out); w.writerow(["kernel", "launches", "total_us", "share_pct"])
for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, round(u, 1), round(100 * u / tot, 2)])
ours = sum(u for k, (n, u) in agg.items() if "dsc::" in k or "xattn" in k or "dpmpp" in k or "region_" in k)
w.writerow(["# launches", len(rows), round(tot, 1), 100.0]); w.writerow(["# dsc kernels share", "", round(ours, 1), round(100 * ours / tot, 2)])
