"""Top stalled SASS instructions per kernel from `ncu --page source --csv` output. usage: ncu_hot.py rep [topN] [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 3: cmd += ["--kernel-name", "regex:" + sys.argv[3]]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
for b in blocks:
    h = b["hdr"]; si = h.index("# Samples"); so = h.index("Source")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print("=====", b["name"][:90], "total samples", tot)
    rows = sorted(b["rows"], key=lambda r: -int(r[si] or 0))[:topn]
    for r in rows:
        st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        idx = b["rows"].index(r)
        print(f"{int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}%  #{idx:5d} {r[so].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
