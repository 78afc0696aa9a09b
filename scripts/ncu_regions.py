"""Samples per code region (between marker instructions) from `ncu --page source --csv`. usage: ncu_regions.py rep kernel-regex [bin] [name-substring]"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]; rx = sys.argv[2]; binw = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
MARK = re.compile(r"LDTM|STTM|UTCHMMA|UTMALDG|UTMASTG|UBLKCP|BAR\.SYNC|SYNCS\.PHASECHK|SYNCS\.ARRIVE|MUFU\.EX2|LDS\.128|UTCBAR|ACQBULK|CCTL|NANOSLEEP|MEMBAR|ATOM|RED|EXIT")
sub = sys.argv[4] if len(sys.argv) > 4 else ""
for b in [b for b in blocks if sub in b["name"]][:1]:
    h = b["hdr"]; si = h.index("# Samples"); so = h.index("Source")
    ex = h.index("# Instructions Executed") if "# Instructions Executed" in h else None
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print("=====", b["name"][:100], "samples", tot, "instrs", len(b["rows"]))
    n = len(b["rows"])
    for s in range(0, n, binw):
        rows = b["rows"][s:s + binw]
        c = sum(int(r[si] or 0) for r in rows)
        e = sum(int(r[ex] or 0) for r in rows) if ex is not None else 0
        marks = {}
        for r in rows:
            m = MARK.search(r[so])
            if m: marks[m.group(0)] = marks.get(m.group(0), 0) + 1
        print(f"[{s:5d},{s+binw:5d}) samples {c:5d} {100*c/max(tot,1):5.1f}%  exec {e:9d}  {' '.join(f'{k}x{v}' for k,v in marks.items())}")
