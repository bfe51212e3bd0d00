#!/usr/bin/env python
"""Aggregates `ncu --page source --print-source sass,cuda --csv` output per CUDA source line:
stall samples, warp instructions executed, dominant stall reasons. Usage: ncu_lines.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue      # SASS rows carry an address; source-line summary rows have "-"
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except (KeyError, ValueError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    agg[(cur, int(r[0]))] = (samples, inst, r[1].strip()[:110], stalls)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {ti}")
print("---- by stall samples")
for (f, l), (s, i, src, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    dom = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{100*s/ts:6.2f}% smp {100*i/ti:6.2f}% ins  {f}:{l:<4d} {src}   [{dom}]")
print("---- by instructions")
for (f, l), (s, i, src, st) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{100*i/ti:6.2f}% ins {100*s/ts:6.2f}% smp  {f}:{l:<4d} {src}")
