#!/usr/bin/env python
"""Like ncu_lines.py, for a report that holds several kernels: aggregates the source page per CUDA source line of the
kernels whose function name contains FILTER. Usage: ncu_lines_fn.py file.csv FILTER [top]"""
import csv, sys
csv.field_size_limit(1 << 30)
flt = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = None; hdr = None; fn = ""
agg = {}
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or flt not in fn: continue
    if r[2] != "-": continue
    d = dict(zip(hdr[4:], r[4:]))
    try: samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except (KeyError, ValueError): continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    exc = int(d.get("L1 Wavefronts Shared Excessive", "0") or 0); wf = int(d.get("L1 Wavefronts Shared", "0") or 0)
    key = (cur, int(r[0]))
    if key in agg:
        a = agg[key]; st = dict(a[3])
        for k, v in stalls.items(): st[k] = st.get(k, 0) + v
        agg[key] = (a[0] + samples, a[1] + inst, a[2], st, a[4] + exc, a[5] + wf)
    else: agg[key] = (samples, inst, r[1].strip()[:100], stalls, exc, wf)
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1; tw = sum(v[5] for v in agg.values()) or 1
print(f"filter {flt}: total samples {ts}, warp instructions {ti}, shared wavefronts {tw} (excessive {sum(v[4] for v in agg.values())})")
print("---- by stall samples")
for (f, l), (s, i, src, st, exc, wf) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    dom = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{100*s/ts:6.2f}% smp {100*i/ti:6.2f}% ins {100*wf/tw:5.1f}% wf ({exc} exc) {f}:{l:<4d} {src}   [{dom}]")
print("---- by instructions")
for (f, l), (s, i, src, st, exc, wf) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{100*i/ti:6.2f}% ins {100*s/ts:6.2f}% smp  {f}:{l:<4d} {src}")
