#!/usr/bin/env python
"""Top SASS instructions by stall samples of the kernels whose name contains FILTER, from `ncu --page source --csv
--print-source cuda,sass`. Usage: ncu_sass_top.py file.csv FILTER [top]"""
import csv, sys
csv.field_size_limit(1 << 30)
flt = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
fn = ""; hdr = None; cur = None; line = 0; seen = {}
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or flt not in fn: continue
    if r[2] == "-": line = int(r[0]); continue
    d = dict(zip(hdr[4:], r[4:]))
    seen.setdefault(r[2], (cur, line, r[3], d))
def num(d, k):
    try: return int(d.get(k, "0") or 0)
    except ValueError: return 0
tot = sum(num(v[3], '# Samples') for v in seen.values()) or 1
print('instructions', len(seen), 'total samples', tot)
for addr, (cur, l, sass, d) in sorted(seen.items(), key=lambda x: -num(x[1][3], '# Samples'))[:top]:
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v)}
    dom = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:2])
    print(f"{100*num(d,'# Samples')/tot:5.2f}% {cur}:{l} {addr[-5:]} {sass[:80]:80s} [{dom}]")
