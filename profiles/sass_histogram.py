#!/usr/bin/env python
"""Instruction histogram per kernel of the built library: `cuobjdump -sass librnsc.so` grouped by mnemonic, with the
mnemonics that prove the sm_100a features in use (UBLKCP = TMA bulk copy, SYNCS = mbarrier, VIADD.16x2, REDUX ...).
Usage: sass_histogram.py path/to/librnsc.so > profiles/rN_sass_histogram.txt"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True, check=True).stdout
fn = None; hist = collections.OrderedDict(); arch = set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip(); hist[fn] = collections.Counter(); continue
    m = re.search(r"arch = (\S+)", line)
    if m: arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn: hist[fn][m.group(1)] += 1
print("arch:", ", ".join(sorted(arch)))
for fn, h in hist.items():
    tot = sum(h.values())
    if tot < 40: continue
    base = collections.Counter()
    for k, v in h.items(): base[k.split(".")[0]] += v
    feats = {k: v for k, v in h.items() if k.split(".")[0] in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "REDUX", "BAR", "ATOMS", "ATOMG", "RED", "MATCH", "VOTE", "SHFL", "LDL", "STL") or ".16x2" in k.upper() or "16X2" in k.upper()}
    print(f"\n{fn[:150]}\n  {tot} instructions; " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)))
    if feats: print("  of note: " + ", ".join(f"{k} {v}" for k, v in sorted(feats.items())))
