"""Per-source-line view of an .ncu-rep (first kernel): instructions executed and stall samples by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# find header row
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
samp_col = next((h for h in hdr if h.startswith('Warp Stall Sampling (All')), None)
stall_cols = [h for h in hdr if h.startswith('stall_') and '(' not in h]
tot_s = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    n = int(float(r[ix['Instructions Executed']] or 0))
    s = int(float(r[ix[samp_col]] or 0)) if samp_col else 0
    st = {c[6:]: int(float(r[ix[c]] or 0)) for c in stall_cols}
    tot_s += s
    lines.append((s, n, r[ix['Source']], st, r[ix.get('Address', 0)] if 'Address' in ix else ''))
print("total samples", tot_s, "instructions", sum(l[1] for l in lines))
# cumulative by position: print SASS in order with samples, compressing cold regions
for i, (s, n, text, st, addr) in enumerate(lines):
    if s >= tot_s / 200:
        why = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"{i:5d} {100 * s / max(tot_s, 1):5.1f}%  x{n:<9d} {text[:90]:90s} {why}")
# instruction volume by contiguous SASS region (split at barriers / exec-count changes)
print("--- regions (start idx, sass lines, executed warp-instructions, samples)")
cur = None
acc = [0, 0, 0, 0]
for i, (s, n, text, st, addr) in enumerate(lines):
    key = n
    isbar = 'BAR.' in text or 'B2R' in text
    if cur is None or (key != cur and abs(key - cur) > 0.2 * max(cur, 1)) or isbar:
        if acc[1]:
            print(f"  idx {acc[0]:5d} len {acc[1]:5d} exec/line {cur:9d} instr {acc[2]:10d} samples {acc[3]:5d}")
        acc = [i, 0, 0, 0]
        cur = key
    acc[1] += 1
    acc[2] += n
    acc[3] += s
print(f"  idx {acc[0]:5d} len {acc[1]:5d} exec/line {cur:9d} instr {acc[2]:10d} samples {acc[3]:5d}")
