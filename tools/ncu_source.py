"""Summarises an `ncu --page source --csv` dump: per-opcode executed counts / stall samples and the hottest SASS lines.
   ncu_source.py file.csv [top_n]"""
import csv, re, sys
from collections import defaultdict
allrows = list(csv.reader(open(sys.argv[1])))
# the dump holds one block per kernel: ["Kernel Name", name], header, lines...; pick block KSEL (env, default last)
import os
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name']
k = int(os.environ.get('KSEL', len(starts) - 1))
rows = allrows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(allrows))]
print("kernel:", rows[0][1], f"(block {k} of {len(starts)})")
hdr = rows[1]
I = {h: i for i, h in enumerate(hdr)}
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ops = defaultdict(lambda: [0, 0, 0])
tot_inst = tot_samp = 0
lines = []
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stall_tot = defaultdict(int)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[I['Source']].strip()
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else '?'
    ex = int(r[I['Instructions Executed']] or 0); sm = int(r[I['# Samples']] or 0)
    ops[op][0] += ex; ops[op][1] += sm; ops[op][2] += 1
    tot_inst += ex; tot_samp += sm
    lines.append((sm, ex, src, r))
    for c in stall_cols: stall_tot[c] += int(r[I[c]] or 0)
print(f"total warp-instructions {tot_inst}, samples {tot_samp}")
print("opcode        executed    share   samples  share  static")
for op, (ex, sm, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{op:12s} {ex:10d}  {100*ex/max(tot_inst,1):5.1f}%  {sm:8d} {100*sm/max(tot_samp,1):5.1f}%  {n}")
print("stall reasons:", ", ".join(f"{c[6:]} {100*v/max(tot_samp,1):.1f}%" for c, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:8]))
print("hottest lines:")
for sm, ex, src, r in sorted(lines, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[I[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"  {sm:6d} samples  {ex:9d} exec  {src[:70]:70s} {st}")
