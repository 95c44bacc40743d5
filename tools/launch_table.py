"""Prints the ncu `--metrics gpu__time_duration.sum --csv` launch list as a compact table."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, gs = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
tot, agg = 0.0, {}
for n, r in enumerate(rows[hi + 1:]):
    name = re.sub(r'\(.*', '', r[kn]).replace('void ', '')
    v = float(r[mv].replace(',', '')) / 1000
    if n < limit:
        print(f"{name[:40]:40s} {r[gs]:16s} {v:9.1f} us")
    tot += v
    agg[name] = agg.get(name, 0) + v
print("---- per kernel (us, share)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k[:40]:40s} {v:10.1f} {100*v/tot:6.1f}%")
print(f"total {tot:.1f} us")
