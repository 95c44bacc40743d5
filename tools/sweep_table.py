"""Tabulates the output of tools/seg_fine_sweep.sh / tools/pieces_sweep.sh: sweep_table.py file..."""
import re, sys, collections
for f in sys.argv[1:]:
    tab = collections.OrderedDict(); keys = []
    for line in open(f):
        m = re.match(r'## \w+=(\d+)', line)
        if m: key = int(m.group(1)); keys.append(key); continue
        m = re.match(r'octave\s+(\d+)\s+(\S+)\s+(\S+)\s+([\d.]+) us', line)
        if m: tab.setdefault((int(m.group(1)), m.group(2), m.group(3)), {})[key] = float(m.group(4))
        m = re.match(r'sum of pyramid launches: ([\d.]+)', line)
        if m: tab.setdefault(('sum', '', ''), {})[key] = float(m.group(1))
    print(f); print('%-28s' % 'knob' + ''.join('%8d' % s for s in keys))
    for k, v in tab.items():
        print('%-28s' % str(k) + ''.join('%8.1f' % v.get(s, 0) for s in keys))
