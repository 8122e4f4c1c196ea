"""Sorted device-side timeline of the LAST traced host-buffer call in a VO_TRACE=1 stderr log. Usage: trace_fmt.py file"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
idx = [i for i, l in enumerate(lines) if 'uploads end' in l]
ev = []
for l in lines[idx[-1]:]:
    m = re.match(r'\[vo trace\]\s+([\d.]+) ms\s+(.*?)(\(enqueued.*)?$', l)
    if m: ev.append((float(m.group(1)), m.group(2).strip()))
    else: print(l)
row = {}
for t, n in sorted(ev):
    k = n.rsplit(' ', 1)
    row.setdefault(k[-1], []).append(f"{k[0]}={t:.3f}")
for b in sorted(row, key=lambda x: int(x) if x.isdigit() else -1): print(b, ' '.join(row[b]))
