"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --clock-control none --csv`) -> per-kernel table.  python scripts/launch_list.py launches.csv [title]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1], errors='replace') if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else (v * 1e3 if r[ui] in ('ms', 'msecond') else v)
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('# %s\n' % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print('| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f | %.1f | %.1f %% |' % (k, c, t / c, t, 100 * t / tot))
