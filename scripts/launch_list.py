"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv) into a markdown table: per kernel launches / mean / total / share."""
import collections
import csv
import sys


def main(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    start = rows.index(hdr)
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e3          # ns -> us
    tot = sum(t for _, t in agg.values())
    with open(out, 'w') as fh:
        fh.write(f'# {title}\n\nsource: `{path}`; per-launch times are cold-cache and serialised by ncu - use the SHARES.\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if t / tot < 0.001:
                continue
            name = k.split('(')[0].replace('void ', '')[:70]
            fh.write(f'| `{name}` | {n} | {t / n:.1f} | {t:.1f} | {100 * t / tot:.1f} % |\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else 'launch list')
