"""profiling helper: per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)
usage: python tools/launch_summary.py <launches.csv>"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0].replace('void ', '')
    t = float(r[14]) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[13], 1e-6)
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
s = sum(v[1] for v in tot.values())
print(f'{"kernel":44s} {"launches":>8s} {"total ms":>10s} {"mean ms":>9s} {"share":>7s}')
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:44s} {n:8d} {t:10.3f} {t/n:9.4f} {100*t/s:6.2f}%')
print(f'{"total":44s} {len(rows):8d} {s:10.3f}')
