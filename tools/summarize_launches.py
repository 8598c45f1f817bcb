"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline='') as fh:
    lines = [l for l in fh if not l.startswith('==')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
total = 0.0
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    name = name.replace('hamt::', '')
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1e3 if unit in ('nsecond', 'ns') else (v if unit in ('usecond', 'us') else v * 1e3)
    a = agg.setdefault(name[:70], [0, 0.0])
    a[0] += 1; a[1] += us; total += us
print(f"total {total/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
print(f"{'kernel':72s} {'n':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:6d} {us/1e3:9.3f} {us/total:7.1%} {us/n:9.1f}")
