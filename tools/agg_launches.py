"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of the step per kernel."""
import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    agg[name][0] += 1; agg[name][1] += v; tot += v
print("total %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%9.1f us %5.1f%% n=%4d avg %7.1f  %s" % (t, 100 * t / tot, n, t / n, k[:100]))
