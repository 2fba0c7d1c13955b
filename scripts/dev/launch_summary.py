"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last full frame, grouped by kernel instantiation."""
import csv, re, sys, collections
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
allr = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    allr.append((row['Kernel Name'], v, row.get('Grid Size')))
names = [a[0] + str(a[2]) for a in allr]
L = None
for cand in range(150, 500):
    if len(names) >= 2 * cand and names[-cand:] == names[-2 * cand:-cand]:
        L = cand
        break
if L is None:
    print('no repeating frame found in', len(allr), 'launches'); sys.exit(1)
fr = allr[-L:]
print('launches per frame %d, sum of durations %.1f us' % (L, sum(v for _, v, _ in fr)))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, g in fr:
    m = re.search(r'([a-z_0-9]+_kernel(<[^>]*>)?)', n)
    key = (m.group(1) if m else n)[:80]
    agg[key][0] += 1
    agg[key][1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print('%-82s %3d %8.1f us  %6.1f avg' % (k, c, t, t / c))
