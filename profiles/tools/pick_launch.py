"""From an ncu gpu__time_duration launch list: index (among launches matching a kernel-name regex) of the longest launch."""
import csv, re, sys
path, pat = sys.argv[1], re.compile(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
k = 0
best = (-1.0, 0)
for row in r:
    name = row.get('Kernel Name', '')
    if not pat.search(name) or row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    if v > best[0]:
        best = (v, k)
    k += 1
print(best[1])
