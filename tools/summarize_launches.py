"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.

usage: python tools/summarize_launches.py launches.csv [first_id [last_id]] > summary.md
"""
import csv, re, sys, collections

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    i = int(r['ID'])
    if lo <= i < hi:
        rows.append((i, r['Kernel Name'], float(r['Metric Value'].replace(',', ''))))
tot = collections.defaultdict(float)
cnt = collections.Counter()
for _, name, ns in rows:
    short = re.sub(r'\(.*$', '', name)
    short = re.sub(r'<unnamed>', 'anon', short)
    if len(short) > 110:
        short = short[:110] + '...'
    tot[short] += ns
    cnt[short] += 1
total = sum(tot.values())
print('| kernel | launches | total ms | share |')
print('|---|---:|---:|---:|')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / total < 0.0005:
        continue
    print('| `%s` | %d | %.3f | %.1f%% |' % (k, cnt[k], v * 1e-6, 100 * v / total))
print('| **total** | %d | %.3f | 100%% |' % (len(rows), total * 1e-6))
mine = sum(v for k, v in tot.items() if 'mpdo::' in k)
print('\nlibrary kernels (`mpdo::`): %.1f%% of the listed device time; ids %d..%d' % (100 * mine / total, rows[0][0], rows[-1][0]))
