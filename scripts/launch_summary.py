"""Aggregate an `ncu --metrics ... --csv` launch list by kernel name (per-launch averages and shares)."""
import collections
import csv
import sys


def main(path, n_last=None):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith('==')))
    by_id = collections.OrderedDict()
    for r in rows:
        d = by_id.setdefault(r['ID'], {'k': r['Kernel Name'].split('(')[0], 'g': r['Grid Size']})
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    items = list(by_id.values())
    if n_last:
        items = items[-n_last:]
    agg = collections.OrderedDict()
    for d in items:
        a = agg.setdefault(d['k'], collections.Counter())
        a['n'] += 1
        for k, v in d.items():
            if k not in ('k', 'g'):
                a[k] += v
    total = sum(a['gpu__time_duration.sum'] for a in agg.values())
    print('%d launches, total %.1f us' % (len(items), total / 1e3))
    for k, a in agg.items():
        line = '%-62s n=%3d  %8.1f us  %5.1f%%' % (k[-62:], a['n'], a['gpu__time_duration.sum'] / 1e3,
                                                  100 * a['gpu__time_duration.sum'] / total)
        for m in a:
            if m in ('n', 'gpu__time_duration.sum'):
                continue
            if 'pct' in m:
                line += '  %s=%.1f' % (m.split('.')[0].split('__')[-1][:18], a[m] / a['n'])
            elif 'bytes' in m:
                line += '  %s=%.0fMB' % (m.split('.')[0].split('__')[-1][:12], a[m] / 1e6)
            elif 'sectors' in m:
                line += '  %s=%.0fMB' % (m.split('.')[0].split('__')[-1][:18], a[m] * 32 / 1e6)
            else:
                line += '  %s=%.3g' % (m.split('.')[0].split('__')[-1][:14], a[m])
        print(line)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
