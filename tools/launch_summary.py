"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last step of a 4-step bench run."""
import collections
import csv
import sys


def main(path, steps=4, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
            for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    per = len(rows) // steps
    last = rows[-per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, u in last:
        key = k.split('(')[0][:100]
        agg[key][0] += 1
        agg[key][1] += v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[u]
    tot = sum(v[1] for v in agg.values())
    print('launches per step: %d, kernel time %.0f us' % (per, tot))
    print('| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('| `%s` | %d | %.1f | %.1f%% |' % (k.replace('|', '/'), c, t, 100 * t / tot))
    mine = sum(t for k, (c, t) in agg.items() if 'mcnn::' in k)
    print('\nour kernels: %.1f%% of kernel time' % (100 * mine / tot))
    return last


if __name__ == '__main__':
    main(sys.argv[1])
