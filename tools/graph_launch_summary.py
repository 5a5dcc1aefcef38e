"""Summarise an `ncu --metrics gpu__time_duration.sum --graph-profiling node --csv` log of bench.py: the kernel nodes of ONE
replay of the step graph (the launches between the last two L2-flush fills of the timed region) as a markdown table.
python tools/graph_launch_summary.py <csv> [title]"""
import collections
import csv
import sys


def main(path, title='launch list'):
    lines = [l for l in open(path) if not l.startswith('==')]
    unit = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
    rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')) * unit[r['Metric Unit']])
            for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    flush = [i for i, (k, v) in enumerate(rows) if 'FillFunctor<unsigned char>' in k and v > 30]
    a, b = flush[0], flush[1]                       # the first timed step: a graph replay between two flushes
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows[a + 1:b]:
        key = k.split('(')[0][:100]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print('## %s\n\nlaunches per step: %d, summed kernel time %.0f us\n' % (title, b - a - 1, tot))
    print('| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('| `%s` | %d | %.1f | %.1f%% |' % (k.replace('|', '/'), c, t, 100 * t / tot))
    mine = sum(t for k, (c, t) in agg.items() if 'mcnn::' in k or 'seq::pool_collapse' in k)
    print('\nour kernels (`mcnn::`, incl. `mcnn::seq::`): %.1f%% of kernel time' % (100 * mine / tot))


if __name__ == '__main__':
    main(*sys.argv[1:3])
