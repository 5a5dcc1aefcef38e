"""Turn an ncu launch list (csv) and an `ncu --set full` report into the markdown summaries kept under profiles/.
python tools/profile_summary.py <launches.csv> <report.ncu-rep> <out_prefix> <label>"""
import collections
import csv
import io
import subprocess
import sys


def launches(path, label):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
            for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    idx = [i for i, (k, v, u) in enumerate(rows) if 'adam_flat' in k]
    a, b = idx[3] + 1, idx[4] + 1                              # the 5th executed step (timed region)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, u in rows[a:b]:
        key = k.split('(')[0][:90]
        agg[key][0] += 1
        agg[key][1] += v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[u]
    tot = sum(v[1] for v in agg.values())
    out = ['# %s -- launch list of one training step (cubes config, eager launch order)' % label, '',
           'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file <csv> python bench.py --steps 3 '
           '--warmup 3 --no-e2e --no-cpu-baseline --no-graph` (one of the timed steps is shown).  ncu serialises the launches and runs them '
           'with cold caches: compare SHARES, not absolutes.  The timed bench replays the same launches as one CUDA graph.', '',
           'launches per step: %d, summed kernel time %.0f us' % (b - a, tot), '',
           '| kernel | launches/step | us/step | share |', '|---|---:|---:|---:|']
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        out.append('| `%s` | %d | %.1f | %.1f%% |' % (k.replace('|', '/'), c, t, 100 * t / tot))
    mine = sum(t for k, (c, t) in agg.items() if 'mcnn::' in k)
    out += ['', 'our kernels (`mcnn::`): %.1f%% of kernel time' % (100 * mine / tot)]
    return '\n'.join(out) + '\n'


def full(path, label):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    h = r[0]
    c = lambda n: h.index(n)
    out = ['# %s -- `ncu --set full` of the top kernels (one training step, cubes config, 64 x 750-edge meshes)' % label, '',
           'Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on -k '
           'regex:"pool_collapse|meshconv_.*_tc_kernel" -s 75 -c 25 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-graph`.',
           'Units: time us, DRAM Mbyte, dynamic shared memory Kbyte per block.  Launches in step order (forward, then backward).', '',
           '| # | kernel | time | grid | regs | dyn smem | DRAM read | DRAM write | L2 throughput % | tensor pipe active % | smem store bank conflicts / wavefronts |',
           '|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|']
    for i, row in enumerate(r[2:]):
        g = lambda n: row[c(n)] if n in h else '-'
        out.append('| %d | `%s` | %.1f | %s | %s | %.1f | %.2f | %.2f | %.1f | %.1f | %s / %s |' % (
            i, g('Kernel Name').split('(')[0], float(g('gpu__time_duration.sum')), g('launch__grid_size'), g('launch__registers_per_thread'),
            float(g('launch__shared_mem_per_block_dynamic')), float(g('dram__bytes_read.sum')), float(g('dram__bytes_write.sum')),
            float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')), float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')),
            g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'), g('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum')))
    return '\n'.join(out) + '\n'


if __name__ == '__main__':
    csv_path, rep, prefix, label = sys.argv[1:5]
    open(prefix + '_launches.md', 'w').write(launches(csv_path, label))
    open(prefix + '_ncu_top_kernels.md', 'w').write(full(rep, label))
