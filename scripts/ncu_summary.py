"""Summarise ncu outputs: launch list csv -> per-kernel table; .ncu-rep raw page -> key metrics."""
import csv, collections, subprocess, sys

def launch_table(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, tot = {}, 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(row['Kernel Name'], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    out = ['| launches | total us | share | avg us | kernel |', '|---|---|---|---|---|']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('| %d | %.1f | %.1f%% | %.1f | `%s` |' % (n, t, 100 * t / tot, t / n, k[:100]))
    out.append('| %d | %.1f | 100%% | | total |' % (sum(a[0] for a in agg.values()), tot))
    return '\n'.join(out)

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']

def rep_table(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append('\n### `%s`\n' % d.get('Kernel Name', '?')[:110])
        out.append('| metric | value | unit |\n|---|---|---|')
        for k in KEYS:
            if k in d:
                out.append('| %s | %s | %s |' % (k, d[k], units[hdr.index(k)]))
    return '\n'.join(out)

if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(launch_table(p) if p.endswith('.csv') else rep_table(p))
