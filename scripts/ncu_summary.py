"""Summarise an `ncu --set full` report (read here, no GPU needed): per-kernel headline metrics and the
warp-stall samples of the source page aggregated by stall reason and by SASS opcode.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
       'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
       'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main(path):
    rows = list(csv.reader(io.StringIO(ncu(['-i', path, '--page', 'raw', '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [d[hdr.index('Kernel Name')].split('(DevProblem')[0] for d in data]
    print('== raw page: %d captured launches' % len(data))
    for n, nm in enumerate(names):
        print('  [%d] %s' % (n, nm))
    for k in RAW:
        if k in hdr:
            i = hdr.index(k)
            print('%-70s %-14s %s' % (k, units[i], '  '.join('%12.6g' % float(d[i]) for d in data)))
    stall = [(i, k) for i, k in enumerate(hdr) if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio')
             and 'not_issued' not in k]
    print('-- warps stalled per issue-active cycle, by reason (>= 0.1 somewhere)')
    for i, k in stall:
        v = [float(d[i]) for d in data]
        if max(v) >= 0.1:
            print('%-70s %s' % (k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''),
                                '  '.join('%12.2f' % x for x in v)))
    # ---- source page
    rows = list(csv.reader(io.StringIO(ncu(['-i', path, '--page', 'source', '--csv']))))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1].split('(DevProblem')[0], 'hdr': None, 'rows': []}
            kern.append(cur)
        elif cur is not None:
            if cur['hdr'] is None:
                cur['hdr'] = r
            else:
                cur['rows'].append(r)
    for k in kern:
        h = k['hdr']
        i_s, i_a, i_e = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
        sc = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
        tab = collections.defaultdict(collections.Counter)
        ex, static = collections.Counter(), collections.Counter()
        for r in k['rows']:
            parts = r[i_s].split()
            op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
            ex[op] += int(r[i_e] or 0)
            static[op] += 1
            for i, c in sc:
                v = int(r[i] or 0)
                if v:
                    tab[op][c[6:]] += v
                    tab['ALL'][c[6:]] += v
        tot = sum(tab['ALL'].values())
        print('\n== source page: %s' % k['name'])
        print('   %d SASS instructions, %d warp-instructions executed, %d stall samples'
              % (len(k['rows']), sum(ex.values()), tot))
        print('   by reason: ' + '  '.join('%s=%.1f%%' % (c, 100.0 * v / tot) for c, v in tab['ALL'].most_common(8)))
        ops = sorted(static, key=lambda o: -sum(tab[o].values()))[:14]
        for op in ops:
            t = sum(tab[op].values())
            print('   %-8s samples %5.1f%%  static %5d  executed %10d   %s'
                  % (op, 100.0 * t / tot, static[op], ex[op], ' '.join('%s=%d' % cv for cv in tab[op].most_common(3))))


if __name__ == '__main__':
    main(sys.argv[1])
