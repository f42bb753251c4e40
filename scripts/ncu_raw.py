"""Key raw metrics of every kernel in an ncu report:  python scripts/ncu_raw.py <report.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name', '')[:70], 'grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
    for k in KEYS:
        if k in d:
            print('   %-70s %s' % (k, d[k]))
    st = sorted(((float(d[k]), k.split('issue_stalled_')[1].split('_per_issue')[0]) for k in hdr
                 if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and d[k]), reverse=True)
    tot = sum(v for v, _ in st)
    print('   stalls (warps per issue-active cycle, total %.2f): ' % tot + ', '.join('%s %.2f' % (n, v) for v, n in st[:9]))
