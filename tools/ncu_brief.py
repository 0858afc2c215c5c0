"""brief of an `ncu --page raw --csv` dump: per kernel the duration, pipe / memory utilisation and the top stall reasons"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr:
            print('  %-85s %s' % (k, r[hdr.index(k)]))
    st = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
            try:
                st.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print('  stalls:', ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
