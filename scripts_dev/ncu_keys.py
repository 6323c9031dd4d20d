"""Key metrics per kernel from an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
K = [('gpu__time_duration.sum', 'time_us'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
     ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'),
     ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_pipe%'),
     ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('launch__registers_per_thread', 'regs'),
     ('smsp__inst_executed.sum', 'inst'),
     ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'),
     ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
     ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'st_notsel'),
     ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'),
     ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'st_mio'),
     ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
     ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma%')]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print(name[:70])
    out = []
    for k, lab in K:
        c = [i for i, h in enumerate(hdr) if h == k or h.endswith('.' + k)]
        if c: out.append(f"{lab}={r[c[0]]}")
    print('   ', '  '.join(out))
