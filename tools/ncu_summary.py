"""`ncu -i X.ncu-rep --page raw --csv` -> a small transposed table of the metrics the design notes quote."""
import csv, sys
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warp_latency_issue_stalled_membar.ratio', 'smsp__average_warp_latency_issue_stalled_wait.ratio',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_branch_resolving.ratio',
        'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_no_instruction.ratio',
        'smsp__average_warp_latency_issue_stalled_sleeping.ratio', 'smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio',
        'lts__t_bytes.sum', 'lts__t_sectors_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
kcol = names.index('Kernel Name')
print('metric,unit,' + ','.join('"%s"' % r[kcol][:70] for r in data))
for m in KEEP:
    cols = [i for i, n in enumerate(names) if n == m or n.endswith('.' + m)]
    if not cols:
        continue
    c = cols[0]
    print('%s,%s,%s' % (m, units[c], ','.join(r[c].replace(',', '') for r in data)))
