#!/bin/bash
# summary metrics of one launch per case:  bash s2_probe_ncu.sh out.txt case...
out=$1; shift
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,l1tex__data_bank_conflicts_pipe_lsu.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for c in "$@"; do
  echo "== $c" >> $out
  python s2_probe.py $c 2 >> $out 2>&1 || continue
  ncu --metrics $M --clock-control none -k regex:'stencil|pool|sep2d|convc|patch|wgrad_rows|strided_vjp' -s 4 -c 1 --csv python s2_probe.py $c 2 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
for r in rows[1:]:
    print('  %-90s %s %s' % (r[-3][:90], r[-1], r[-2]) if r[-3]!='launch__block_size' else '  kernel %s' % r[4][:100])
" >> $out
done
