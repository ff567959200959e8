"""Print the metrics we track from an `ncu --page raw --csv` export.  python scripts/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_imc_miss_per_warp_active.pct', 'smsp__warp_issue_stalled_drain_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct']

rows = list(csv.reader(open(sys.argv[1])))
H, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(H)}
for r in rows[2:]:
    print('-----', r[idx['Kernel Name']][:60], 'grid', r[idx.get('Grid Size', 0)], 'block', r[idx.get('Block Size', 0)])
    for w in WANT:
        if w in idx and r[idx[w]] not in ('', 'n/a'):
            print(f"  {w:72s} {r[idx[w]][:24]:>24s} {units[idx[w]]}")
