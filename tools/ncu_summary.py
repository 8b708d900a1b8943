"""Extract the judged metrics from an .ncu-rep (read here, on the CPU box) into a small text summary."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x', 'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']

for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f'== {rep}: {r[hdr.index("Kernel Name")][:90]}')
        for k in KEYS:
            if k in hdr:
                print(f'   {k:90s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
