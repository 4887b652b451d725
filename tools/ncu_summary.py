"""Print the key metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_xu.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
idx = [(w, h.index(w)) for w in want if w in h]
units = rows[1]
for r in rows[2:]:
    print('---')
    for w, i in idx:
        print('  %-72s %s %s' % (w, r[i][:90], units[i]))
