"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio' ]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ''
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if pat not in name:
        continue
    print('==', name[:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'   {w:90s} {r[i]:>16s} {units[i]}')
    if len(sys.argv) > 3:
        for i, h in enumerate(hdr):
            if sys.argv[3] in h:
                print(f'   {h:90s} {r[i]:>16s} {units[i]}')
