"""Key metrics from an .ncu-rep (per kernel launch).  usage: python tools/ncu_raw.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'launch__shared_mem_per_block_dynamic', 'sm__ctas_launched.sum']
import re

PIPES = re.compile(r"^(sm__inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active|sm__pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active)$")
STALL = re.compile(r"^smsp__average_warps?_(latency_)?issue_stalled_(\w+?)(_per_warp_active\.pct|\.ratio|_per_warp_active\.ratio)$")


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for r in rows[2:]:
    print('----')
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(' ', w, '=', r[i][:120], units[i])
    # busiest execution pipes and the top warp-stall reasons (whatever this ncu version calls them)
    pipes = sorted(((fnum(r[i]), h) for i, h in enumerate(hdr) if PIPES.match(h)), reverse=True)[:8]
    for v, h in pipes:
        print('  pipe', h, '=', round(v, 2), '%')
    stalls = sorted(((fnum(r[i]), h) for i, h in enumerate(hdr) if STALL.match(h)), reverse=True)[:8]
    for v, h in stalls:
        print('  stall', h, '=', round(v, 3))
