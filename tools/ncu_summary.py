"""Selected metrics of every kernel in an ncu report as JSON (the format of profiles/*_ncu_full_summary.json).
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > summary.json"""
import csv
import json
import subprocess
import sys

KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio')
out = []
for rep in sys.argv[1:]:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {'report': rep.split('/')[-1], 'Kernel Name': r[hdr.index('Kernel Name')]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i] + ' ' + units[i]).strip()
        out.append(d)
json.dump(out, sys.stdout, indent=1)
