#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (run where ncu is installed): python profiles/ncu_summary.py [--json out.json] rep [...]
--json: also write {"kernels": {short kernel name: {dram_bytes_read, dram_bytes_write, us, launches}}} averaged per launch (bench.py's roofline.traffic)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']

import json
import re
args = sys.argv[1:]
json_out = None
if args and args[0] == '--json':
    json_out, args = args[1], args[2:]
agg = {}


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


for rep in args:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:110])
        short = re.sub(r'^(void )?(elm::)?(\(anonymous namespace\)::)?', '', r[hdr.index('Kernel Name')]).split('<')[0].split('(')[0]
        a = agg.setdefault(short, {'dram_bytes_read': [], 'dram_bytes_write': [], 'us': [], 'launches': 0})
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        for key, dst in (('dram__bytes_read.sum', 'dram_bytes_read'), ('dram__bytes_write.sum', 'dram_bytes_write')):
            if key in hdr:
                a[dst].append(num(r[hdr.index(key)]) * scale.get(units[hdr.index(key)], 1.0))
        if 'gpu__time_duration.sum' in hdr:
            i = hdr.index('gpu__time_duration.sum')
            a['us'].append(num(r[i]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(units[i], 1.0))
        a['launches'] += 1
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'   {k:90s} {r[i]:>16s} {units[i]}')

if json_out:
    for a in agg.values():
        for k in ('dram_bytes_read', 'dram_bytes_write', 'us'):
            v = sorted(a[k])
            a[k + '_per_launch'] = a[k]
            a[k] = v[len(v) // 2] if v else 0.0  # median over the captured launches
    with open(json_out, 'w') as f:
        json.dump({'kernels': agg}, f, indent=1)
