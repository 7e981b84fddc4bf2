#!/usr/bin/env python
"""Turn an ncu report (scratch, gpurun_out/) into the committed text summary + profiles/traffic.json.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_rN_name.txt <cells per launch> "<command line>" [--update-traffic-json] """
import csv
import json
import os
import subprocess
import sys

rep, out, cells, cmdline = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__warps_eligible.avg.per_cycle_active']
lines, traffic = [], []
for r in rows[2:]:
    lines.append('---- ' + r[hdr.index('Kernel Name')])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"{w} = {r[i]} {units[i]}")
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'byte': 1, 'Kbyte': 1e3}
    rd = float(r[hdr.index('dram__bytes_read.sum')]) * scale[units[hdr.index('dram__bytes_read.sum')]]
    wr = float(r[hdr.index('dram__bytes_write.sum')]) * scale[units[hdr.index('dram__bytes_write.sum')]]
    dur = float(r[hdr.index('gpu__time_duration.sum')]) * {'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1}[units[hdr.index('gpu__time_duration.sum')]]
    traffic.append(rd + wr)
    lines.append(f"=> DRAM traffic per launch = {(rd + wr) / 1e9:.3f} GB = {(rd + wr) / cells:.1f} B per lattice update (algorithmic minimum 432 B); {(rd + wr) / dur / 1e12:.2f} TB/s under ncu")
head = f"command: {cmdline}\nreport : {rep} (scratch); selected raw metrics below.  Numbers under ncu are cold-cache, serialised replays.\n"
open(out, 'w').write(head + '\n'.join(lines) + '\n')
print('\n'.join(lines))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 5 and sys.argv[5] == "--update-traffic-json":  # only for the capture of the headline kernel: bench.py reads profiles/traffic.json
    json.dump({"dram_bytes_per_update": sum(traffic) / len(traffic) / cells, "measured_at": f"{cells} cells per launch, D3Q27 cumulant fp64 A-A, mean of one even and one odd launch",
           "source": os.path.relpath(out, root), "streaming": "AA"}, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
