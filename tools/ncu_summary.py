"""Summarise an .ncu-rep (first kernel) into the handful of counters the design is argued from.
usage: python tools/ncu_summary.py report.ncu-rep [tiles]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {vals[i]:>22s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop, stall, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = r[ix['Source']].split()
    op = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
    n = int(float(r[ix['Instructions Executed']] or 0))
    byop[op] += n
    tot += n
    for k in ix:
        if k.startswith('stall_') and '(' not in k:
            stall[k] += int(float(r[ix[k]] or 0))
print("warp instructions:", tot, (f"= {tot / tiles:.0f} per tile" if tiles else ""))
print("top opcodes:", ", ".join(f"{o} {n / (tiles or 1):.1f}" for o, n in byop.most_common(14)))
ts = sum(stall.values()) or 1
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / ts:.0f}%" for k, v in stall.most_common(7)))
