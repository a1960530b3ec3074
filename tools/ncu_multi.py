"""Summarise every kernel of an .ncu-rep (the counters the design is argued from + stall-sample shares).
usage: python tools/ncu_multi.py report.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__cycles_elapsed.max']
stall = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("## " + r[ix['Kernel Name']][:120])
    for h in want:
        if h in ix:
            print(f"{h:72s} {r[ix[h]]:>22s} {units[ix[h]]}")
    ss = collections.Counter()
    for h in stall:
        try:
            ss[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] += float(r[ix[h]].replace(',', '') or 0)
        except ValueError:
            pass
    tot = sum(ss.values())
    if tot:
        print("stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in ss.most_common(7)))
    print()
