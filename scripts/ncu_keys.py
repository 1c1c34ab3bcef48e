#!/usr/bin/env python
"""Key counters of an .ncu-rep for the gather kernels: python scripts/ncu_keys.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_tex_mem_texture.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.max',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w} = {vals[i]} {units[i]}")
st = [(float(vals[i].replace(',', '')), h) for i, h in enumerate(hdr)
      if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and vals[i] not in ('', 'n/a')]
for v, h in sorted(st, reverse=True)[:6]:
    print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]} = {v:.2f}")
