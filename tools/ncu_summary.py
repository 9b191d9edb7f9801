#!/usr/bin/env python
"""Summarise an .ncu-rep (first captured kernel) into the handful of numbers DESIGN.md / bench.py quote.
usage: tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_write.sum.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second', 'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
out = []
for r in rows[2:]:
    out.append("== " + " | ".join(r[hdr.index(k)] for k in ('Kernel Name', 'Grid Size', 'Block Size') if k in hdr))
    for i, h in enumerate(hdr):
        if h in KEEP[3:]:
            out.append(f"{h:92s} {r[i]} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                st.append((float(r[i]), h))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    for v, h in sorted(st, reverse=True)[:8]:
        out.append(f"stall {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {100 * v / tot:5.1f} %")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("ncu --set full --clock-control none (B200); " + rep + "\n" + txt + "\n")
