#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics we steer by."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
keys = [
 ("Kernel Name", "kernel"),
 ("gpu__time_duration.sum", "ns"),
 ("launch__registers_per_thread", "regs"),
 ("launch__occupancy_limit_registers","occ_regs"),
 ("launch__occupancy_limit_shared_mem","occ_smem"),
 ("sm__warps_active.avg.per_cycle_active", "warps/SM"),
 ("dram__bytes_read.sum", "dram_rd"),
 ("dram__bytes_write.sum", "dram_wr"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
 ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wave%"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_waves"),
 ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_waves"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
 ("smsp__inst_executed.sum", "inst"),
 ("smsp__inst_executed_op_local_st.sum","local_st"),
 ("smsp__inst_executed_op_local_ld.sum","local_ld"),
 ("lts__t_sector_hit_rate.pct","l2hit%"),
]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----", d.get("Kernel Name","")[:100], d.get("Grid Size",""), d.get("Block Size",""))
    for k, nm in keys[1:]:
        if k in d: print(f"  {nm:12s} {d[k]}")
    st = sorted(((float(d[s]) if d[s] not in ('','nan','-nan') else 0.0, s.split('stalled_')[1].split('_per_issue')[0]) for s in stall), reverse=True)[:7]
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st))
