"""Summary of one `ncu --set full` capture for profiles/: the headline metrics from the raw page plus the per-function and
per-line instruction shares from the source page.  Usage: python scripts/ncu_summary.py RAW.csv REP.ncu-rep "header line" > profiles/...txt"""
import csv, subprocess, sys

raw, rep, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.max"]
print(title)
for k in keep:
    if k in hdr:
        i = hdr.index(k)
        print("%-88s %s %s" % (k, vals[i], units[i]))
for i, h in enumerate(hdr):
    if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        print("%-88s %s %s" % (h, vals[i], units[i]))
print()
print("instruction / stall / shared-memory wavefront shares by function and by source line (scripts/ncu_lines.py):")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "40"])
