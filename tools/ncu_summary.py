"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel. Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_op_red.sum"]
for r in rows[2:]:
    print("=" * 100)
    print(r[H.index("Kernel Name")][:160])
    for k in keys:
        if k in H:
            print(f"  {k:75s} {r[H.index(k)]} {rows[1][H.index(k)]}")
    tot = 0
    st = []
    for i, h in enumerate(H):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                v = float(r[i]); tot += v; st.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  stall samples: " + ", ".join(f"{n} {100*v/tot:.1f}%" for v, n in st[:7]))
