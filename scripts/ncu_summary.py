"""Summarise an `ncu --page raw --csv` dump: one block of selected metrics per profiled launch.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv > profiles/NAME.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:72s} {r[i]:>22s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                st.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_")[0]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("   warp stall reasons per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
