"""Prints selected metrics of every kernel in an ncu raw-page CSV.
usage: python scripts/ncu_metrics.py file.csv [substring ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = sys.argv[2:] or ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
                        "launch__registers_per_thread", "launch__waves_per_multiprocessor",
                        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64",
                        "smsp__issue_active.avg.pct", "sm__inst_issued.avg.pct",
                        "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "sm__pipe_fp64_cycles_active", "smsp__inst_executed_pipe_fp64",
                        "l1tex__data_bank_conflicts_pipe_lsu", "smsp__warps_eligible.avg.per_cycle",
                        "sm__warps_active.avg.pct", "launch__shared_mem"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90])
    for i, h in enumerate(hdr):
        if any(w in h for w in want):
            print("   %-90s %s %s" % (h, r[i], rows[1][i]))
