"""Extracts the counters bench.py's K1 roofline needs from an `ncu --set full` raw-page CSV
(one launch of the default c2 walk: 4096 chains x 10^4 steps, thin 1) into
profiles/k1_counters.json.
usage: python scripts/ncu_counters.py profiles/<tag>_ncu_full_k1.csv [chains steps thin]"""
import csv, json, os, sys
path = sys.argv[1]
chains, steps, thin = (int(v) for v in (sys.argv[2:5] or (4096, 10000, 1)))
rows = list(csv.reader(open(path)))
hdr, units, r = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
cyc = get("sm__cycles_elapsed.max")
nsm = int(get("launch__sm_count")) if "launch__sm_count" in hdr else 148
fp64_pct = get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
act = get("sm__cycles_active.avg") if "sm__cycles_active.avg" in hdr else cyc
out = {
    "source": os.path.basename(path) + " (ncu --set full --clock-control none, one launch)",
    "kernel": r[hdr.index("Kernel Name")],
    "chains": chains, "steps": steps, "thin": thin,
    "grid": int(get("launch__grid_size")), "block": int(get("launch__block_size")),
    "registers": int(get("launch__registers_per_thread")),
    "waves_per_sm": get("launch__waves_per_multiprocessor"),
    "duration_us_under_ncu": get("gpu__time_duration.sum"),
    "warp_inst": get("smsp__inst_executed.sum"),
    # FP64 pipe: pct of peak (1 warp instruction per 2 cycles per sub-partition) x active cycles
    "fp64_warp_inst": fp64_pct / 100.0 * act * nsm * 4 / 2.0,
    "fp64_pipe_pct_active": fp64_pct,
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
    "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
    "smem_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "smem_bank_conflicts": get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(path)), "k1_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
