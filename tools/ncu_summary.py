"""Summarise an .ncu-rep (raw page) into a small text table for profiles/ (run here, no GPU needed)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("==", d["Kernel Name"][:110])
        for k in KEYS:
            if k in d: print(f"   {k:75s} {d[k]:>16s} {u[k]}")
        st = sorted(((float(d[k] or 0), k[len(STALL):-len('_per_issue_active.ratio')]) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)
        print("   stalls (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
if __name__ == "__main__":
    main(sys.argv[1])
