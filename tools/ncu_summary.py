"""Developer tool: condense an `ncu --set full` report into the handful of numbers quoted in DESIGN.md / profiles/README.md.
Usage: ncu_summary.py <report.ncu-rep> [kernel substring]  -> text on stdout (one block per captured launch)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "inst_executed", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    stall = [c for c in h if "issue_stalled" in c and "per_issue_active" in c and "not_issued" not in c]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        if sub not in name:
            continue
        print(f"== {name[:110]}")
        for w in WANT:
            if w in h:
                print(f"{w:78s} {units[h.index(w)]:16s} {r[h.index(w)]}")
        st = sorted([(float(r[h.index(c)].replace(',', '') or 0), c) for c in stall], reverse=True)[:7]
        for v, c in st:
            print(f"stall per issue: {c.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):60s} {v:.2f}")
        print()


if __name__ == "__main__":
    main()
