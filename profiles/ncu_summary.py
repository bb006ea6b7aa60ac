"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv > X.csv) into the handful
of numbers DESIGN.md / bench.py quote.  Usage: python profiles/ncu_summary.py X.csv [pattern...]"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_sectors_op_red.sum", "lts__t_requests_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "sm__sass_inst_executed_op_global_red.sum",
    "smsp__sass_inst_executed_op_global_red.sum",
]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in hdr:
            short = k.split(".", 2)[-1] if k.split(".")[0].isupper() and "Triage" in k else k
            if short in KEYS or k in KEYS or any(p in k for p in extra):
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_warp_active.pct"):
                try:
                    stalls.append((float(d[k].replace(",", "")), k))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print("  stall %-64s %.1f %%" % (k.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))


if __name__ == "__main__":
    main()
