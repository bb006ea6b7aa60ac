"""Turn ncu captures into profiles/r2/ncu_stamps.json: the profiler-only figures bench.py quotes (DRAM bytes per
launch of the ray-cast, FP64-pipe share of the ICP kernel), each stamped with the SHA-1 of the kernel sources it was
measured on.  bench.py re-hashes those sources at run time and reports null when they have changed since.

    ncu -i gpurun_out/grid16k.ncu-rep --page raw --csv > gpurun_out/grid16k_raw.csv      (same for icp360 / icp1080)
    python profiles/scripts/stamp_ncu.py gpurun_out/grid16k_raw.csv gpurun_out/icp360_raw.csv gpurun_out/icp1080_raw.csv
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200/csrc/"
SOURCES = {"grid_raycast": [PKG + "b2s_grid.cu", PKG + "b2s_common.cuh"],
           "icp": [PKG + "b2s_icp_kernel.cuh", PKG + "b2s_common.cuh"]}


def sha(files):
    h = hashlib.sha1()
    for rel in files:
        with open(os.path.join(ROOT, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def row(path, pattern):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pattern in d.get("Kernel Name", ""):
            return d
    raise SystemExit("no %s launch in %s" % (pattern, path))


def num(d, key):
    return float(d[key].replace(",", ""))


def duration_us(path, d):
    rows = list(csv.reader(open(path)))
    unit = dict(zip(rows[0], rows[1]))["gpu__time_duration.sum"]
    return num(d, "gpu__time_duration.sum") * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3, "s": 1e6}[unit]


def main():
    grid_csv, icp360_csv, icp1080_csv = sys.argv[1:4]
    units = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    out = {}
    rows = list(csv.reader(open(grid_csv)))
    hdr, unit_row = rows[0], rows[1]
    d = row(grid_csv, "grid_raycast")
    u = dict(zip(hdr, unit_row))
    dram = sum(num(d, k) * units[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    grid = [int(v) for v in d["Grid Size"].strip("()").split(",")]
    out["grid_raycast"] = {
        "kernel": d["Kernel Name"].split("(")[0], "scans": grid[0] * 256 // 1080, "dram_bytes_per_launch": int(dram),
        "warp_instructions": int(num(d, "smsp__inst_executed.sum")),
        "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "duration_us_under_ncu": duration_us(grid_csv, d),
        "sources": SOURCES["grid_raycast"], "sha1": sha(SOURCES["grid_raycast"])}
    for key, path in (("icp_360", icp360_csv), ("icp_1080", icp1080_csv)):
        d = row(path, "icp_batch_kernel")
        out[key] = {
            "kernel": d["Kernel Name"].split("(")[0],
            "fp64_pipe_active_pct": num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "fp64_instruction_issue_pct": num(d, "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"),
            "warp_instructions": int(num(d, "smsp__inst_executed.sum")),
            "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "achieved_occupancy_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": int(num(d, "launch__registers_per_thread")),
            "duration_us_under_ncu": duration_us(path, d),
            "note": "executed-instruction figures of one launch under ncu --set full (serialised, cold cache); the share "
                    "of the FP64 pipe is what the kernel is quoted against, not a brute-force-equivalent rate",
            "sources": SOURCES["icp"], "sha1": sha(SOURCES["icp"])}
    path = os.path.join(ROOT, "profiles", "r2", "ncu_stamps.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
