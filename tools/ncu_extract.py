#!/usr/bin/env python
"""Key `ncu --set full` metrics of every kernel in an .ncu-rep -> JSON (what profiles/*.json holds).
usage: ncu_extract.py report.ncu-rep out.json"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, r):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled"):
                try:
                    d.setdefault("stall", {})[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(float(v.replace(",", "")), 3)
                except ValueError:
                    pass
            if h in KEYS:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if h.startswith("dram__bytes") or h.startswith("lts__t_bytes"):
                    d[h + " [bytes]"] = x * SCALE.get(u, 1)
                elif h == "gpu__time_duration.sum":
                    d[h + " [us]"] = x * SCALE.get(u, 1)
                else:
                    d[h] = x
        res.append(d)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for d in res:
        if d.get("stall"):
            top = sorted(d["stall"].items(), key=lambda kv: -kv[1])[:6]
            print("   stalls (warps per issue):", ", ".join("%s %.2f" % kv for kv in top))
        print("%-60s %8.1f us  dram %.1f MB" % (d["kernel"][:60], d.get("gpu__time_duration.sum [us]", 0),
                                              (d.get("dram__bytes_read.sum [bytes]", 0) + d.get("dram__bytes_write.sum [bytes]", 0)) / 1e6))


if __name__ == "__main__":
    main()
