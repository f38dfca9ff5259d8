"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean us, share of the total."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = name.replace("lfb::", "").replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        rows.append((name, us))
    agg = defaultdict(list)
    for n, us in rows:
        agg[n].append(us)
    tot = sum(us for _, us in rows)
    print("%-40s %8s %10s %8s" % ("kernel", "launches", "mean us", "share"))
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-40s %8d %10.1f %7.1f%%" % (n[:40], len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
