#!/usr/bin/env python
"""Summary of the per-task cycle counts a -DLFB_DP_PROF build of k_dp prints (stdin): per class, mean and max of each phase."""
import re, sys, collections
rows = collections.defaultdict(list)
for line in sys.stdin:
    m = re.match(r"dp_task G=(\d+) R=(\d+) blk \d+ nmax (\d+): setup (\d+) pre (\d+) tilt (\d+) main (\d+) \(par (\d+) loop (\d+) chk (\d+)\) tails (\d+) total (\d+) start \d+ parwait (\d+) pareval (\d+)", line)
    if m:
        v = list(map(int, m.groups()))
        rows[(v[0], v[1])].append(v[2:])
names = ["nmax", "setup", "pre", "tilt", "main", "par", "loop", "chk", "tails", "total", "parwait", "pareval"]
for k in sorted(rows):
    r = rows[k]
    n = len(r)
    print("G=%d R=%d: %d tasks" % (k[0], k[1], n))
    print("   mean " + "  ".join("%s %d" % (nm, sum(x[i] for x in r) / n) for i, nm in enumerate(names)))
    print("   max  " + "  ".join("%s %d" % (nm, max(x[i] for x in r)) for i, nm in enumerate(names)))
