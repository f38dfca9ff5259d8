#!/usr/bin/env python
"""Per-source-line executed instructions and stall samples of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_lines.py report.ncu-rep kernel-substring [top]"""
import csv, collections, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
SORTK = 0 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, '', collections.Counter()])
stall_cols = {}
kstall = collections.Counter()
fname = cur = None
active = False
iE = iS = None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        active = kname in r[1]; continue
    if r[0] == 'Line No':
        iE = r.index('Instructions Executed'); iS = r.index('# Samples')
        stall_cols = {i: h[6:] for i, h in enumerate(r) if h.startswith('stall_') and '(' not in h}; continue
    if not active:
        continue
    if r[0] != '':
        cur = (fname, int(r[0])); agg[cur][2] = r[1].strip()[:100]; continue
    try:
        agg[cur][0] += int(r[iE]); agg[cur][1] += int(r[iS])
        for i, nm in stall_cols.items():
            if r[i] not in ('', '0'):
                agg[cur][3][nm] += int(r[i]); kstall[nm] += int(r[i])
    except Exception:
        pass
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("kernel~%s: %d warp instructions, %d samples" % (kname, tot, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][SORTK])[:top]:
    why = " ".join("%s:%d" % (a, b) for a, b in v[3].most_common(3))
    print("%-18s:%4d %5.1f%% inst %5.1f%% samp  [%s]  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ts, why, v[2][:70]))
print("stall samples of the kernel:", ", ".join("%s %d" % kv for kv in kstall.most_common(8)))
