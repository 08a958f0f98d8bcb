#!/usr/bin/env python
"""Hottest SASS lines of one kernel in an ncu report (stall samples per instruction, in program order):
    python tools/ncu_hot.py rep.ncu-rep kernel-regex [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
data = []
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= max(si, src, ie):
        break           # next kernel instance
    data.append((int(r[si] or 0), i, r[src].strip(), r[ie]))
tot = sum(d[0] for d in data) or 1
print("kernel %s: %d SASS lines, %d samples" % (pat, len(data), tot))
for s, i, t, e in sorted(sorted(data, reverse=True)[:n], key=lambda x: x[1]):
    print("%6d %5.1f%%  @%4d exec %9s  %s" % (s, 100.0 * s / tot, i, e, t[:100]))
