#!/usr/bin/env python
"""Stall-reason totals of an ncu source page, EXCLUDING instructions that every CTA executes equally often (the idle
spin of CTAs without work): keeps only instructions whose executed count is below a threshold.
usage: ncu_busy.py <report> <max_exec_count>"""
import csv, io, subprocess, sys
rep, cap = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for _, h in stalls}
n_s = n_e = 0
for r in data:
    e = int(r[iex])
    if e == 0 or e > cap: continue
    n_s += int(r[ismp]); n_e += e
    for i, h in stalls: tot[h] += int(r[i])
print("instructions kept: %d executions, %d samples" % (n_e, n_s))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-24s %6.1f %%" % (h, 100.0 * v / max(n_s, 1)))
