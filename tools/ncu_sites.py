#!/usr/bin/env python
"""Per inline SITE of one source line: executed warp instructions and average active threads.
usage: ncu_sites.py <report.ncu-rep> <lib.so> <kernel-substring> <file:line>"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern, want = sys.argv[1:5]
wf, wl = want.split(":"); wl = int(wl)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ia, iex, ismp, ithr = (hdr.index(k) for k in ("Address", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("pt_kernel.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
offs, cur, infn = {}, ("?", 0), False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = kern in l
    elif not infn:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
    if m: offs[int(m.group(1), 16)] = (cur, m.group(2))
base = int(data[0][ia], 16)
prev = None
for r in data:
    o = int(r[ia], 16) - base
    c, txt = offs.get(o, (("?", 0), ""))
    if c == (wf, wl):
        print("%06x  exec %12d  thr %5.1f  smp %7d  %s" % (o, int(r[iex]), float(r[ithr]), int(r[ismp]), txt[:60]))
