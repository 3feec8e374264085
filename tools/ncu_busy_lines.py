#!/usr/bin/env python
"""Like ncu_lines.py, for instructions executed at most <cap> times (the CTAs that do real work in a small-region
run), ranking source lines by non-barrier samples.  usage: ncu_busy_lines.py <rep> <lib.so> <kernel-substr> <cap>"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern, cap = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ia, iex, ismp, ibar = (hdr.index(k) for k in ("Address", "Instructions Executed", "# Samples", "stall_barrier"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("pt_kernel.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
offs, cur, infn = {}, ("?", 0), False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."): infn = kern in l
    elif not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
    if m: offs[int(m.group(1), 16)] = cur
base = int(data[0][ia], 16)
agg = {}; tot = 0
for r in data:
    e = int(r[iex])
    if e == 0 or e > cap: continue
    s = int(r[ismp]) - int(r[ibar])
    key = offs.get(int(r[ia], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += e; tot += s
src = {}
def text(f, n):
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "path_tracer_b200/csrc", f)
    if os.path.exists(p):
        if p not in src: src[p] = open(p).read().splitlines()
        return src[p][n - 1].strip()[:80] if 0 < n <= len(src[p]) else ""
    return ""
print("non-barrier samples", tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "50"))]:
    print("%-24s %6.2f%% exec %8d  %s" % ("%s:%d" % key, 100.0 * a[0] / max(tot, 1), a[1], text(*key)))
