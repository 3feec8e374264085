#!/usr/bin/env python
"""Join an ncu source-page (SASS) dump with nvdisasm line info: per source line samples / instructions.
usage: ncu_lines.py <report.ncu-rep> <lib.so> [kernel-substring]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "render_kernelILb1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ia, isrc, iex, ismp, ithr = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
ino = hdr.index("stall_no_inst")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("pt_kernel.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, infn = [], ("?", 0), False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = kern in l
    elif not infn:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
base = int(data[0][ia], 16)
off2line = {o: c for o, c, _ in lines}
agg = {}
tot_s = tot_e = 0
for r in data:
    o = int(r[ia], 16) - base
    key = off2line.get(o, ("?", 0))
    a = agg.setdefault(key, [0, 0, 0.0, 0])
    s, e = int(r[ismp]), int(r[iex])
    a[0] += s; a[1] += e; a[2] += float(r[ithr]) * e; a[3] += int(r[ino])
    tot_s += s; tot_e += e
src_cache = {}
def text(f, n):
    for d in ("path_tracer_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:70] if 0 < n <= len(src_cache[p]) else ""
    return ""
print("total samples %d, warp instr %d" % (tot_s, tot_e))
print("%-26s %7s %7s %6s %7s  %s" % ("file:line", "smp%", "instr%", "thr", "noinst%", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "45"))]:
    print("%-26s %7.2f %7.2f %6.1f %7.1f  %s" % ("%s:%d" % key, 100 * a[0] / tot_s, 100 * a[1] / tot_e, a[2] / max(a[1], 1), 100 * a[3] / max(a[0], 1), text(*key)))
