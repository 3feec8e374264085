#!/usr/bin/env python
"""Per source line: samples, instructions, chosen stall columns, local-memory sectors (ncu source page joined with
nvdisasm -g line info).  usage: ncu_lines2.py <report.ncu-rep> <lib.so> <kernel-substring> [sort-column]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
sort = sys.argv[4] if len(sys.argv) > 4 else "# Samples"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern in d:
            dis = d
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
base = int(data[0][col["Address"]], 16)
off2line = {o: c for o, c, _ in lines}
want = ["# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_no_inst", "stall_lg",
        "L2 Theoretical Sectors Local", "L1 Wavefronts Shared Excessive"]
agg, tot = {}, [0] * len(want)
for r in data:
    key = off2line.get(int(r[col["Address"]], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0] * len(want))
    for j, w in enumerate(want):
        try:
            v = int(float(r[col[w]]))
        except (ValueError, KeyError):
            v = 0
        a[j] += v; tot[j] += v
src_cache = {}
def text(f, n):
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "path_tracer_b200/csrc", f)
    if os.path.exists(p):
        if p not in src_cache: src_cache[p] = open(p).read().splitlines()
        return src_cache[p][n - 1].strip()[:64] if 0 < n <= len(src_cache[p]) else ""
    return ""
print("totals: " + ", ".join("%s %d" % (w, t) for w, t in zip(want, tot)))
print("%-22s %6s %6s %6s %6s %6s %6s %6s %6s %8s %8s  %s" % ("file:line", "smp%", "ins%", "lsb%", "ssb%", "bar%", "wait%", "noi%", "lg%", "local", "shexc", "source"))
si = want.index(sort)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][si])[:int(os.environ.get("TOP", "40"))]:
    pc = [100.0 * a[j] / max(tot[j], 1) for j in range(8)]
    print("%-22s %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %8d %8d  %s" % ("%s:%d" % key, *pc, a[8], a[9], text(*key)))
