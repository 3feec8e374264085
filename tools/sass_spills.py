#!/usr/bin/env python
"""Which source lines carry local-memory (spill) instructions in a kernel's SASS?  usage: sass_spills.py lib.so kernel-substring"""
import os, re, subprocess, sys, tempfile, collections
lib, kern = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in d: continue
    infn, cur = False, ("?", 0)
    agg = collections.Counter(); n_inst = 0
    for l in d.splitlines():
        if l.startswith("\t.section\t.text."):
            infn = kern in l
        elif not infn: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
        if m:
            n_inst += 1
            op = m.group(2)
            if re.search(r"\b(LDL|STL)\b", op): agg[(cur, "LDL" if "LDL" in op else "STL")] += 1
    print("%s: %d instructions, %d local-memory instructions" % (f, n_inst, sum(agg.values())))
    for (c, k), n in sorted(agg.items(), key=lambda kv: (kv[0][0][0], kv[0][0][1])):
        print("  %s:%d %s x%d" % (c[0], c[1], k, n))
