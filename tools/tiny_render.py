"""A few tiny renders through the C-ABI (for compute-sanitizer runs).  usage: tiny_render.py [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from path_tracer_b200 import render as R
if len(sys.argv) > 1: R.lib().pt_debug_set_kernel(int(sys.argv[1]))
sc, cam, _ = scenes.load_c1()
a = R.render(sc, cam, 96, 64, 3, 50)
print("c1", float(a.mean()), R.stats()["scans"])
if os.environ.get("PT_TINY_LPT"):  # large enough for the cost probe, the LPT order, express CTAs and hand-offs
    a = R.render(sc, cam, 320, 192, 12, 50)
    print("c1 with LPT order and express CTAs", float(a.mean()), R.stats()["scans"], R.stats()["kernel_launches"])
for name in ("media", "moving", "shapes"):
    s, c = scenes.ALL[name](4 / 3)
    a = R.render(s, c, 48, 36, 2, 50)
    print(name, float(a.mean()), R.stats()["scans"])
# flat trees (breadth-first tree passes, grazing index): a 1 200-triangle mesh; the Cornell box (media, vector-order rays)
s, c = scenes.c4_mesh(4 / 3, nx=30, nz=10)
a = R.render(s, c, 64, 48, 2, 50)
print("mesh with trees", float(a.mean()), R.stats()["scans"])
if os.environ.get("PT_TINY_LPT"):  # enough rays per CTA for the tree lists to spill into global memory
    a = R.render(s, c, 480, 300, 2, 50)
    print("mesh with trees, full pools (tree lists spill)", float(a.mean()), R.stats()["scans"])
s, c = scenes.cornell(1.0)
a = R.render(s, c, 40, 40, 2, 50)
print("cornell", float(a.mean()), R.stats()["scans"])
# progressive rendering: two legs of one image
if hasattr(R, "render_resume"):
    sc, cam, _ = scenes.load_c1()
    from path_tracer_b200 import abi
    reg = abi.pt_region(0, 0, 64, 48, 1)
    fb, st = R.render_resume(sc, cam, 64, 48, 0, 2, 50, reg)
    fb, st = R.render_resume(sc, cam, 64, 48, 2, 4, 50, reg, state=st)
    print("resume", float(fb.mean()))
