"""Two tiny renders (default scene, RTIOW) for a quick compute-sanitizer pass of the last kernel change."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from path_tracer_b200 import render as R
sc, cam, _ = scenes.load_c1()
a = R.render(sc, cam, 96, 64, 2, 50)
print("c1", float(a.mean()), R.stats()["scans"])
s, c = scenes.rtiow(16 / 9)
a = R.render(s, c, 128, 72, 2, 50)
print("rtiow", float(a.mean()), R.stats()["scans"])
