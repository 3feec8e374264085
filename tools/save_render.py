"""Render a workload at a given spp through pt_render and save the framebuffer: python tools/save_render.py workload spp out.npy"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from path_tracer_b200 import render as R
sc, cam, w, h, spp, d = bench.load_workload(sys.argv[1])
spp = int(sys.argv[2])
img = R.render(sc, cam, w, h, spp, d)
st = R.stats()
print("%s %dx%dx%d: kernel %.2f ms, scans %d, NaN pixels %d, mean %.7f" % (sys.argv[1], w, h, spp, st["kernel_ms"], st["scans"], int(np.isnan(img).any(axis=2).sum()), float(np.nanmean(img))))
np.save(sys.argv[3], img.astype(np.float32))
