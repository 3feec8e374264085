"""Emulate every rank of an N-GPU WEAK-scaling bench step on one GPU: the default scene at 800 x (480 N), rank r
renders rows r::N.  usage: timeline_weak.py N"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, scenes
from path_tracer_b200 import render as R, make_camera
n = int(sys.argv[1])
L = R.lib()
sc, cam, (w, h, spp, d) = scenes.load_c1()
H = h * n
cam_n = cam  # bench.py keeps the camera: the same view sampled with N times the rows
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(n):
    reg = R.rows_region(w, H, r, n)
    for i in range(2):
        ev0.record()
        ds.render_region(cam_n, w, H, spp, d, reg, fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        ev1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 16)()
    L.pt_debug_timeline(ds._h, out)
    print("rank %d of %d: total %.2f ms; dry %.2f, done %.2f; regular work out %.2f / %.2f; handed off %d; longest stay %d" % (
        r, n, ev0.elapsed_time(ev1), out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4], out[9]), flush=True)
