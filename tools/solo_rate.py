import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scenes
from path_tracer_b200 import render as R, abi
from oracle.pyoracle import CPort
sc, cam, (w, h, spp, d) = scenes.load_c1()
L = R.lib(); L.pt_debug_set_kernel(1)
reg = abi.pt_region(300, 200, 148, 1, 1)
import ctypes as C
cp = CPort()
cp.lib.pt_oracle_render_region_ex.argtypes = [C.c_int]*4 + [C.c_void_p]*4 + [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
from path_tracer_b200.scene import camera_c
s, keep = sc.as_c(); c = camera_c(cam)
out = np.zeros((1,148,3),np.float32); ps = np.zeros((1,148),np.uint32)
cp.lib.pt_oracle_render_region_ex(w,h,spp,d,C.addressof(c),C.addressof(s),C.addressof(reg),out.ctypes.data,148*3,1,0,None,ps.ctypes.data)
print("pixel scans: max %d mean %.0f" % (ps.max(), ps.mean()))
for t in (32, 16, 8, 1):
    L.pt_debug_set_team_size(t)
    for rep in range(2):
        R.render_region(sc, cam, w, h, spp, d, reg); st = R.stats()
    print("team %2d: kernel %.2f ms -> %.2f us per round of the heaviest pixel" % (t, st["kernel_ms"], st["kernel_ms"]*1e3/ps.max()))
