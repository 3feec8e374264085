"""Time pt_render of a workload with a given libptb200 build (experiments): python tools/variant_time.py LIB [workload] [spp] [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from path_tracer_b200 import abi
from path_tracer_b200.scene import camera_c
sys.argv += [None] * 4
lib_path, workload, spp_arg, reps = sys.argv[1], sys.argv[2] or "c1", sys.argv[3], int(sys.argv[4] or 4)
import bench
sc, cam, w, h, spp, d = bench.load_workload(workload)
if spp_arg:
    spp = int(spp_arg)
L = C.CDLL(lib_path)
L.pt_render.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
L.pt_get_stats.argtypes = [C.c_void_p]
L.pt_last_error.restype = C.c_char_p
s, keep = sc.as_c()
c = camera_c(cam)
out = np.empty((h, w, 3), np.float32)
ms = []
for i in range(reps):
    rc = L.pt_render(w, h, spp, d, C.addressof(c), C.addressof(s), out.ctypes.data)
    assert rc == 0, L.pt_last_error()
    st = abi.pt_stats(); L.pt_get_stats(C.byref(st)); ms.append(st.kernel_ms)
print("%-28s %s %dx%dx%d: kernel ms %s -> best %.1f Mpaths/s, fb mean %.6f, scans %d" % (os.path.basename(lib_path), workload, w, h, spp,
      " ".join("%.2f" % m for m in ms), w * h * spp / min(ms) / 1e3, float(out.mean()), st.scans), flush=True)
