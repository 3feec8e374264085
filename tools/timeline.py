import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, scenes
from path_tracer_b200 import render as R
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
if len(sys.argv) > 2: R.lib().pt_debug_set_kernel(int(sys.argv[2]))
sc, cam, (w, h, _, d) = scenes.load_c1()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
L = R.lib()
for i in range(3):
    ds.render_region(cam, w, h, spp, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, 0)
    out = (C.c_ulonglong * 5)()
    L.pt_debug_timeline(ds._h, out)
    print("spp %d: queue dry at %.2f ms, done at %.2f ms -> tail %.1f%%; rate %.1f Mpaths/s; first/last CTA out of regular work %.2f / %.2f ms; heavy pixels %d" % (
        spp, out[0] / 1e6, out[1] / 1e6, 100 * (out[1] - out[0]) / out[1], w * h * spp / (out[1] / 1e3), out[2] / 1e6, out[3] / 1e6, out[4]))
