import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, scenes
from path_tracer_b200 import render as R
# emulate rank 0 of an N-GPU strong-scaling run on one GPU: rows 0::N
L = R.lib()
sc, cam, (w, h, spp, d) = scenes.load_c1()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (1, 2, 4, 8):
    reg = R.rows_region(w, h, 0, n)
    for i in range(2):
        ev0.record()
        ds.render_region(cam, w, h, spp, d, reg, fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        ev1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 5)()
    L.pt_debug_timeline(ds._h, out)
    ms = ev0.elapsed_time(ev1)
    print("N=%d (rows 0::%d, %d pixels): total %.2f ms -> speed-up %.2fx if all ranks alike; main kernel: dry %.2f, done %.2f; regular work out %.2f / %.2f; handed off %d" % (
        n, n, reg.w * reg.h, ms, 0, out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4]))
