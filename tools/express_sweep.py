"""C1 (or another workload) frame time against the number of express CTAs: python tools/express_sweep.py [workload] [spp] n1 n2 ..."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from path_tracer_b200 import render as R
workload = sys.argv[1] if len(sys.argv) > 1 else "c1"
sc, cam, w, h, spp, d = bench.load_workload(workload)
if len(sys.argv) > 2 and int(sys.argv[2]) > 0: spp = int(sys.argv[2])
L = R.lib()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in [int(a) for a in sys.argv[3:]] or [-1]:
    L.pt_debug_set_express(n)
    ms = []
    for i in range(4):
        ev0.record()
        ds.render_region(cam, w, h, spp, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        ev1.record(); torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
    out = (C.c_ulonglong * 11)()
    L.pt_debug_timeline(ds._h, out)
    tun = (C.c_int * 5)()
    L.pt_debug_frame_tuning(ds._h, tun)
    print("%s express %3d: ms %s  (dry %.2f, done %.2f, regular CTAs out %.2f / %.2f ms, handed off %d, queue wait avg %.2f max %.2f ms) fb %.6f" % (
        workload, n, " ".join("%.2f" % m for m in ms), out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4],
        out[7] / max(out[4], 1) / 1e6, out[8] / 1e6, float(fb.mean())), flush=True)
    print("      probe: heavy rate %d, express %d, mean %.2f scans/sample, heavy share %.3f, deepest sample %d" % (tun[0], tun[1], tun[2] / 1e3, tun[3] / 1e3, tun[4]), flush=True)
