"""Where does a frame's tail come from?  python tools/tail_probe.py [workload] [spp] -- the timeline with / without the LPT order and
with forced numbers of express CTAs (PTB200_LIB selects the build; a -DPT_FORCE_HEAVY_RATE build switches the hand-off off)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from path_tracer_b200 import render as R
workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
sc, cam, w, h, spp, d = bench.load_workload(workload)
if len(sys.argv) > 2: spp = int(sys.argv[2])
L = R.lib()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for lpt, n in [(1, -1), (1, 17), (0, -1)]:
    L.pt_debug_set_lpt(lpt); L.pt_debug_set_express(n)
    ms = []
    for i in range(3):
        ev0.record()
        ds.render_region(cam, w, h, spp, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        ev1.record(); torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
    out = (C.c_ulonglong * 11)()
    L.pt_debug_timeline(ds._h, out)
    print("%s %d spp, lpt %d express %2d: ms %s (dry %.1f, done %.1f, regular CTAs out %.1f / %.1f, handed off %d, wait avg %.1f max %.1f ms)" % (
        workload, spp, lpt, n, " ".join("%.1f" % m for m in ms), out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4],
        out[7] / max(out[4], 1) / 1e6, out[8] / 1e6), flush=True)
    print("      service: %d rounds, %.1f rays per round, %.0f rounds per handed-off pixel, longest stay %d rounds; items scanned in place %d" % (
        out[5], out[6] / max(out[5], 1), out[6] / max(out[4], 1), out[9], out[10]), flush=True)
