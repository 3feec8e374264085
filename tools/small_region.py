import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, scenes
from path_tracer_b200 import render as R, abi
# usage: small_region.py x0 y0 w h spp  -> time of a small region (short rounds: the latency of one round)
x0, y0, rw, rh, spp = (int(a) for a in sys.argv[1:6])
L = R.lib()
if len(sys.argv) > 6: L.pt_debug_set_kernel(int(sys.argv[6]))
if len(sys.argv) > 7: L.pt_debug_set_team_size(int(sys.argv[7]))
sc, cam, (w, h, _, d) = scenes.load_c1()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((rh, rw, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    ds.counters(reset=True)
    ev0.record()
    ds.render_region(cam, w, h, spp, d, abi.pt_region(x0, y0, rw, rh, 1), fb.data_ptr(), rw * 3, torch.cuda.current_stream().cuda_stream)
    ev1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 11)()
    L.pt_debug_timeline(ds._h, out)
    ms = ev0.elapsed_time(ev1)
    paths, scans = ds.counters(reset=True)
    print("region %dx%d+%d+%d spp %d: %.3f ms, %d scans (%.1f per pixel), handed off %d; service rounds %d (%.1f rays), longest stay %d" % (
        rw, rh, x0, y0, spp, ms, scans, scans / (rw * rh), out[4], out[5], out[6] / max(out[5], 1), out[9]))
