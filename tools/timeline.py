import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, scenes
from path_tracer_b200 import render as R
# usage: timeline.py spp kernel(0 wave,1 lane) lpt(0/1) [express [team [cull]]]
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = R.lib()
if len(sys.argv) > 2: L.pt_debug_set_kernel(int(sys.argv[2]))
if len(sys.argv) > 3: L.pt_debug_set_lpt(int(sys.argv[3]))
if len(sys.argv) > 4: L.pt_debug_set_express(int(sys.argv[4]))
if len(sys.argv) > 5: L.pt_debug_set_team_size(int(sys.argv[5]))
if len(sys.argv) > 6: L.pt_debug_set_cull(int(sys.argv[6]))
sc, cam, (w, h, _, d) = scenes.load_c1()
ds = R.DeviceScene(sc, 0)
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    ev0.record()
    ds.render_region(cam, w, h, spp, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
    ev1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 11)()
    L.pt_debug_timeline(ds._h, out)
    ms = ev0.elapsed_time(ev1)
    print("args %s: total %.2f ms (%.1f Mpaths/s); main kernel: dry %.2f, done %.2f; CTAs out of regular work %.2f / %.2f ms; handed off %d" % (
        " ".join(sys.argv[1:]), ms, w * h * spp / ms / 1e3, out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4]))
    print("    service: %d rounds, %.1f rays per round, %.0f rounds per handed-off pixel (longest stay %d); wait in queue avg %.2f ms, max %.2f ms" % (
        out[5], out[6] / max(out[5], 1), out[6] / max(out[4], 1), out[9], out[7] / max(out[4], 1) / 1e6, out[8] / 1e6))
