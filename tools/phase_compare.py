"""Per-phase cycles of a -DPT_PHASE_TIMING build of any vintage (raw ctypes: no dependency on newer symbols).
usage: PT_PHASE_TIMING=1 python tools/phase_compare.py LIB [workload] [spp]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from path_tracer_b200 import abi
from path_tracer_b200.scene import camera_c
lib, workload = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "c1")
sc, cam, w, h, spp, d = bench.load_workload(workload)
if len(sys.argv) > 3: spp = int(sys.argv[3])
L = C.CDLL(lib)
L.pt_scene_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
L.pt_render_region_device.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
L.pt_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
s, keep = sc.as_c(); c = camera_c(cam)
hnd = C.c_void_p()
assert L.pt_scene_upload(C.addressof(s), 0, C.byref(hnd)) == 0
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
region = abi.pt_region(0, 0, w, h, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print("==", os.path.basename(lib), workload, spp, flush=True)
for i in range(3):
    ev0.record()
    assert L.pt_render_region_device(hnd, w, h, spp, d, C.addressof(c), C.addressof(region), C.c_void_p(fb.data_ptr()), w * 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    ev1.record(); torch.cuda.synchronize()
    out = (C.c_ulonglong * 11)()
    sys.stderr.flush()
    L.pt_debug_timeline(hnd, out)
    print("   %.2f ms; dry %.2f done %.2f, regular CTAs out %.2f / %.2f, handed off %d" % (ev0.elapsed_time(ev1), out[0] / 1e6, out[1] / 1e6, out[2] / 1e6, out[3] / 1e6, out[4]), flush=True)
