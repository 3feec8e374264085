"""(ray, chunk) items that did not fit the item list and were scanned in place, per workload (pt_debug_timeline out[10]).
usage: python tools/inplace_items.py LIB workload spp"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from path_tracer_b200 import abi
from path_tracer_b200.scene import camera_c
lib, workload, spp = sys.argv[1], sys.argv[2], int(sys.argv[3])
sc, cam, w, h, _, d = bench.load_workload(workload)
L = C.CDLL(lib)
L.pt_scene_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
L.pt_render_region_device.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
L.pt_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
L.pt_scene_read_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
s, keep = sc.as_c(); c = camera_c(cam)
hnd = C.c_void_p()
assert L.pt_scene_upload(C.addressof(s), 0, C.byref(hnd)) == 0
fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
region = abi.pt_region(0, 0, w, h, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    ev0.record()
    assert L.pt_render_region_device(hnd, w, h, spp, d, C.addressof(c), C.addressof(region), C.c_void_p(fb.data_ptr()), w * 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    ev1.record(); torch.cuda.synchronize()
out = (C.c_ulonglong * 11)()
L.pt_debug_timeline(hnd, out)
paths, scans = C.c_ulonglong(), C.c_ulonglong()
L.pt_scene_read_counters(hnd, C.byref(paths), C.byref(scans), 0)
print("%-16s %s x%d: %.2f ms, %d scans, %d items scanned in place (%.4f per scan)" % (os.path.basename(lib), workload, spp, ev0.elapsed_time(ev1), scans.value,
      out[10], out[10] / max(scans.value, 1)), flush=True)
