"""Non-finite pixels of a full-size render, checked against the oracle pixel by pixel:
python tools/nonfinite_pixels.py [workload] [spp] -- renders the workload through pt_render, lists the pixels whose
value is NaN / Inf, runs the oracle on each of them (and on their 8 neighbours) at the same spp and compares the bits.
The reference produces NaNs of its own (a zero-length scatter direction, a 0/0 rectangle plane: DESIGN.md section 3);
parity means reproducing exactly those."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from path_tracer_b200 import abi, render as R
from oracle.pyoracle import CPort
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else "c5"
sc, cam, w, h, spp, d = bench.load_workload(workload)
if len(sys.argv) > 2:
    spp = int(sys.argv[2])
fb = R.render(sc, cam, w, h, spp, d)
bad = np.argwhere(~np.isfinite(fb).all(axis=2))
print("%s %dx%dx%d: %d non-finite pixel(s), mean of the finite ones %.6f" % (workload, w, h, spp, len(bad), float(fb[np.isfinite(fb).all(axis=2)].mean())), flush=True)
oracle = CPort()
mismatch = 0
for y, x in bad[:16]:
    x0, y0 = max(0, x - 1), max(0, y - 1)
    x1, y1 = min(w, x + 2), min(h, y + 2)
    region = abi.pt_region(int(x0), int(y0), int(x1 - x0), int(y1 - y0), 1)
    want, _ = oracle.render_region(sc, cam, w, h, spp, d, region, nthreads=bench.host_threads())
    got = fb[y0:y1, x0:x1]
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    mismatch += int((~same).sum())
    print("  pixel (%d, %d): GPU %s oracle %s; 3x3 neighbourhood words equal: %d of %d" % (x, y, fb[y, x], want[y - y0, x - x0], int(same.sum()), same.size), flush=True)
print("mismatching words: %d" % mismatch)
sys.exit(1 if mismatch else 0)
