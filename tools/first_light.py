"""Scratch GPU check: CUDA vs C oracle on small renders, then timing of C1."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from path_tracer_b200 import render as R
from oracle.pyoracle import CPort, compare

cp = CPort()
print("devices", R.device_count(), "fp32 peak", R.measure_fp32_peak(0), flush=True)
bad = 0
def check(name, sc, cam, w, h, spp, d):
    global bad
    t = time.time(); a = R.render(sc, cam, w, h, spp, d); tg = time.time() - t
    t = time.time(); b, cnt = cp.render(sc, cam, w, h, spp, d); tc = time.time() - t
    mae, psnr, same = compare(a, b)
    st = R.stats()
    ok = mae <= 1e-3 and psnr >= 50
    bad += (not ok)
    print("%-16s %4dx%-4d spp %-4d d %-2d  mae %.3e psnr %6.2f same %.5f  gpu %.3fs (kernel %.2f ms) cpu %.2fs scans gpu/cpu %d/%d %s"
          % (name, w, h, spp, d, mae, psnr, same, tg, st["kernel_ms"], tc, st["scans"], cnt.scans, "OK" if ok else "FAIL"), flush=True)
    return a, b

for name, fn in scenes.ALL.items():
    sc, cam = fn(64 / 48)
    check(name, sc, cam, 64, 48, 8, 50)
    check(name, sc, cam, 33, 17, 5, 7)
for seed in range(4):
    sc, cam = scenes.random_scene(seed, aspect=64 / 48)
    check("random%d" % seed, sc, cam, 64, 48, 8, 50)
sc, cam, meta = scenes.load_c1()
a, b = check("c1", sc, cam, 200, 120, 16, 50)
check("c1", sc, cam, 800, 480, 2, 50)
# timing, C1 full size
for spp in (10, 100):
    t = time.time(); a = R.render(sc, cam, 800, 480, spp, 50); dt = time.time() - t
    st = R.stats()
    print("C1 800x480 spp %d: wall %.3f s, kernel %.2f ms -> %.1f Mpaths/s (kernel), %.1f (e2e); scans/path %.3f; h2d %.2f ms d2h %.2f ms"
          % (spp, dt, st["kernel_ms"], 800 * 480 * spp / st["kernel_ms"] / 1e3, 800 * 480 * spp / dt / 1e6,
             st["scans"] / (800 * 480 * spp), st["h2d_ms"], st["d2h_ms"]), flush=True)
np.save(os.path.join(ROOT, "gpurun_out", "c1_800x480x100_gpu.npy"), a)
print("FAILURES", bad)
