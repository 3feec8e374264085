#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Regenerate tests/golden/ from the reference ITSELF.

Must run in the container that has /root/reference (it needs oracle/_ref/*,
built by `make -C oracle`).  Everything written here is data, produced by the
unmodified reference code compiled in place:

  c1_scene.ptsc.gz     the default scene + camera of src/main.cpp:67-183, captured from
                       the unmodified main.cpp by oracle/_ref/capture_main
  rng_kat.json         xorshift32 / LocalPseudoRNG known answers (xorshift.hpp, rtweekend.hpp)
  ref_renders.npz      linear fp32 framebuffers of small renders by libptref.so
                       (render_pixel<> of render.hpp:25-106, executor seeding :130-133)
  ref_hashes.json      sha256 of larger reference framebuffers (incl. the survey's
                       800x480x32 hash of the default scene) and of the USE_SINGLE_TASK
                       executor's images (oracle/_ref/libptref_st.so)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pyoracle import Ref  # noqa: E402
import scenes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SMALL = [(64, 48, 4, 50), (64, 48, 16, 3), (33, 17, 5, 7)]
SINGLE_TASK_CASES = [("c1", 64, 48, 4), ("c1", 200, 120, 4), ("shapes", 64, 48, 4), ("media", 64, 48, 4)]  # = tests/test_oracle.py


def main(full=False):
    os.makedirs(GOLDEN, exist_ok=True)
    subprocess.check_call([sys.executable, os.path.join(HERE, "decode_images.py")])
    tmp = os.path.join(HERE, "_ref", "c1_scene.ptsc")
    subprocess.check_call([os.path.join(HERE, "_ref", "capture_main"), os.path.join(HERE, "_ref", "images"), tmp])
    with open(tmp, "rb") as f, gzip.GzipFile(os.path.join(GOLDEN, "c1_scene.ptsc.gz"), "wb", 9, mtime=0) as g:
        shutil.copyfileobj(f, g)

    ref = Ref()
    kat = {"xorshift32": {}, "float_t": {}, "unit_vec": {}, "in_unit_ball": {}, "in_unit_disk": {}, "vec_t": {}}
    for seed in (2463534242, 0, 1, 2, 799, 800, 383999, 100000, 4294967295):
        kat["xorshift32"][str(seed)] = [int(v) for v in ref.xorshift(seed, 16)]
        kat["float_t"][str(seed)] = [float(v).hex() for v in ref.floats(seed, 16)]
        for kind, name in enumerate(("unit_vec", "in_unit_ball", "in_unit_disk", "vec_t")):
            kat[name][str(seed)] = [[float(c).hex() for c in v] for v in ref.vecs(seed, kind, 6)]
    with open(os.path.join(GOLDEN, "rng_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    renders = {}
    c1, c1cam, _ = scenes.load_c1()
    for (w, h, spp, d) in SMALL:
        renders["c1_%dx%dx%dx%d" % (w, h, spp, d)] = ref.render_region(c1, c1cam, w, h, spp, d)
        for name in scenes.REFERENCE_COMPATIBLE:
            sc, cam = scenes.ALL[name](w / h)
            renders["%s_%dx%dx%dx%d" % (name, w, h, spp, d)] = ref.render_region(sc, cam, w, h, spp, d)
    for seed in range(4):
        sc, cam = scenes.random_scene(seed, aspect=64 / 48)
        renders["random%d_64x48x4x50" % seed] = ref.render_region(sc, cam, 64, 48, 4, 50)
    renders["c1_200x120x16x50"] = ref.render_region(c1, c1cam, 200, 120, 16, 50)
    np.savez_compressed(os.path.join(GOLDEN, "ref_renders.npz"), **renders)

    hashes = {}
    hashes["c1_200x120x64x50"] = hashlib.sha256(ref.render_region(c1, c1cam, 200, 120, 64, 50).tobytes()).hexdigest()
    if full:  # ~35 s on 8 cores: the reference's own render<800,480,32>() entry point
        hashes["c1_800x480x32x50_render_full"] = hashlib.sha256(
            ref.render_full(c1, c1cam, 800, 480, 32, dynamic=True).tobytes()).hexdigest()
    else:
        path = os.path.join(GOLDEN, "ref_hashes.json")
        if os.path.exists(path):
            old = json.load(open(path))
            if "c1_800x480x32x50_render_full" in old:
                hashes["c1_800x480x32x50_render_full"] = old["c1_800x480x32x50_render_full"]
    # the reference's USE_SINGLE_TASK executor (render.hpp:113-122): ref_driver.cpp compiled with -DUSE_SINGLE_TASK
    st_path = os.path.join(HERE, "_ref", "libptref_st.so")
    if os.path.exists(st_path):
        st = Ref(st_path)
        for name, w, h, spp in SINGLE_TASK_CASES:
            if name == "c1":
                sc, cam = c1, c1cam
            else:
                sc, cam = scenes.ALL[name](w / h)
            img = st.render_full(sc, cam, w, h, spp, nthreads=1)
            hashes["single_task_%s_%dx%dx%dx50" % (name, w, h, spp)] = hashlib.sha256(img.tobytes()).hexdigest()
    else:
        path = os.path.join(GOLDEN, "ref_hashes.json")
        if os.path.exists(path):
            hashes.update({k: v for k, v in json.load(open(path)).items() if k.startswith("single_task_")})
    with open(os.path.join(GOLDEN, "ref_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print("golden: %d renders, hashes %s" % (len(renders), sorted(hashes)))


if __name__ == "__main__":
    main(full="--full" in sys.argv)
