#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Decode the reference's two image textures
(/root/reference/images/{Xilinx.jpg,SYCL.png}, loaded by main.cpp:133,145) to
binary PPM with PIL, into oracle/_ref/images/.  stb (the reference's decoder) is
not in this image; which decoder produced the texels is irrelevant for parity as
long as the oracle and the CUDA path read the same bytes (the texels end up
embedded in tests/golden/c1_scene.ptsc)."""
import os
import sys

from PIL import Image

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/images"
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "_ref", "images")
os.makedirs(dst, exist_ok=True)
for name in sorted(os.listdir(src)):
    im = Image.open(os.path.join(src, name)).convert("RGB")
    with open(os.path.join(dst, name + ".ppm"), "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % im.size)
        f.write(im.tobytes())
    print(name, im.size)
