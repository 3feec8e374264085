"""The C++ host mirror: the reference's UNMODIFIED src/main.cpp, compiled against
path_tracer_b200/include + compat (build/sycl-rt-b200, built by `make host` where /root/reference exists),
must build exactly the scene the reference headers build, and render it through libptb200.so."""
import os
import subprocess

import numpy as np
import pytest

import scenes
from path_tracer_b200 import Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "build", "sycl-rt-b200")
IMAGES = os.path.join(ROOT, "oracle", "_ref", "images")

needs_binary = pytest.mark.skipif(not (os.path.exists(BINARY) and os.path.isdir(IMAGES)),
                                  reason="build/sycl-rt-b200 or the decoded reference images are not built")


def _run(tmp_path, extra_env):
    env = dict(os.environ, PT_IMAGE_DIR=IMAGES, **extra_env)
    return subprocess.run([BINARY], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)


@needs_binary
def test_unmodified_main_builds_the_reference_scene(tmp_path, c1):
    """Scene + camera flattened by OUR headers == scene captured from the REFERENCE headers, field for field."""
    dump = str(tmp_path / "scene.ptsc")
    r = _run(tmp_path, {"PT_DUMP_SCENE": dump})
    sc, cam, meta = Scene.load(dump)
    want, want_cam, want_meta = c1
    assert meta == want_meta and cam.tobytes() == want_cam.tobytes()
    for k, v in want.arrays().items():
        got = sc.arrays()[k]
        assert got.shape == v.shape, k
        for field in v.dtype.names:
            if field != "_pad":
                assert np.array_equal(got[field], v[field]), (k, field)
    assert np.array_equal(sc.texture_bytes, want.texture_bytes)
    # without a GPU the run must fail loudly, never fall back to a CPU render
    from path_tracer_b200 import render as R
    if R.device_count() == 0:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_unmodified_main_in_single_task_mode(tmp_path, cport):
    """The reference's application compiled with -DUSE_SINGLE_TASK (CMake option USE_SINGLE_TASK, build_parameters.hpp:5-9)
    against the host mirror: render<>() goes to pt_render_single_task, and out.png is the single-task image of the oracle
    (one generator for the whole image, render.hpp:113-122) for the scene the binary itself built."""
    from PIL import Image
    binary = BINARY + "-st"
    if not os.path.exists(binary):
        pytest.skip("build/sycl-rt-b200-st not built (needs /root/reference at build time)")
    dump = str(tmp_path / "scene.ptsc")
    env = dict(os.environ, PT_IMAGE_DIR=IMAGES, PT_DUMP_SCENE=dump)
    r = subprocess.run([binary], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    sc, cam, (w, h, spp, d) = Scene.load(dump)
    assert (w, h, spp, d) == (32, 24, 100, 50)
    img = np.asarray(Image.open(str(tmp_path / "out.png")).convert("RGB")).astype(np.int32)
    want, _ = cport.render_single_task(sc, cam, w, h, spp, d)
    want8 = (256 * np.clip(np.sqrt(want), 0.0, 0.999)).astype(np.int32)
    assert np.array_equal(img[::-1], want8)  # (bit-identical pixels: the tone map sees the same floats)
    par, _ = cport.render(sc, cam, w, h, spp, d)
    assert not np.array_equal(img[::-1], (256 * np.clip(np.sqrt(par), 0.0, 0.999)).astype(np.int32))


@needs_binary
@pytest.mark.gpu
def test_unmodified_main_renders_out_png(tmp_path, cport, c1):
    """End to end: ./sycl-rt -> out.png (main.cpp:33-59 tone map and flip), checked against the oracle."""
    from PIL import Image
    r = _run(tmp_path, {})
    assert r.returncode == 0, r.stderr
    img = np.asarray(Image.open(str(tmp_path / "out.png")).convert("RGB")).astype(np.int32)
    sc, cam, (w, h, spp, d) = c1
    assert img.shape == (h, w, 3)
    from path_tracer_b200 import abi
    rows = abi.pt_region(0, 7, w, (h - 7 + 39) // 40, 40)
    want, _ = cport.render_region(sc, cam, w, h, spp, d, rows)
    want8 = (256 * np.clip(np.sqrt(want), 0.0, 0.999)).astype(np.int32)
    got8 = img[::-1][7::40]  # the PNG is written top row first (main.cpp:41)
    diff = np.abs(got8 - want8)
    assert diff.max() <= 2 and (diff > 0).mean() < 0.01
