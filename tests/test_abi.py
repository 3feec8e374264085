"""CPU-only: the C-ABI library loads, exports every symbol include/pt_abi.h declares, validates its
arguments, and FAILS LOUDLY (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import scenes
from path_tracer_b200 import Scene, abi, camera_c, make_camera
from path_tracer_b200 import render as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pt_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pt_[a-z0-9_]+|render)\s*\(", hdr))
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    L = R.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.pt_abi_version() == abi.PT_ABI_VERSION


def test_pod_layouts_match_header():
    hdr = open(os.path.join(ROOT, "include", "pt_abi.h")).read()
    assert "float lens_radius;" in hdr and C.sizeof(abi.pt_camera) == 96
    assert abi.TEXTURE_DT.fields["offset"][1] == 40 and abi.TEXTURE_DT.itemsize == 56
    assert abi.SPHERE_DT.fields["material"][1] == 36


def test_argument_validation_precedes_device_use():
    sc, cam = scenes.spheres_basic()
    s, keep = sc.as_c()
    c = camera_c(cam)
    fb = np.zeros((4, 4, 3), np.float32)
    L = R.lib()
    assert L.pt_render(4, 4, 1, 50, None, C.addressof(s), fb.ctypes.data) == abi.PT_ERR_INVALID_ARGUMENT
    assert L.pt_render(0, 4, 1, 50, C.addressof(c), C.addressof(s), fb.ctypes.data) == abi.PT_ERR_INVALID_ARGUMENT
    assert L.pt_render(4, 4, 0, 50, C.addressof(c), C.addressof(s), fb.ctypes.data) == abi.PT_ERR_INVALID_ARGUMENT
    assert b"null" in L.pt_last_error() or b"positive" in L.pt_last_error()
    assert L.pt_set_num_gpus(0) == abi.PT_ERR_INVALID_ARGUMENT


def test_no_cpu_fallback_without_gpu():
    if R.device_count() > 0:
        pytest.skip("a GPU is visible")
    sc, cam = scenes.spheres_basic()
    with pytest.raises(R.PathTracerError) as e:
        R.render(sc, cam, 8, 8, 1, 50)
    assert e.value.code == abi.PT_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(R.PathTracerError) as e:
        R.render_single_task(sc, cam, 8, 8, 1, 50)
    assert e.value.code == abi.PT_ERR_NO_DEVICE
    from path_tracer_b200 import abi as A
    with pytest.raises(R.PathTracerError) as e:
        R.render_resume(sc, cam, 8, 8, 0, 1, 50, A.pt_region(0, 0, 8, 8, 1))
    assert e.value.code == abi.PT_ERR_NO_DEVICE
    with pytest.raises(R.PathTracerError):
        R.DeviceScene(sc, 0)
    with pytest.raises(R.PathTracerError):
        R.measure_fp32_peak(0)


def test_ptscene_roundtrip(tmp_path, c1):
    sc, cam, meta = c1
    assert meta == (800, 480, 100, 50) and sc.n_hittables == 496
    a = sc.arrays()
    assert len(a["spheres"]) == 490 and len(a["rects"]) == 1 and len(a["triangles"]) == 4
    assert len(a["boxes"]) == 1 and len(a["media"]) == 1 and sc.texture_bytes.size == 3 + 3 * (1024 * 512 + 1280 * 559)
    assert int((a["spheres"]["time0"] != a["spheres"]["time1"]).sum()) == 178  # moving spheres
    p = tmp_path / "x.ptsc"
    sc.save(str(p), cam, *meta)
    sc2, cam2, meta2 = Scene.load(str(p))
    assert meta2 == meta and cam2.tobytes() == cam.tobytes()
    for k, v in sc.arrays().items():
        for field in v.dtype.names:  # field-wise: struct padding bytes are not data
            assert np.array_equal(v[field], sc2.arrays()[k][field]), (k, field)
    assert np.array_equal(sc.texture_bytes, sc2.texture_bytes)


def test_python_camera_equals_reference_camera(c1):
    """make_camera restates camera.hpp:67-87; the fixture's camera was built by the reference constructor."""
    _, cam, _ = c1
    look_from, look_at = np.float32([13, 3, 3]), np.float32([0, -1, 0])
    d = look_at - look_from
    focus = np.sqrt(np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
    mine = make_camera(look_from, look_at, (0, 1, 0), 40.0, np.float32(800) / np.float32(480), 0.04, focus, 0.0, 1.0)
    assert mine.tobytes() == cam.tobytes()


def test_python_camera_equals_reference_constructor(ref):
    rs = np.random.RandomState(5)
    for _ in range(20):
        lf, la = rs.uniform(-10, 10, 3), rs.uniform(-2, 2, 3)
        args = (lf, la, (0, 1, 0), float(rs.uniform(15, 70)), float(rs.uniform(0.8, 2.0)), float(rs.uniform(0, 0.2)),
                float(rs.uniform(1, 15)), 0.0, float(rs.rand()))
        # harness-side helper: numpy's float32 tan is not glibc's tanf, so allow an ulp or two here;
        # the product's C++ camera (path_tracer_b200/include/pt/scene.hpp) is checked bit for bit in test_host.py
        a = np.frombuffer(make_camera(*args).tobytes(), np.float32)
        b = np.frombuffer(ref.make_camera(*args).tobytes(), np.float32)
        assert np.allclose(a, b, rtol=2e-6, atol=1e-6)
