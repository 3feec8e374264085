"""sin / cos / log / pow(x, 5) / asin / atan2 exactly as the reference's host computes them
(path_tracer_b200/csrc/pt_glibc_math.cuh).

CPU: the header, compiled for the host, against the running libm on ALL 2^32 binary32 arguments of sinf, cosf, logf,
powf(x, 5), asinf and atanf, and on 400 million (y, x) pairs of atan2f (tests/host/math_check.cpp; 10-40 s per function
on 8 cores).  GPU: the device build against libm's values on several million arguments -- random bit patterns plus the
ranges the renderer uses."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libmath_check.so")
SRC = os.path.join(ROOT, "tests", "host", "math_check.cpp")
DEPS = [SRC, os.path.join(ROOT, "tests", "host", "pt_hostshim.h")] + \
    [os.path.join(ROOT, "path_tracer_b200", "csrc", n) for n in ("pt_glibc_math.cuh", "pt_device.cuh")]
NAMES = ["sinf", "cosf", "logf", "powf(x, 5)", "asinf", "atanf"]


def harness():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w",
                               "-I" + os.path.join(ROOT, "tests", "host"), "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "path_tracer_b200", "csrc"), SRC, "-o", SO, "-lm"])
    lib = C.CDLL(SO)
    lib.math_check.restype = C.c_uint64
    lib.math_check.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.math_libm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.math_check_atan2.restype = C.c_uint64
    lib.math_check_atan2.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("kind", range(6))
def test_restatement_equals_libm_on_every_float(kind):
    """Exhaustive: every one of the 2^32 bit patterns (NaN == NaN)."""
    bad = C.c_uint32()
    n = harness().math_check(kind, 0, 0xFFFFFFFF, 1, C.byref(bad))
    assert n == 0, (NAMES[kind], n, hex(bad.value))


def test_atan2_restatement_equals_libm():
    """Two arguments: 400 million pairs (random bit patterns, comparable magnitudes, points of the unit circle -- what
    sphere.hpp:13-17 passes) and every pair of 16 special values."""
    by, bx = C.c_uint32(), C.c_uint32()
    n = harness().math_check_atan2(1, 400_000_000, C.byref(by), C.byref(bx))
    assert n == 0, (n, hex(by.value), hex(bx.value))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", range(6))
def test_device_equals_libm(kind):
    """kind: 0 sin, 1 cos, 2 log, 3 pow(x, 5), 4 asin, 5 atan2 (pairs)."""
    from path_tracer_b200 import render as R
    rs = np.random.RandomState(17 + kind)
    parts = [rs.randint(0, 2 ** 32, size=3_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
             rs.uniform(-7, 7, 1_000_000).astype(np.float32),          # theta, phi of in_unit_ball (rtweekend.hpp:70-80)
             rs.uniform(-3000, 3000, 1_000_000).astype(np.float32),    # 10 x coordinate of the checker texture (texture.hpp:43)
             rs.uniform(0, 1, 1_000_000).astype(np.float32),           # log of a uniform draw (constant_medium.hpp:65)
             rs.uniform(-1, 2, 1_000_000).astype(np.float32),          # 1 - cosine of the Schlick term, asin of a normal's y
             np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, 1e-38, 120.0, -120.0, 0.785398, 2.0 ** -12, 0.5, 0.975], np.float32)]
    x = np.concatenate(parts)
    if kind == 5:  # atan2: pairs; half of them points of the unit circle (sphere.hpp:15)
        a = rs.uniform(0, 2 * np.pi, x.size // 4)
        x = np.concatenate([x[: x.size // 2 * 2], np.stack([np.sin(a), np.cos(a)], axis=1).astype(np.float32).reshape(-1)])
    x = np.ascontiguousarray(x)
    n = x.size // 2 if kind == 5 else x.size
    want = np.zeros(n, np.float32)
    harness().math_libm({0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 6}[kind], n, x.ctypes.data, want.ctypes.data)
    got = np.zeros(n, np.float32)
    L = R.lib()
    L.pt_debug_math.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    assert L.pt_debug_math({0: 0, 1: 1, 2: 2, 3: 5, 4: 3, 5: 4}[kind], n, x.ctypes.data, got.ctypes.data) == 0
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (kind, int((~same).sum()), got[~same][:5], want[~same][:5])
