"""Culling is invisible (CPU): the product's closest-hit scan -- path_tracer_b200/csrc/pt_prims.cuh compiled by
g++ through tests/host/pt_hostshim.h -- returns the same winner (t bits, object, RNG state) with chunk boxes, flat
trees and the grazing index as with all of them switched off, for camera rays, scattered rays and rays built to
attack the margins (tests/host/scan_check.cpp).  No GPU and no oracle involved: this pins the PROOF obligations of
DESIGN.md ("chunk culling", "flat culling") on millions of rays per scene; parity of the scan itself with the
reference is the GPU tests' job."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import scenes
from path_tracer_b200.scene import camera_c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libscan_check.so")
SRC = [os.path.join(ROOT, "tests", "host", "scan_check.cpp"), os.path.join(ROOT, "path_tracer_b200", "csrc", "pt_pack.cpp")]
DEPS = SRC + [os.path.join(ROOT, "tests", "host", "pt_hostshim.h")] + \
    [os.path.join(ROOT, "path_tracer_b200", "csrc", n) for n in ("pt_prims.cuh", "pt_device.cuh", "pt_stage.cuh", "pt_packed.h", "pt_pack.h")]


class Result(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("mismatches", C.c_uint64), ("hits", C.c_uint64), ("flat_nodes", C.c_uint64),
                ("graze_nodes", C.c_uint64), ("triangle_tests", C.c_uint64), ("graze_tests", C.c_uint64),
                ("brute_triangle_tests", C.c_uint64), ("bad_ray", C.c_float * 7), ("t_cull", C.c_float),
                ("t_brute", C.c_float), ("id_cull", C.c_int), ("id_brute", C.c_int), ("n_trees", C.c_int),
                ("tree_levels", C.c_int), ("tree_leaves", C.c_int), ("gtree_leaves", C.c_int)]


def load_harness():
    """Build (when stale) and load tests/host/scan_check.cpp."""
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w",
                               "-I" + os.path.join(ROOT, "tests", "host"), "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "path_tracer_b200", "csrc")] + SRC + ["-o", SO])
    lib = C.CDLL(SO)
    lib.scan_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.scan_check_make_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    return lib


def make_rays(sc, cam, mode, seed, n):
    """The harness's adversarial rays as arrays: ([n, 7] float32 {o, d, time}, [n] uint32 generator states)."""
    lib = load_harness()
    s, keep = sc.as_c()
    c = camera_c(cam)
    rays = np.zeros((n, 7), dtype=np.float32)
    seeds = np.zeros(n, dtype=np.uint32)
    assert lib.scan_check_make_rays(C.addressof(s), C.addressof(c), mode, seed, n, rays.ctypes.data, seeds.ctypes.data) == 0
    return rays, seeds


@pytest.fixture(scope="session")
def checker():
    lib = load_harness()

    def run(sc, cam, mode, seed, n_rays):
        s, keep = sc.as_c()
        c = camera_c(cam)
        res = Result()
        assert lib.scan_check(C.addressof(s), C.addressof(c), mode, seed, n_rays, C.addressof(res)) == 0
        return res
    return run


MODES = {0: "camera", 1: "scattered", 2: "grazing", 3: "far origin", 4: "degenerate"}


def _check(res, what):
    assert res.mismatches == 0, (what, "ray", list(res.bad_ray), "culled", res.t_cull, res.id_cull, "brute", res.t_brute,
                                 res.id_brute)


@pytest.mark.parametrize("mode", sorted(MODES))
def test_config4_mesh_culling_is_invisible(checker, mode):
    """The 10 002-triangle mesh of BASELINE config 4: three box levels and a grazing index."""
    sc, cam = scenes.c4_mesh()
    res = checker(sc, cam, mode, 11, 60000)
    _check(res, ("c4", MODES[mode]))
    assert res.n_trees == 2 and res.tree_levels == 3 and res.tree_leaves == (10002 + 7) // 8
    if mode < 2:
        assert res.hits > res.rays // 2
        # the point of it all: a ray looks at a small fraction of the 10 002 triangles
        assert res.triangle_tests * 40 < res.brute_triangle_tests


@pytest.mark.parametrize("name", ["triangle_mesh", "shapes", "ties", "media", "cornell", "rtiow", "moving", "motion_blur"])
def test_scene_culling_is_invisible(checker, name):
    sc, cam = getattr(scenes, name)(4 / 3)
    res = checker(sc, cam, -1, 5, 150000)
    _check(res, name)
    assert res.hits > 0


@pytest.mark.parametrize("seed", range(6))
def test_random_flat_scene_culling_is_invisible(checker, seed):
    """Random soups of rectangles (all three axes), triangles (slivers and degenerate ones included), boxes and
    spheres, on both sides of constant media, at coordinates from 1e-3 to 1e4."""
    rs = np.random.RandomState(100 + seed)
    scale = float(10.0 ** rs.uniform(-3, 4))
    s = scenes.Scene()
    mats = [s.lambertian((0.5, 0.5, 0.5)), s.metal((0.8, 0.8, 0.8), 0.1), s.dielectric(1.5)]
    n = 150 + 120 * seed
    for i in range(n):
        k = rs.randint(0, 10)
        c = (rs.uniform(-1, 1, 3) * scale).astype(np.float32)
        m = mats[rs.randint(0, 3)]
        if k < 5:
            e = rs.uniform(-0.15, 0.15, (2, 3)) * scale
            if k == 0:
                e[1] = e[0] * rs.uniform(-2, 2) + rs.uniform(-1e-6, 1e-6, 3) * scale  # a sliver
            if i % 97 == 0:
                e[1] = 0  # degenerate
            s.triangle(c, (c + e[0]).astype(np.float32), (c + e[1]).astype(np.float32), m)
        elif k < 7:
            a = rs.uniform(0.01, 0.3, 2) * scale
            s.rect(c[0], c[0] + a[0], c[1], c[1] + a[1], c[2], m, axis=int(rs.randint(0, 3)))
        elif k < 9:
            s.box(c, (c + rs.uniform(0.01, 0.2, 3) * scale).astype(np.float32), m)
        else:
            s.sphere(c, 0.05 * scale, m)
        if i in (n // 3, 2 * n // 3) and seed % 2:
            s.medium_sphere(c, 0.3 * scale, 2.0 / scale, (0.9, 0.9, 0.9))
    cam = scenes.make_camera(tuple(np.float32(scale) * np.array([2.5, 1, 2], np.float32)), (0, 0, 0), (0, 1, 0), 50.0, 4 / 3,
                             0.01 * scale, 3.0 * scale)
    res = checker(s, cam, -1, seed, 150000)
    _check(res, ("random flats", seed, scale))
    assert res.n_trees >= 2 and res.hits > 0
