"""GPU, ray level: the CUDA closest-hit scan (chunk boxes, flat trees, grazing index, vector-order fallback) against the
oracle's hit_world (render.hpp:30-51) on the adversarial rays of tests/host/scan_check.cpp -- camera rays, scattered
rays, rays grazing object planes, rays aimed at vertices from far away, axis-parallel / denormal / NaN / infinite rays.
Per ray: the same object, the same t (bit for bit; NaN == NaN), the same generator state afterwards."""
import numpy as np
import pytest

import scenes
from path_tracer_b200 import abi
from path_tracer_b200 import render as R
from test_flat_culling import make_rays
from test_parity_gpu import _flat_soup

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def _compare(sc, cam, cport, mode, seed, n, what):
    rays, seeds = make_rays(sc, cam, mode, seed, n)
    want_t, want_i, want_rng = cport.hit_world_batch(sc, rays, seeds)
    ds = R.DeviceScene(sc, 0)
    try:
        got_t, got_i, got_rng = ds.closest_hit(cam, rays, seeds, 0)
        ord_t, ord_i, ord_rng = ds.closest_hit(cam, rays, seeds, 1)
    finally:
        ds.close()
    kinds = np.array([int(e["kind"]) for e in sc.arrays()["order"]] + [-1])
    medium = (kinds[want_i] == abi.HIT_MEDIUM) | (kinds[got_i] == abi.HIT_MEDIUM)
    for name, (t, i, rng) in (("culled", (got_t, got_i, got_rng)), ("vector order", (ord_t, ord_i, ord_rng))):
        same_t = (_bits(t) == _bits(want_t)) | (np.isnan(t) & np.isnan(want_t))
        exact = same_t & (i == want_i) & (rng == want_rng)
        bad = ~exact & ~medium
        assert not bad.any(), (what, name, int(bad.sum()), rays[bad][:3].tolist(), t[bad][:3], want_t[bad][:3], i[bad][:3], want_i[bad][:3])
        # (a constant_medium's t goes through log(): glibc's logf, restated bit for bit in pt_glibc_math.cuh)
        assert exact.all(), (what, name, "medium", int((~exact).sum()))
    return int((want_i >= 0).sum())


@pytest.mark.parametrize("name", ["shapes", "ties", "rect_axes", "triangle_mesh", "media", "cornell", "moving", "rtiow"])
def test_ray_level_parity(cport, name):
    sc, cam = scenes.ALL[name](4 / 3) if name in scenes.ALL else getattr(scenes, name)(4 / 3)
    hits = _compare(sc, cam, cport, -1, 3, 100000, name)
    assert hits > 1000


def test_ray_level_parity_mesh(cport):
    """The 10 002-triangle mesh: three box levels and the grazing index, every ray mode."""
    sc, cam = scenes.c4_mesh()
    for mode in range(5):
        _compare(sc, cam, cport, mode, 7 + mode, 40000, ("c4", mode))


@pytest.mark.parametrize("with_media", [False, True])
def test_ray_level_parity_flat_soup(cport, with_media):
    sc, cam = _flat_soup(31, 500, with_media)
    _compare(sc, cam, cport, -1, 5, 150000, ("soup", with_media))


def test_rectangle_plane_nan_is_the_references(cport):
    """A ray IN the plane of a rectangle (d_k == 0, o_k == k) has t = 0/0 = NaN there, which the reference accepts and
    which poisons its running closest hit (rectangle.hpp:35-41): later spheres are rejected, later
    flat objects accepted whatever their t.  Pinned here on hand-made rays through scenes of all three rectangle axes."""
    s = scenes.Scene()
    m = s.lambertian((0.5, 0.5, 0.5))
    s.sphere((0, 0, -3), 1.0, m)                      # before the rectangle: hit normally
    s.rect(-1, 1, -1, 1, 0.5, m)                      # xy rectangle at z = 0.5
    s.sphere((0, 0, -6), 1.0, m)                      # after it: rejected once the closest hit is NaN
    s.triangle((-5, -5, -9), (5, -5, -9), (0, 5, -9), m)  # after it: accepted whatever its t
    s.rect(-1, 1, -1, 1, 2.0, m, axis=abi.AXIS_XZ)
    s.rect(-1, 1, -1, 1, -2.0, m, axis=abi.AXIS_YZ)
    s.box((3, 3, 3), (4, 4, 4), m)
    cam = scenes.make_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 5.0)
    rays = np.array([
        [0.2, 0.1, 0.5, 1.0, 0.3, 0.0, 0],     # in the plane z = 0.5 of the xy rectangle
        [9.0, 9.0, 0.5, -1.0, -1.0, 0.0, 0],   # the same plane, far outside the rectangle
        [0.3, 2.0, 0.1, 1.0, 0.0, -0.2, 0],    # in the plane y = 2 of the xz rectangle
        [-2.0, 0.3, 0.2, 0.0, 0.5, 1.0, 0],    # in the plane x = -2 of the yz rectangle
        [3.5, 3.5, 4.0, 0.4, -0.3, 0.0, 0],    # in the plane of the box's +z side
        [0.0, 0.0, 5.0, 0.0, 0.0, -1.0, 0],    # axis parallel, no NaN: through sphere, rectangle, sphere, triangle
        [0.2, 0.1, 0.5, 0.0, 0.0, 0.0, 0],     # a zero direction
    ], dtype=np.float32)
    seeds = np.arange(1, len(rays) + 1, dtype=np.uint32)
    want_t, want_i, _ = cport.hit_world_batch(s, rays, seeds)
    assert np.isnan(want_t[:5]).sum() >= 3, want_t  # the scenario really produces the reference's NaN hits
    ds = R.DeviceScene(s, 0)
    try:
        for mode in (0, 1):
            t, i, _ = ds.closest_hit(cam, rays, seeds, mode)
            assert np.array_equal(i, want_i), (mode, i, want_i)
            assert np.array_equal(np.isnan(t), np.isnan(want_t)) and np.array_equal(_bits(t)[~np.isnan(t)], _bits(want_t)[~np.isnan(t)])
    finally:
        ds.close()


def test_render_level_nan_scenario(cport):
    """The NaN quirk through the whole renderer.  Pixel (0,0) has seed 0, a generator that returns zeros forever, so its
    Lambertian bounces add unit_vec = (-1, 0, -0) to the normal (rtweekend.hpp:60-67): off an xy rectangle the scattered
    direction is (-1, 0, 1), d_y == 0 exactly.  An xz rectangle placed exactly at the height of that bounce point then has
    t = 0/0 = NaN for the scattered ray -- accepted by the reference, poisoning its scan.  The wavefront kernel sees such
    a ray when it is stored and runs that round through the per-ray sequential scan; the image must be the oracle's,
    NaN for NaN, in both kernels."""
    w, h, spp = 8, 6, 3

    def build(k_y=None):
        s = scenes.Scene()
        s.rect(-20, 20, -20, 20, -1.0, s.lambertian((0.8, 0.7, 0.6)))                      # the wall the camera looks at
        if k_y is not None:
            s.rect(-30, 30, -0.9, 4.0, k_y, s.lambertian((0.2, 0.9, 0.2)), axis=abi.AXIS_XZ)  # the plane of the bounce
        s.sphere((-3, 0, 2), 1.0, s.metal((0.9, 0.9, 0.9), 0.0))                          # listed after it: rejected by a NaN
        s.triangle((-9, -9, 3.5), (9, -9, 3.5), (0, 9, 3.5), s.lambertian((0.3, 0.3, 0.9)))  # listed after it: accepted
        return s

    cam = scenes.make_camera((0, 0, 3), (0, 0, -1), (0, 1, 0), 40.0, w / h, 0.0, 4.0)
    # the first bounce of pixel (0,0): u = v = 0 and a zero lens offset (seed 0)
    ray0 = np.array(list(cam["origin"]) + list(np.float32(cam["lower_left_corner"]) - np.float32(cam["origin"])) + [0.0], dtype=np.float32)
    out = cport.hit_scatter(build(), ray0, 0)
    assert out[0] == 1.0 and out[11] == 1.0 and out[19] == 0.0, out  # hit, scattered, direction.y == 0 exactly
    sc = build(k_y=float(out[16]))  # the xz rectangle exactly at the bounce point's height
    t, idx, _ = cport.hit_world_batch(sc, np.array([list(out[15:21]) + [0.0]], np.float32), np.array([0], np.uint32))
    assert np.isnan(t[0]), (t, idx)  # the reference's scan of the scattered ray is poisoned
    want, _ = cport.render(sc, cam, w, h, spp, 50)
    L = R.lib()
    try:
        for kernel in (0, 1):
            L.pt_debug_set_kernel(kernel)
            got = R.render(sc, cam, w, h, spp, 50)
            same = (_bits(got) == _bits(want)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (kernel, got[0, 0], want[0, 0], int((~same).sum()))
    finally:
        L.pt_debug_set_kernel(0)
