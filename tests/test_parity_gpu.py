"""GPU parity tests: the sm_100a kernel, called through the C-ABI (libptb200.so), against the CPU oracle
on the same inputs, and against the reference's own committed outputs (tests/golden/ref_renders.npz).

Tolerance (BASELINE.json north_star): per-pixel mean-abs-error <= 1e-3 on the linear framebuffer and
PSNR >= 50 dB (peak 1.0).  The kernel is held to far more: + - * / sqrt are single IEEE-754 operations in the
reference's order and sin / cos / log / pow / asin / atan2 are glibc's own algorithms operation by operation
(path_tracer_b200/csrc/pt_glibc_math.cuh, pinned against the running libm on every float by
tests/test_glibc_math.py), so these tests require EVERY pixel to be BIT-identical to the oracle's (NaN for NaN)
and the closest-hit scan counts to be EXACTLY equal.
"""
import os

import numpy as np
import pytest

import scenes
from oracle.pyoracle import compare
from path_tracer_b200 import abi
from path_tracer_b200 import render as R

pytestmark = pytest.mark.gpu

MAE_TOL = 1e-3
PSNR_TOL = 50.0
SAME_TOL = 1.0  # the fraction of bit-identical pixels


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def assert_parity(got, want, what, same_tol=SAME_TOL):
    assert np.isfinite(want).all() == np.isfinite(got).all(), what
    nan = np.isnan(want) & np.isnan(got)  # (a pixel that is NaN in the reference has to be NaN here: compare the rest)
    mae, psnr, same = compare(np.where(nan, 0.0, got), np.where(nan, 0.0, want))
    assert mae <= MAE_TOL and psnr >= PSNR_TOL and same >= same_tol, (what, mae, psnr, same)
    return mae, psnr, same


def test_extension_is_loaded_and_gpu_visible():
    assert R.device_count() >= 1
    assert R.lib().pt_abi_version() == abi.PT_ABI_VERSION
    tf, mhz = R.measure_fp32_peak(0)
    assert 30.0 < tf < 90.0 and 900 < mhz < 2200


@pytest.mark.parametrize("name", list(scenes.ALL))
@pytest.mark.parametrize("cfg", [(64, 48, 8, 50), (33, 17, 5, 7), (40, 30, 16, 1)])
def test_scene_parity_vs_oracle(cport, name, cfg):
    w, h, spp, d = cfg
    sc, cam = scenes.ALL[name](w / h)
    got = R.render(sc, cam, w, h, spp, d)
    st = R.stats()
    want, cnt = cport.render(sc, cam, w, h, spp, d)
    assert_parity(got, want, (name, cfg))
    assert st["scans"] == cnt.scans and st["paths"] == cnt.paths == w * h * spp


@pytest.mark.parametrize("seed", range(8))
def test_random_scene_parity_vs_oracle(cport, seed):
    sc, cam = scenes.random_scene(seed, n_objects=40 + 15 * seed, aspect=64 / 48)
    got = R.render(sc, cam, 64, 48, 8, 50)
    st = R.stats()
    want, cnt = cport.render(sc, cam, 64, 48, 8, 50)
    assert_parity(got, want, ("random", seed))
    assert st["scans"] == cnt.scans


def test_parity_vs_reference_goldens(golden_renders, c1):
    """Against framebuffers produced by the UNMODIFIED reference (libptref.so), committed as fixtures."""
    c1_scene, c1_cam, _ = c1
    n = 0
    for key in golden_renders.files:
        name, cfg = key.rsplit("_", 1)
        w, h, spp, d = (int(v) for v in cfg.split("x"))
        if name == "c1":
            sc, cam = c1_scene, c1_cam
        elif name.startswith("random"):
            sc, cam = scenes.random_scene(int(name[6:]), aspect=w / h)
        else:
            sc, cam = scenes.ALL[name](w / h)
        assert_parity(R.render(sc, cam, w, h, spp, d), golden_renders[key], key)
        n += 1
    assert n >= 30


def test_c1_default_scene_full_size_sampled_rows(cport, c1):
    """BASELINE config 1 at full size and full spp (800x480, 100 spp, depth 50): every 24th row vs the oracle."""
    sc, cam, (w, h, spp, d) = c1
    got = R.render(sc, cam, w, h, spp, d)
    assert np.isfinite(got).all()
    rows = abi.pt_region(0, 5, w, (h - 5 + 23) // 24, 24)
    want, _ = cport.render_region(sc, cam, w, h, spp, d, rows)
    assert_parity(got[5::24], want, "c1 full-size rows")
    # 8-bit means of the reference's 100-spp render: R,G,B = 143.39/158.42/167.95,
    # through main.cpp:41-49's tone map
    img8 = (256 * np.clip(np.sqrt(got), 0.0, 0.999)).astype(np.int32)
    assert np.allclose(img8.reshape(-1, 3).mean(0), [143.39, 158.42, 167.95], atol=0.02)


def test_c1_partition_invariance_and_determinism(c1):
    """Size-independent properties at the full BASELINE size: any partition is BIT-identical to the whole."""
    sc, cam, (w, h, _, d) = c1
    spp = 10
    full = R.render(sc, cam, w, h, spp, d)
    again = R.render(sc, cam, w, h, spp, d)
    assert np.array_equal(_bits(full), _bits(again))
    for n in (2, 8):
        for r in (0, n - 1):
            part = R.render_region(sc, cam, w, h, spp, d, R.rows_region(w, h, r, n))
            assert np.array_equal(_bits(part), _bits(full[r::n])), (n, r)
    tile = abi.pt_region(123, 77, 200, 90, 1)
    assert np.array_equal(_bits(R.render_region(sc, cam, w, h, spp, d, tile)), _bits(full[77:167, 123:323]))


@pytest.mark.parametrize("name", ["media", "ties", "shapes", "moving"])
def test_team_size_invariance(cport, name):
    """Lanes-per-pixel (team size) is a scheduling choice: every power of two gives the same bits."""
    sc, cam = scenes.ALL[name](64 / 48)
    L = R.lib()
    try:
        base = None
        for t in (1, 2, 4, 8, 16, 32):
            L.pt_debug_set_team_size(t)
            img = R.render(sc, cam, 64, 48, 6, 50)
            if base is None:
                base = img
                want, cnt = cport.render(sc, cam, 64, 48, 6, 50)
                assert_parity(img, want, (name, t))
                assert R.stats()["scans"] == cnt.scans
            else:
                assert np.array_equal(_bits(img), _bits(base)), (name, t)
    finally:
        L.pt_debug_set_team_size(0)


@pytest.mark.parametrize("name", ["media", "shapes", "moving", "c1"])
def test_lane_kernel_equals_wavefront_kernel(name):
    """The two schedulers (wavefront = default, lane = pixel-per-lane-team) share the device functions and
    must produce the same bits."""
    if name == "c1":
        sc, cam, _ = scenes.load_c1()
        w, h, spp = 160, 96, 8
    else:
        sc, cam = scenes.ALL[name](64 / 48)
        w, h, spp = 64, 48, 6
    L = R.lib()
    wave = R.render(sc, cam, w, h, spp, 50)
    wave_scans = R.stats()["scans"]
    try:
        L.pt_debug_set_kernel(1)
        lane_img = R.render(sc, cam, w, h, spp, 50)
        assert R.stats()["scans"] == wave_scans
    finally:
        L.pt_debug_set_kernel(0)
    assert np.array_equal(_bits(wave), _bits(lane_img))


def test_rtiow_config2_layout(cport):
    """BASELINE config 2 layout (about 480 static spheres) at reduced image size, full depth."""
    sc, cam = scenes.rtiow(16 / 9)
    got = R.render(sc, cam, 160, 90, 16, 50)
    st = R.stats()
    want, cnt = cport.render(sc, cam, 160, 90, 16, 50)
    assert_parity(got, want, "rtiow")
    assert st["scans"] == cnt.scans


def test_config2_rtiow_full_size(cport):
    """BASELINE config 2 at full size: RTIOW random spheres, 1920x1080, 64 spp, depth 50; every 60th row
    against the oracle at full spp (133 M paths on the GPU, 2 M on the CPU)."""
    w, h, spp, d = 1920, 1080, 64, 50
    sc, cam = scenes.rtiow(w / h)
    got = R.render(sc, cam, w, h, spp, d)
    assert np.isfinite(got).all()
    rows = abi.pt_region(0, 11, w, (h - 11 + 59) // 60, 60)
    want, _ = cport.render_region(sc, cam, w, h, spp, d, rows)
    assert_parity(got[11::60], want, "config 2 rows")


def test_config3_cornell_tile_at_full_spp(cport):
    """BASELINE config 3 (Cornell box with smoke boxes and a light, 1024x1024, 1024 spp): a tile of the
    full-size image at FULL spp with true global seeds -- never reduced spp, per-pixel sums differ."""
    w, h, spp, d = 1024, 1024, 1024, 50
    sc, cam = scenes.cornell(1.0)
    tile = abi.pt_region(480, 300, 48, 16, 1)
    got = R.render_region(sc, cam, w, h, spp, d, tile)
    scans = R.stats()["scans"]
    want, cnt = cport.render_region(sc, cam, w, h, spp, d, tile)
    assert_parity(got, want, "config 3 tile")
    # 2.4 M constant_medium hits, each drawing log(rng) (constant_medium.hpp:65) and scattering by sin / cos
    # (rtweekend.hpp:70-80): glibc's logf / sinf / cosf, restated bit for bit
    assert scans == cnt.scans and cnt.as_dict()["accepts"][abi.HIT_MEDIUM] > 0


def test_config5_motion_blur_tile_at_full_spp(cport):
    """BASELINE config 5 (4K, moving spheres, depth of field, 4096 spp): a tile of the 3840x2160 image at
    full spp, plus the multi-GPU partition property on it (rows of the tile from another rank's region)."""
    w, h, spp, d = 3840, 2160, 4096, 50
    sc, cam = scenes.motion_blur(w / h)
    tile = abi.pt_region(1900, 700, 32, 8, 1)
    got = R.render_region(sc, cam, w, h, spp, d, tile)
    scans = R.stats()["scans"]
    want, cnt = cport.render_region(sc, cam, w, h, spp, d, tile)
    assert_parity(got, want, "config 5 tile")
    assert scans == cnt.scans
    # the same pixels rendered as part of rank 3's rows of an 8-GPU interleave are bit-identical
    rows = abi.pt_region(1900, 700 + 3, 32, 1, 8)
    part = R.render_region(sc, cam, w, h, spp, d, rows)
    assert np.array_equal(_bits(part[0]), _bits(got[3]))


def test_config5_nan_pixel_is_the_references(cport):
    """The full 3840x2160x4096 frame of BASELINE config 5 has exactly one non-finite pixel, (903, 1336)
    (tools/nonfinite_pixels.py c5), and it is the REFERENCE's: some sample's scatter direction degenerates
    and the NaN flows into the pixel's sum (render.hpp:95-104).  Parity means reproducing it -- NaN where the
    reference has NaN, identical bits around it -- not avoiding it."""
    w, h, spp, d = 3840, 2160, 4096, 50
    sc, cam = scenes.motion_blur(w / h)
    tile = abi.pt_region(902, 1335, 3, 3, 1)
    got = R.render_region(sc, cam, w, h, spp, d, tile)
    scans = R.stats()["scans"]
    want, cnt = cport.render_region(sc, cam, w, h, spp, d, tile)
    assert np.isnan(want[1, 1]).all() and np.isnan(got[1, 1]).all()
    finite = np.isfinite(want).all(axis=2)
    assert finite.sum() == 8 and np.array_equal(_bits(got[finite]), _bits(want[finite]))
    assert scans == cnt.scans


def test_scene_larger_than_shared_memory(cport):
    """A scan blob beyond the 227 KB shared-memory budget is streamed from L2 instead of staged."""
    sc, cam = scenes.triangle_mesh(16 / 9, nx=40, nz=32)  # 5 122 triangles x 48 B = 246 KB
    assert len(sc.arrays()["triangles"]) * 48 > 232448
    got = R.render(sc, cam, 48, 27, 2, 50)
    st = R.stats()
    want, cnt = cport.render(sc, cam, 48, 27, 2, 50)
    assert_parity(got, want, "big triangle mesh")
    assert st["scans"] == cnt.scans


def test_edge_cases(cport):
    sc, cam = scenes.spheres_basic()
    # depth 0: no bounce allowed, every sample is black (render.hpp:58,91)
    assert not R.render(sc, cam, 16, 12, 3, 0).any()
    # 1x1 image = pixel (0,0) = seed 0 = an all-zero RNG stream (xorshift.hpp:56)
    one = R.render(sc, cam, 1, 1, 7, 50)
    want, _ = cport.render(sc, cam, 1, 1, 7, 50)
    assert np.array_equal(_bits(one), _bits(want))
    # spp 1 / depth 1 / ragged sizes
    for (w, h, spp, d) in [(31, 9, 1, 50), (7, 33, 2, 1), (130, 3, 3, 2)]:
        got = R.render(sc, cam, w, h, spp, d)
        want, _ = cport.render(sc, cam, w, h, spp, d)
        assert_parity(got, want, (w, h, spp, d))
    # empty region is a no-op
    out = R.render_region(sc, cam, 16, 12, 1, 50, abi.pt_region(0, 0, 0, 0, 1))
    assert out.size == 0


def test_error_paths():
    sc, cam = scenes.spheres_basic()
    with pytest.raises(R.PathTracerError) as e:
        R.render_region(sc, cam, 16, 12, 1, 50, abi.pt_region(10, 0, 10, 1, 1))
    assert e.value.code == abi.PT_ERR_INVALID_ARGUMENT
    bad = scenes.spheres_basic()[0]
    bad._lists["spheres"][0]["material"] = 10 ** 6
    bad._frozen = None
    with pytest.raises(R.PathTracerError) as e:
        R.render(bad, cam, 8, 8, 1, 50)
    assert e.value.code == abi.PT_ERR_INVALID_ARGUMENT


def test_device_resident_api_matches_host_api(c1):
    import torch
    sc, cam, (w, h, _, d) = c1
    ds = R.DeviceScene(sc, 0)
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    ds.render_region(cam, w, h, 4, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3,
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    paths, scans = ds.counters()
    assert paths == w * h * 4 and scans > paths
    host = R.render(sc, cam, w, h, 4, d)
    assert np.array_equal(_bits(fb.cpu().numpy()), _bits(host))
    ds.close()


def test_concurrent_scenes_on_two_streams(c1):
    """Two device scenes launched back to back on two streams of one GPU (pt_abi.h, "Concurrency"): the CTAs of the
    render kernel wait for each other, so two half-resident grids would hang until the watchdog; the cooperative launch
    makes the runtime order them.  Both frames must be complete and bit-identical to a render on its own."""
    import torch
    sc, cam, (w, h, _, d) = c1
    sc2, cam2 = scenes.media(w / h)
    alone = [R.render(sc, cam, w, h, 6, d), R.render(sc2, cam2, w, h, 6, d)]
    a, b = R.DeviceScene(sc, 0), R.DeviceScene(sc2, 0)
    fbs = [torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0") for _ in range(2)]
    streams = [torch.cuda.Stream(device=0), torch.cuda.Stream(device=0)]
    torch.cuda.synchronize()
    for rep in range(3):
        for ds, c, fb, st in ((a, cam, fbs[0], streams[0]), (b, cam2, fbs[1], streams[1])):
            ds.render_region(c, w, h, 6, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, st.cuda_stream)
    torch.cuda.synchronize()
    for fb, want in zip(fbs, alone):
        assert np.array_equal(_bits(fb.cpu().numpy()), _bits(want))
    a.close(), b.close()


def test_single_process_multi_gpu(c1):
    if R.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc, cam, (w, h, _, d) = c1
    one = R.render(sc, cam, w, h, 4, d)
    R.set_num_gpus(2)
    try:
        two = R.render(sc, cam, w, h, 4, d)
    finally:
        R.set_num_gpus(1)
    assert np.array_equal(_bits(one), _bits(two))


# ---------------------------------------------------------------- chunk culling (pt_kernel.cu "CHUNK CULLING")
def _scene_by_name(name):
    if name == "c1":
        sc, cam, _ = scenes.load_c1()
        return sc, cam
    if name.startswith("random"):
        return scenes.random_scene(int(name[6:]), n_objects=90, aspect=4 / 3)
    return getattr(scenes, name)(4 / 3)


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("name", ["c1", "rtiow", "motion_blur", "moving", "media", "shapes", "random1", "random5"])
def test_chunk_culling_is_invisible(name, kernel):
    """Skipping the chunks whose box a ray misses is an optimisation only: the same bits with and without it, in both
    kernels (the wavefront kernel's (ray, chunk) items and the lane kernel's per-lane chunk lists)."""
    sc, cam = _scene_by_name(name)
    L = R.lib()
    try:
        L.pt_debug_set_kernel(kernel)
        L.pt_debug_set_cull(0)
        plain = R.render(sc, cam, 120, 90, 6, 50)
        scans_plain = R.stats()["scans"]
        L.pt_debug_set_cull(1)
        culled = R.render(sc, cam, 120, 90, 6, 50)
        assert np.array_equal(_bits(plain), _bits(culled)), name
        assert R.stats()["scans"] == scans_plain
    finally:
        L.pt_debug_set_cull(1)
        L.pt_debug_set_kernel(0)


@pytest.mark.parametrize("scale", [1.0, 6.0, 40.0, 300.0, 5000.0, 2.0e6])
def test_far_camera_uses_wider_box_sets(cport, scale):
    """Origins far from the scene are served by box sets with larger margins, and beyond the last bound by no culling at
    all (the reference's own discriminant is noise there): parity with the oracle at every distance."""
    sc, _ = scenes.rtiow(4 / 3, n=6)
    cam = scenes.make_camera(tuple(np.float32(scale) * np.array([13, 2, 3], np.float32)), (0, 0, 0), (0, 1, 0),
                             20.0 / scale, 4 / 3, 0.0, 10.0 * scale)
    got = R.render(sc, cam, 64, 48, 6, 50)
    st = R.stats()
    want, cnt = cport.render(sc, cam, 64, 48, 6, 50)
    assert_parity(got, want, scale)
    assert st["scans"] == cnt.scans


def test_chunk_boxes_follow_the_shutter(cport):
    """The boxes of moving spheres cover their sweep over the camera's shutter interval; a device scene recomputes them
    when a render brings another interval (also one that extrapolates the spheres' own time range)."""
    import torch
    sc, _ = scenes.motion_blur(4 / 3)
    ds = R.DeviceScene(sc, 0)
    w, h = 96, 72
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    try:
        for t0, t1 in [(0.0, 1.0), (0.25, 0.3), (-1.5, 3.0), (0.0, 1.0), (0.9, 0.1)]:
            cam = scenes.make_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, w / h, 0.1, 10.0, t0, t1)
            ds.render_region(cam, w, h, 4, 50, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3,
                             torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            want, _ = cport.render(sc, cam, w, h, 4, 50)
            assert_parity(fb.cpu().numpy(), want, (t0, t1))
    finally:
        ds.close()


def test_degenerate_directions_are_not_culled_wrongly(cport):
    """Axis-parallel rays (zero direction components: infinite slab reciprocals) through a grid of spheres."""
    s = scenes.Scene()
    for i in range(-8, 9):
        for j in range(-8, 9):
            s.sphere((float(i), float(j), 0.0), 0.45, s.lambertian((0.5 + 0.02 * i, 0.5, 0.5 + 0.02 * j)))
    # an orthographic-like camera: tiny field of view from far away, looking straight down the z axis
    cam = scenes.make_camera((0, 0, 50), (0, 0, 0), (0, 1, 0), 20.0, 1.0, 0.0, 50.0)
    got = R.render(s, cam, 65, 65, 4, 50)
    want, cnt = cport.render(s, cam, 65, 65, 4, 50)
    assert_parity(got, want, "grid")
    assert R.stats()["scans"] == cnt.scans


def test_item_list_overflow_is_scanned_in_place(cport):
    """A grazing camera makes every ray cross many chunk boxes, so a full CTA's (ray, chunk) items do not fit the item
    list; what does not fit is scanned where it was found.  Large enough that every CTA holds a full pool."""
    import ctypes as C
    import torch
    rs = np.random.RandomState(5)
    sc = scenes.Scene()
    mats = [sc.lambertian((0.7, 0.3, 0.3)), sc.metal((0.8, 0.8, 0.8), 0.1), sc.dielectric(1.5), sc.lightsource((2, 2, 2))]
    for k in range(48):  # 48 clusters of 16 small spheres in a row along z: a ray down the row crosses every cluster's box
        for i in range(16):
            c = np.array([rs.uniform(-1.5, 1.5), rs.uniform(-1.5, 1.5), 2.0 * k + rs.uniform(0, 1)], dtype=np.float32)
            if i % 4 == 0:
                sc.sphere(c, 0.08, mats[i % 3], center1=c + np.array([0.1, 0, 0], np.float32), time0=0.0, time1=1.0)
            else:
                sc.sphere(c, 0.08, mats[(i + k) % 4])
    cam = scenes.make_camera((0, 0, -12), (0, 0, 50), (0, 1, 0), 1.4, 16 / 9, 0.0, 10.0, 0.0, 1.0)
    w, h = 640, 360
    ds = R.DeviceScene(sc, 0)
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    try:
        ds.render_region(cam, w, h, 2, 50, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 11)()
        R.lib().pt_debug_timeline(ds._h, out)
        assert out[10] > 1000, "the item list never overflowed: the test does not test what it says"
        paths, scans = ds.counters()
    finally:
        ds.close()
    got = fb.cpu().numpy()
    # the lane kernel has no item list: same bits, same scan count
    L = R.lib()
    try:
        L.pt_debug_set_kernel(1)
        lane = R.render(sc, cam, w, h, 2, 50)
        assert np.array_equal(_bits(got), _bits(lane)) and R.stats()["scans"] == scans
    finally:
        L.pt_debug_set_kernel(0)
    want, cnt = cport.render(sc, cam, w, h, 2, 50)
    assert_parity(got, want, "row of clusters")
    assert scans == cnt.scans


def test_item_list_shares_follow_the_demand():
    """A scene with static spheres only asks for more (ray, chunk) items a round than the static share of the item list
    used to hold (3/8 of it, the default scene's optimum: config 2 scanned 0.084 items per closest-hit scan in place, one
    thread doing 16 spheres' work while its warp waits, and took 19 % longer).  The shares follow the demand now
    (pt_wave.cu, "The shares of the item list follow the demand"): next to nothing is scanned in place -- measured on
    this short frame, whose first rounds still start from the fixed shares: 0.014 items per scan, against 0.24 before."""
    import ctypes as C
    import torch
    w, h = 960, 540  # every CTA holds a full pool
    sc, cam = scenes.rtiow(w / h)
    ds = R.DeviceScene(sc, 0)
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    try:
        ds.render_region(cam, w, h, 8, 50, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 11)()
        R.lib().pt_debug_timeline(ds._h, out)
        paths, scans = ds.counters()
    finally:
        ds.close()
    assert paths == w * h * 8
    assert out[10] < 0.05 * scans, (out[10], scans)


def test_more_flat_objects_than_the_unit_table_holds(cport):
    """Flat objects beyond the (ray, object) unit table are scanned sequentially per ray."""
    sc, cam = scenes.triangle_mesh(4 / 3, nx=24, nz=10)
    assert len(sc.triangles) > 256
    got = R.render(sc, cam, 96, 72, 4, 50)
    st = R.stats()
    want, cnt = cport.render(sc, cam, 96, 72, 4, 50)
    assert_parity(got, want, "mesh")
    assert st["scans"] == cnt.scans


def test_scene_too_large_for_shared_memory(cport):
    """9 000 spheres: the scan blob streams from L2 instead of shared memory, and there are more chunk-box blocks than
    the short rounds' table holds (they fall back to one thread per ray)."""
    rs = np.random.RandomState(11)
    s = scenes.Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian((0.5, 0.5, 0.5)))
    mats = [s.lambertian((0.7, 0.3, 0.3)), s.metal((0.8, 0.8, 0.8), 0.1), s.dielectric(1.5), s.lambertian((0.2, 0.4, 0.8))]
    for i in range(9000):
        c = np.array([rs.uniform(-20, 20), rs.uniform(0.1, 3.0), rs.uniform(-20, 20)], dtype=np.float32)
        if i % 5 == 0:
            s.sphere(c, 0.1, mats[i % 4], center1=c + np.array([0, 0.2, 0], np.float32), time0=0.0, time1=1.0)
        else:
            s.sphere(c, 0.1, mats[i % 4])
    cam = scenes.make_camera((30, 6, 8), (0, 1, 0), (0, 1, 0), 30.0, 4 / 3, 0.0, 10.0, 0.0, 1.0)
    for w, h, spp in ((64, 48, 3), (400, 300, 1)):
        got = R.render(s, cam, w, h, spp, 50)
        st = R.stats()
        want, cnt = cport.render(s, cam, w, h, spp, 50)
        assert_parity(got, want, ("large", w, h))
        assert st["scans"] == cnt.scans


def test_negative_radius_and_nested_spheres(cport):
    """RTIOW's hollow glass bubble (a sphere of negative radius inside a glass sphere: same geometry, flipped normal,
    sphere.hpp:81) and concentric shells around it -- chunk boxes are built from |radius|."""
    s = scenes.Scene()
    s.sphere((0, -100.5, -1), 100, s.lambertian((0.8, 0.8, 0.0)))
    s.sphere((0, 0, -1), 0.5, s.dielectric(1.5))
    s.sphere((0, 0, -1), -0.45, s.dielectric(1.5))
    s.sphere((0, 0, -1), 0.2, s.lambertian((0.1, 0.2, 0.5)))
    s.sphere((-1, 0, -1), 0.5, s.metal((0.8, 0.6, 0.2), 0.0))
    s.sphere((1, 0, -1), -0.5, s.metal((0.8, 0.8, 0.8), 0.3))
    for i in range(40):
        s.sphere((-2 + 0.1 * i, -0.4, -0.3 - 0.02 * i), 0.05 if i % 2 else -0.05, s.lambertian((0.5, 0.1 + 0.02 * i, 0.3)))
    cam = scenes.make_camera((-2, 2, 1), (0, 0, -1), (0, 1, 0), 35.0, 4 / 3, 0.0, 3.4)
    got = R.render(s, cam, 160, 120, 8, 50)
    st = R.stats()
    want, cnt = cport.render(s, cam, 160, 120, 8, 50)
    assert_parity(got, want, "hollow glass")
    assert st["scans"] == cnt.scans


# ---------------------------------------------------------------- flat culling (pt_prims.cuh "FLAT CULLING", BASELINE config 4)
def _flat_soup(seed, n, with_media):
    """Rectangles (all three axes), triangles, boxes and a few spheres in random order -- enough of each kind for a
    tree per kind -- optionally on both sides of constant media."""
    rs = np.random.RandomState(seed)
    s = scenes.Scene()
    mats = [s.lambertian((0.6, 0.5, 0.4)), s.metal((0.8, 0.8, 0.8), 0.1), s.dielectric(1.5), s.lambertian(s.checker((0.1, 0.1, 0.1), (0.9, 0.9, 0.9))),
            s.lightsource((3, 3, 3))]
    s.sphere((0, -1000, 0), 1000, mats[0])
    for i in range(n):
        k = rs.randint(0, 10)
        c = np.array([rs.uniform(-5, 5), rs.uniform(0.1, 3.0), rs.uniform(-5, 5)], dtype=np.float32)
        m = mats[rs.randint(0, 5) if i % 9 == 0 else rs.randint(0, 4)]
        if k < 5:
            e = rs.uniform(-0.6, 0.6, (2, 3))
            s.triangle(c, (c + e[0]).astype(np.float32), (c + e[1]).astype(np.float32), m)
        elif k < 7:
            a = rs.uniform(0.1, 0.9, 2)
            s.rect(c[0], c[0] + a[0], c[1], c[1] + a[1], c[2], m, axis=int(rs.randint(0, 3)) if with_media else abi.AXIS_XY)
        elif k < 9:
            s.box(c, (c + rs.uniform(0.1, 0.7, 3)).astype(np.float32), m)
        else:
            s.sphere(c, 0.25, m)
        if with_media and i in (n // 3, 2 * n // 3):
            s.medium_sphere(c, 1.2, 0.8, (0.9, 0.9, 0.9))
    cam = scenes.make_camera((11, 4, 9), (0, 1, 0), (0, 1, 0), 35.0, 4 / 3, 0.05, 14.0)
    return s, cam


@pytest.mark.parametrize("with_media", [False, True])
def test_flat_trees_parity_vs_oracle(cport, with_media):
    """Trees over rectangles, triangles and boxes; with media the trees behind the first medium are walked
    sequentially per ray (LATE), and spheres behind a medium force the sequential scan of the whole list."""
    sc, cam = _flat_soup(21 + with_media, 400, with_media)
    for (w, h, spp) in ((96, 72, 6), (20, 12, 9)):  # (the small one runs in short rounds from the start)
        got = R.render(sc, cam, w, h, spp, 50)
        st = R.stats()
        want, cnt = cport.render(sc, cam, w, h, spp, 50)
        assert_parity(got, want, ("flat soup", with_media, w, h))
        assert st["scans"] == cnt.scans


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("name", ["mesh", "soup", "soup_media"])
def test_flat_culling_is_invisible(name, kernel):
    """Box trees and the grazing index are an optimisation only: the same bits with and without them, both kernels."""
    if name == "mesh":
        sc, cam = scenes.c4_mesh(4 / 3, nx=30, nz=12)
    else:
        sc, cam = _flat_soup(5, 300, name == "soup_media")
    L = R.lib()
    try:
        L.pt_debug_set_kernel(kernel)
        L.pt_debug_set_cull(0)
        plain = R.render(sc, cam, 120, 90, 4, 50)
        scans_plain = R.stats()["scans"]
        L.pt_debug_set_cull(1)
        culled = R.render(sc, cam, 120, 90, 4, 50)
        assert np.array_equal(_bits(plain), _bits(culled)), name
        assert R.stats()["scans"] == scans_plain
    finally:
        L.pt_debug_set_cull(1)
        L.pt_debug_set_kernel(0)


def test_config4_mesh_tiles_at_full_spp(cport):
    """BASELINE config 4 (10 002 triangles, checker + image textures, 1920x1080, 256 spp, depth 50): tiles of the
    full-size image at FULL spp with true global seeds against the oracle's brute-force scan."""
    w, h, spp, d = 1920, 1080, 256, 50
    sc, cam = scenes.c4_mesh(w / h)
    assert len(sc.arrays()["triangles"]) == 10002
    for (x0, y0) in ((950, 420), (300, 560), (1500, 250)):
        tile = abi.pt_region(x0, y0, 16, 6, 1)
        got = R.render_region(sc, cam, w, h, spp, d, tile)
        scans = R.stats()["scans"]
        want, cnt = cport.render_region(sc, cam, w, h, spp, d, tile)
        assert_parity(got, want, ("config 4 tile", x0, y0))
        assert scans == cnt.scans


def test_single_task_mode_is_the_references(cport):
    """pt_render_single_task = the reference built with -DUSE_SINGLE_TASK (render.hpp:113-122): one generator with the
    default seed for the whole image, pixels x-major.  Against the oracle's restatement (itself bit-identical to
    oracle/_ref/libptref_st.so) and the committed hashes of the reference's own output."""
    import hashlib
    import json
    hashes = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_hashes.json")))
    for name, w, h, spp in (("c1", 64, 48, 4), ("shapes", 64, 48, 4), ("media", 64, 48, 4)):
        if name == "c1":
            sc, cam, _ = scenes.load_c1()
        else:
            sc, cam = scenes.ALL[name](w / h)
        got = R.render_single_task(sc, cam, w, h, spp, 50)
        want, cnt = cport.render_single_task(sc, cam, w, h, spp, 50)
        assert np.array_equal(_bits(got), _bits(want)), name
        assert R.stats()["scans"] == cnt.scans
        assert hashlib.sha256(got.tobytes()).hexdigest() == hashes["single_task_%s_%dx%dx%dx50" % (name, w, h, spp)]
    # a mesh with trees, and not the per-pixel-seeded image
    sc, cam = scenes.triangle_mesh(16 / 9, nx=12, nz=6)
    got = R.render_single_task(sc, cam, 48, 27, 2, 50)
    want, _ = cport.render_single_task(sc, cam, 48, 27, 2, 50)
    assert np.array_equal(_bits(got), _bits(want))
    assert not np.array_equal(_bits(got), _bits(R.render(sc, cam, 48, 27, 2, 50)))


def test_tree_lists_spill_to_global_memory(cport):
    """A frame with a full ray pool over the mesh: the (ray, node) / (ray, leaf) items of a round do not fit the 32 KB of
    shared-memory lists and continue in global memory -- and nothing is walked in place.  Rows against the oracle."""
    import ctypes as C
    import torch
    w, h, spp, d = 640, 360, 4, 50
    sc, cam = scenes.c4_mesh(w / h)
    ds = R.DeviceScene(sc, 0)
    fb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    ds.render_region(cam, w, h, spp, d, R.rows_region(w, h, 0, 1), fb.data_ptr(), w * 3, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    L = R.lib()
    spilled = C.c_ulonglong(0)
    L.pt_debug_tree_spilled.argtypes = [C.c_void_p, C.c_void_p]
    assert L.pt_debug_tree_spilled(ds._h, C.byref(spilled)) == 0
    tl = (C.c_ulonglong * 11)()
    L.pt_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
    assert L.pt_debug_timeline(ds._h, tl) == 0
    assert spilled.value > 100_000 and tl[10] == 0, (spilled.value, tl[10])
    got = fb.cpu().numpy()
    ds.close()
    rows = R.rows_region(w, h, 5, 45)  # rows 5, 50, 95, ...
    want, cnt = cport.render_region(sc, cam, w, h, spp, d, rows)
    assert np.array_equal(_bits(got[5::45]), _bits(want))


def test_config4_culling_pays(cport):
    """Config 4 at reduced size: the trees must make the frame at least 10x faster than testing every triangle."""
    w, h, spp = 480, 270, 2
    sc, cam = scenes.c4_mesh(w / h)
    L = R.lib()
    R.render(sc, cam, w, h, spp, 50)  # warm-up
    culled = R.render(sc, cam, w, h, spp, 50)
    ms_culled, scans = R.stats()["kernel_ms"], R.stats()["scans"]
    try:
        L.pt_debug_set_cull(0)
        plain = R.render(sc, cam, w, h, spp, 50)
        ms_plain = R.stats()["kernel_ms"]
        assert R.stats()["scans"] == scans
    finally:
        L.pt_debug_set_cull(1)
    assert np.array_equal(_bits(plain), _bits(culled))
    print("config 4 at %dx%dx%d: %.2f ms with trees, %.2f ms without" % (w, h, spp, ms_culled, ms_plain))
    assert ms_plain >= 10.0 * ms_culled


# ---------------------------------------------------------------- progressive rendering
@pytest.mark.parametrize("kernel", [0, 1])
def test_resume_is_bit_identical_to_one_launch(c1, kernel):
    """render.hpp:94-105 keeps {RNG, sum} per pixel across its sample loop: samples 0..40 followed by 40..100 from the
    saved state must leave exactly the framebuffer (and the state) of one launch 0..100."""
    sc, cam, (w, h, _, d) = c1
    region = abi.pt_region(300, 200, 96, 40, 1)
    L = R.lib()
    try:
        L.pt_debug_set_kernel(kernel)
        whole = R.render_region(sc, cam, w, h, 100, d, region)
        fb_a, st_a = R.render_resume(sc, cam, w, h, 0, 40, d, region)
        assert np.array_equal(_bits(fb_a), _bits(R.render_region(sc, cam, w, h, 40, d, region)))
        fb_b, st_b = R.render_resume(sc, cam, w, h, 40, 100, d, region, st_a)
        assert np.array_equal(_bits(fb_b), _bits(whole))
        fb_c, st_c = R.render_resume(sc, cam, w, h, 0, 100, d, region)
        assert np.array_equal(_bits(st_c), _bits(st_b)) and np.array_equal(_bits(fb_c), _bits(whole))
        # three uneven steps, the LPT order on in one of them (>= 32768 pixels and >= 8 samples) and off in the others
        big = abi.pt_region(0, 100, 800, 48, 1)
        one = R.render_region(sc, cam, w, h, 12, d, big)
        fb, st = R.render_resume(sc, cam, w, h, 0, 1, d, big)
        fb, st = R.render_resume(sc, cam, w, h, 1, 10, d, big, st)
        fb, st = R.render_resume(sc, cam, w, h, 10, 12, d, big, st)
        assert np.array_equal(_bits(fb), _bits(one))
    finally:
        L.pt_debug_set_kernel(0)
    with pytest.raises(R.PathTracerError):
        R.render_resume(sc, cam, w, h, 5, 5, d, region, st_a)


def test_staged_framebuffer_stores_are_invisible(c1):
    """Finished pixels go through a float4 staging image and a coalescing resolve pass (16-byte stores); with the stage
    off they are written straight into the caller's rows, three scalars each: same bits, both kernels, also for a tile
    with a row pitch and for widths that are not a multiple of 4 (which always take the scalar path)."""
    sc, cam, (w, h, _, d) = c1
    L = R.lib()
    try:
        for kernel in (0, 1):
            L.pt_debug_set_kernel(kernel)
            for region in (abi.pt_region(0, 0, w, 64, 1), abi.pt_region(40, 100, 128, 32, 3), abi.pt_region(8, 8, 50, 20, 1)):
                L.pt_debug_set_fb_stage(1)
                a = R.render_region(sc, cam, w, h, 6, d, region)
                launches = R.stats()["kernel_launches"]
                L.pt_debug_set_fb_stage(0)
                b = R.render_region(sc, cam, w, h, 6, d, region)
                assert np.array_equal(_bits(a), _bits(b)), (kernel, region.w)
                assert launches == R.stats()["kernel_launches"] + (1 if region.w % 4 == 0 else 0)
    finally:
        L.pt_debug_set_fb_stage(1)
        L.pt_debug_set_kernel(0)
