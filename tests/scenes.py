"""Scene builders for the parity tests (flat pt_scene via path_tracer_b200.Scene).

Every builder returns (scene, camera).  They exercise each primitive,
material and texture of the hot path plus the edge cases
the reference's semantics create: exact-tie overlaps, objects on both sides of
a constant_medium in list order, empty scenes, moving-sphere time classes.
"""
import os

import numpy as np

from path_tracer_b200 import Scene, abi, make_camera

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_c1():
    """The reference's default scene (main.cpp:67-183), captured from the unmodified main.cpp."""
    return Scene.load(os.path.join(GOLDEN, "c1_scene.ptsc.gz"))


def _cam(aspect, look_from=(13, 2, 3), look_at=(0, 0, 0), vfov=20.0, aperture=0.1, focus=10.0, t0=0.0, t1=0.0):
    return make_camera(look_from, look_at, (0, 1, 0), vfov, aspect, aperture, focus, t0, t1)


def test_image(w=16, h=8, seed=7):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)


def spheres_basic(aspect=4 / 3):
    s = Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian(s.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9))))
    s.sphere((0, 1, 0), 1.0, s.dielectric(1.5))
    s.sphere((-4, 1, 0), 1.0, s.lambertian((0.4, 0.2, 0.1)))
    s.sphere((4, 1, 0), 1.0, s.metal((0.7, 0.6, 0.5), 0.0))
    s.sphere((2, 0.5, 2), 0.5, s.metal((0.8, 0.8, 0.9), 0.3))
    s.sphere((1, 0.4, -2), 0.4, s.dielectric(1.5, (1.0, 0.5, 0.5)))
    s.sphere((3, 0.3, 1), -0.25 + 0.55, s.lightsource((4, 4, 4)))
    return s, _cam(aspect)


def rtiow(aspect=16 / 9, n=11, seed=1):
    """RTIOW-style random spheres (BASELINE config 2 layout), all static."""
    rs = np.random.RandomState(seed)
    f = np.float32
    s = Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian((0.5, 0.5, 0.5)))
    for a in range(-n, n):
        for b in range(-n, n):
            choose = rs.rand()
            c = (f(a + 0.9 * rs.rand()), f(0.2), f(b + 0.9 * rs.rand()))
            if np.linalg.norm(np.array(c) - np.array([4, 0.2, 0])) <= 0.9:
                continue
            if choose < 0.8:
                s.sphere(c, 0.2, s.lambertian(tuple(rs.rand(3) * rs.rand(3))))
            elif choose < 0.95:
                s.sphere(c, 0.2, s.metal(tuple(0.5 + 0.5 * rs.rand(3)), 0.5 * rs.rand()))
            else:
                s.sphere(c, 0.2, s.dielectric(1.5))
    s.sphere((0, 1, 0), 1.0, s.dielectric(1.5))
    s.sphere((-4, 1, 0), 1.0, s.lambertian((0.4, 0.2, 0.1)))
    s.sphere((4, 1, 0), 1.0, s.metal((0.7, 0.6, 0.5), 0.0))
    return s, _cam(aspect)


def moving(aspect=4 / 3):
    """Moving spheres in two (time0,time1) classes interleaved with static ones; shutter 0..1."""
    rs = np.random.RandomState(3)
    s = Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian((0.5, 0.5, 0.5)))
    for i in range(40):
        c = np.array([rs.uniform(-5, 5), 0.3, rs.uniform(-5, 5)], dtype=np.float32)
        m = s.lambertian(tuple(rs.rand(3))) if i % 3 else s.metal(tuple(0.5 + 0.5 * rs.rand(3)), 0.2)
        if i % 4 == 0:
            s.sphere(c, 0.3, m)
        elif i % 4 in (1, 2):
            s.sphere(c, 0.3, m, center1=c + np.array([0, rs.uniform(0, 0.5), 0], dtype=np.float32), time0=0.0, time1=1.0)
        else:
            s.sphere(c, 0.3, m, center1=c + np.array([rs.uniform(0, 0.3), 0, 0], dtype=np.float32), time0=0.25, time1=0.75)
    return s, _cam(aspect, t0=0.0, t1=1.0)


def shapes(aspect=4 / 3):
    """xy_rect, triangles, boxes with solid / checker / image textures (u,v only on sphere/rect/box)."""
    s = Scene()
    img = s.image(test_image())
    img5 = s.image_view(img, 5.0)
    s.sphere((0, -1000, 0), 1000, s.lambertian(s.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9))))
    s.rect(-2, 2, 0, 2, -2, s.lambertian(img))
    s.sphere((0, 1, 0), 1.0, s.lambertian(img5))
    s.triangle((3, 0, 1.3), (2.75, 1.5, 1.05), (3, 0, 0.8), s.lambertian((0.68, 0.5, 0.1)))
    s.triangle((2.5, 0, 1.3), (2.75, 1.5, 1.05), (3, 0, 1.3), s.metal((0.89, 0.73, 0.29), 0.1))
    s.triangle((3, 0, 0.8), (2.75, 1.5, 1.05), (2.5, 0, 0.8), s.lambertian((0, 0, 1)))
    s.triangle((2.5, 0, 0.8), (2.75, 1.5, 1.05), (2.5, 0, 1.3), s.dielectric(1.5))
    s.box((-3.5, 0, 1), (-2.5, 1.5, 2), s.metal((0.7, 0.6, 0.5), 0.25))
    s.box((1.5, 0, -1.5), (2.5, 0.8, -0.5), s.lambertian(img))
    s.sphere((1, 2.5, 1), 0.3, s.lightsource((10, 8, 6)))
    return s, _cam(aspect, look_from=(9, 3, 6), vfov=30.0, aperture=0.05, focus=11.0)


def rect_axes(aspect=4 / 3):
    """Top-level xz / yz rectangles: accepted by the C-ABI as a superset (the reference's
    hittable_t only holds xy_rect at top level, render.hpp:22-23), so oracle-port only."""
    s, cam = shapes(aspect)
    s.rect(-1, 1, -1, 1, 3.0, s.lightsource(s.checker((3, 0, 0), (0, 3, 0))), axis=abi.AXIS_XZ)
    s.rect(0, 1, 0, 1, -3.0, s.lambertian((0.3, 0.8, 0.3)), axis=abi.AXIS_YZ)
    s.rect(-4, 4, -4, 4, 0.01, s.metal((0.9, 0.9, 0.9), 0.05), axis=abi.AXIS_XZ)
    return s, cam


def media(aspect=4 / 3):
    """constant_medium with sphere and box boundaries, objects before AND after them in list order."""
    s = Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian((0.5, 0.5, 0.5)))
    s.sphere((-2, 1, 0), 1.0, s.lambertian((0.8, 0.2, 0.2)))
    s.medium_sphere((0, 1, 0), 1.0, 1.0, (1, 1, 1))
    s.sphere((0, 1, 0), 0.4, s.metal((0.9, 0.9, 0.9), 0.0))  # inside the smoke, listed after it
    s.medium_box((1.5, 0, -1), (3.0, 1.5, 0.5), 0.7, (0.1, 0.1, 0.1))
    s.sphere((2.2, 0.7, 1.5), 0.5, s.dielectric(1.5))
    s.box((-1, 0, 2), (0, 0.8, 3), s.lambertian((0.2, 0.2, 0.8)))
    s.medium_sphere((4, 0.6, 2), 0.6, 3.0, (0.2, 0.9, 0.3), center1=(4, 1.0, 2), time0=0.0, time1=1.0)
    s.sphere((1, 3, 1), 0.5, s.lightsource((8, 8, 8)))
    return s, _cam(aspect, look_from=(8, 2.5, 7), vfov=30.0, aperture=0.0, focus=10.0, t0=0.0, t1=1.0)


def ties(aspect=4 / 3):
    """Exact-t ties: duplicated spheres (earlier wins), duplicated rects / triangles (later wins),
    a rect coplanar with a box face, a sphere duplicated around a medium."""
    s = Scene()
    s.sphere((0, 1, 0), 1.0, s.lambertian((0.9, 0.1, 0.1)))
    s.sphere((0, 1, 0), 1.0, s.lambertian((0.1, 0.9, 0.1)))  # same sphere: loses every tie
    s.rect(-3, -1, 0, 2, 0.5, s.lambertian((0.1, 0.1, 0.9)))
    s.rect(-3, -1, 0, 2, 0.5, s.lambertian((0.9, 0.9, 0.1)))  # same rect: wins every tie
    s.triangle((2, 0, 0), (3, 2, 0), (4, 0, 0), s.lambertian((0.9, 0.1, 0.9)))
    s.triangle((2, 0, 0), (3, 2, 0), (4, 0, 0), s.lambertian((0.1, 0.9, 0.9)))
    s.box((-1, 2.2, -1), (1, 3.2, 1), s.lambertian((0.6, 0.6, 0.6)))
    s.rect(-1, 1, 2.2, 3.2, 1.0, s.lambertian((1.0, 0.5, 0.0)))  # coplanar with the box's +z face, listed later
    s.rect(-1, 1, 2.2, 3.2, -1.0, s.lambertian((0.0, 0.5, 1.0)))
    s.sphere((0, -1000, 0), 1000, s.lambertian((0.5, 0.5, 0.5)))
    s.sphere((0, 1, 3), 0.7, s.metal((0.8, 0.8, 0.8), 0.0), center1=(0, 1.3, 3), time0=0.0, time1=1.0)
    s.sphere((0, 1, 3), 0.7, s.lambertian((0.3, 0.3, 0.3)))  # static twin of the moving sphere at t=0
    return s, _cam(aspect, look_from=(3, 3, 9), look_at=(0, 1, 0), vfov=35.0, aperture=0.0, focus=9.0, t0=0.0, t1=0.0)


def empty(aspect=4 / 3):
    return Scene(), _cam(aspect)


def single_light(aspect=4 / 3):
    s = Scene()
    s.sphere((0, 0, 0), 2.0, s.lightsource((2, 3, 4)))
    return s, _cam(aspect)


def triangle_mesh(aspect=16 / 9, nx=12, nz=6, seed=5):
    """Grid of 4-triangle pyramids on a two-triangle ground (BASELINE config 4 layout, small)."""
    rs = np.random.RandomState(seed)
    s = Scene()
    g = s.lambertian(s.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9)))
    s.triangle((-20, 0, -20), (-20, 0, 20), (20, 0, -20), g)
    s.triangle((20, 0, 20), (20, 0, -20), (-20, 0, 20), g)
    for i in range(nx):
        for j in range(nz):
            x, z = -6 + i * 1.0, -3 + j * 1.0
            apex = (x + 0.25, 0.5 + 0.3 * rs.rand(), z + 0.25)
            c = [(x, 0, z), (x + 0.5, 0, z), (x + 0.5, 0, z + 0.5), (x, 0, z + 0.5)]
            mats = [s.lambertian(tuple(rs.rand(3))), s.metal(tuple(0.5 + 0.5 * rs.rand(3)), 0.1 * rs.rand()),
                    s.lambertian(tuple(rs.rand(3))), s.dielectric(1.5)]
            for k in range(4):
                s.triangle(c[k], apex, c[(k + 1) % 4], mats[k])
    s.sphere((0, 3, 0), 1.0, s.lambertian(s.image(test_image(32, 16, 9))))
    return s, _cam(aspect, look_from=(10, 4, 8), vfov=30.0, aperture=0.05, focus=12.0)


def c4_mesh(aspect=16 / 9, nx=100, nz=25, seed=4):
    """BASELINE config 4: ~10 000 triangles -- an nx x nz grid of 4-triangle pyramids (the reference's pyramid
    pattern, main.cpp:113-126) on a two-triangle ground, Lambertian checker / solid with some metal and glass
    faces; image textures only on a few spheres and an xy_rect (triangles do not write u, v: triangle.hpp:94-98).
    Pyramid placement and apex heights come from numpy's RandomState (not LocalPseudoRNG): oracle and GPU get
    the same vectors either way."""
    rs = np.random.RandomState(seed)
    f = np.float32
    s = Scene()
    ground = s.lambertian(s.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9)))
    img = s.image(test_image(64, 32, 9))
    palette = [s.lambertian(tuple(rs.rand(3))) for _ in range(10)]
    palette += [s.lambertian(s.checker(tuple(rs.rand(3)), tuple(rs.rand(3)))) for _ in range(2)]
    palette += [s.metal(tuple(0.5 + 0.5 * rs.rand(3)), 0.1 * rs.rand()) for _ in range(3)] + [s.dielectric(1.5)]
    pitch, base = 0.25, 0.2
    half_x, half_z = 0.5 * nx * pitch, 0.5 * nz * pitch
    gx, gz = half_x + 8.0, half_z + 8.0
    s.triangle((-gx, 0, -gz), (-gx, 0, gz), (gx, 0, -gz), ground)
    s.triangle((gx, 0, gz), (gx, 0, -gz), (-gx, 0, gz), ground)
    for i in range(nx):
        for j in range(nz):
            x, z = f(-half_x + i * pitch), f(-half_z + j * pitch)
            apex = (x + f(0.5 * base), f(0.15 + 0.25 * rs.rand()), z + f(0.5 * base))
            c = [(x, 0, z), (x + f(base), 0, z), (x + f(base), 0, z + f(base)), (x, 0, z + f(base))]
            for k in range(4):
                s.triangle(c[k], apex, c[(k + 1) % 4], palette[rs.randint(0, len(palette))])
    s.sphere((-4, 1.6, -1), 1.0, s.lambertian(img))
    s.sphere((3, 1.3, 0.5), 0.8, s.lambertian(s.image_view(img, 4.0)))
    s.sphere((0, 2.2, -2), 0.7, s.metal((0.8, 0.8, 0.9), 0.05))
    s.rect(-half_x, half_x, 0, 3, -half_z - 1.0, s.lambertian(img))
    return s, _cam(aspect, look_from=(0, 7, 15), look_at=(0, 0.2, 0), vfov=42.0, aperture=0.05, focus=16.0)


def cornell(aspect=1.0):
    """BASELINE config 3: Cornell box.  hittable_t has no rotate/translate and only xy_rect at top level
    (render.hpp:22-23), so floor / ceiling / side walls / light are thin boxes and the back wall an xy_rect;
    two boxes are wrapped in constant_medium smoke (white / black), one is plain."""
    s = Scene()
    red, white, green = s.lambertian((0.65, 0.05, 0.05)), s.lambertian((0.73, 0.73, 0.73)), s.lambertian((0.12, 0.45, 0.15))
    light = s.lightsource((15, 15, 15))
    s.box((555, 0, 0), (556, 555, 555), green)        # left wall
    s.box((-1, 0, 0), (0, 555, 555), red)             # right wall
    s.box((213, 554, 227), (343, 555, 332), light)    # ceiling light
    s.box((0, -1, 0), (555, 0, 555), white)           # floor
    s.box((0, 555, 0), (555, 556, 555), white)        # ceiling
    s.rect(0, 555, 0, 555, 555, white)                # back wall (xy_rect, k = z)
    s.medium_box((130, 0, 65), (295, 165, 230), 0.01, (1, 1, 1))
    s.medium_box((265, 0, 295), (430, 330, 460), 0.01, (0, 0, 0))
    s.box((60, 0, 300), (160, 100, 400), white)
    return s, make_camera((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, aspect, 0.0, 10.0, 0.0, 0.0)


def motion_blur(aspect=16 / 9, seed=2):
    """BASELINE config 5: the default scene's 22x22 grid with ALL small Lambertian spheres moving,
    depth of field (aperture 0.1) and a 0..1 shutter."""
    rs = np.random.RandomState(seed)
    f = np.float32
    s = Scene()
    s.sphere((0, -1000, 0), 1000, s.lambertian(s.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9))))
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = rs.rand()
            c = np.array([f(a + 0.9 * rs.rand()), f(0.2), f(b + 0.9 * rs.rand())], dtype=np.float32)
            if np.linalg.norm(c - np.array([4, 0.2, 0])) <= 0.9:
                continue
            if choose < 0.8:
                c1 = c + np.array([0, f(0.5 * rs.rand()), 0], dtype=np.float32)
                s.sphere(c, 0.2, s.lambertian(tuple(rs.rand(3) * rs.rand(3))), center1=c1, time0=0.0, time1=1.0)
            elif choose < 0.95:
                s.sphere(c, 0.2, s.metal(tuple(0.5 + 0.5 * rs.rand(3)), 0.5 * rs.rand()))
            else:
                s.sphere(c, 0.2, s.dielectric(1.5))
    s.sphere((0, 1, 0), 1.0, s.dielectric(1.5))
    s.sphere((-4, 1, 0), 1.0, s.lambertian((0.4, 0.2, 0.1)))
    s.sphere((4, 1, 0), 1.0, s.metal((0.7, 0.6, 0.5), 0.0))
    return s, make_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0, 0.0, 1.0)


def random_scene(seed, n_objects=60, aspect=4 / 3):
    """Mixed random scene; the seed decides kinds, materials, order."""
    rs = np.random.RandomState(seed)
    s = Scene()
    img = s.image(test_image(8, 8, seed))
    texs = [s.checker(tuple(rs.rand(3)), tuple(rs.rand(3))), img]

    def material():
        k = rs.randint(0, 10)
        if k < 4:
            return s.lambertian(tuple(rs.rand(3)))
        if k < 5:
            return s.lambertian(texs[rs.randint(0, 2)])
        if k < 7:
            return s.metal(tuple(0.4 + 0.6 * rs.rand(3)), rs.rand() * 0.6)
        if k < 9:
            return s.dielectric(1.3 + 0.4 * rs.rand(), tuple(0.7 + 0.3 * rs.rand(3)))
        return s.lightsource(tuple(5 * rs.rand(3)))

    s.sphere((0, -500, 0), 500, s.lambertian(texs[0]))
    for _ in range(n_objects):
        k = rs.randint(0, 12)
        c = np.array([rs.uniform(-6, 6), rs.uniform(0.2, 2.0), rs.uniform(-6, 6)], dtype=np.float32)
        if k < 5:
            s.sphere(c, rs.uniform(0.2, 0.8), material())
        elif k < 7:
            t0 = rs.choice([0.0, 0.2])
            s.sphere(c, rs.uniform(0.2, 0.6), material(), center1=c + rs.uniform(-0.4, 0.4, 3).astype(np.float32),
                     time0=t0, time1=t0 + rs.choice([0.5, 1.0]))
        elif k < 8:
            a0, b0 = rs.uniform(-5, 4), rs.uniform(0, 2)
            s.rect(a0, a0 + rs.uniform(0.3, 2), b0, b0 + rs.uniform(0.3, 2), rs.uniform(-5, 5), material())
        elif k < 10:
            v = [c + rs.uniform(-1, 1, 3).astype(np.float32) for _ in range(3)]
            m = material()
            while s._lists["materials"][m]["kind"] == abi.MAT_LAMBERTIAN and \
                    s._lists["textures"][s._lists["materials"][m]["texture"]]["kind"] == abi.TEX_IMAGE:
                m = material()  # image textures on triangles read indeterminate u,v in the reference
            s.triangle(v[0], v[1], v[2], m)
        elif k < 11:
            s.box(c - 0.4, c + rs.uniform(0.2, 0.9, 3).astype(np.float32), material())
        else:
            if rs.rand() < 0.5:
                s.medium_sphere(c, rs.uniform(0.4, 1.0), rs.uniform(0.3, 2.0), tuple(rs.rand(3)))
            else:
                s.medium_box(c - 0.5, c + 0.5, rs.uniform(0.3, 2.0), tuple(rs.rand(3)))
    return s, _cam(aspect, look_from=(10, 3, 9), look_at=(0, 0.8, 0), vfov=35.0, aperture=0.08, focus=13.0, t0=0.0, t1=1.0)


ALL = {
    "spheres_basic": spheres_basic, "moving": moving, "shapes": shapes, "media": media, "ties": ties,
    "empty": empty, "single_light": single_light, "triangle_mesh": triangle_mesh, "rect_axes": rect_axes,
    "cornell": cornell,
}
# (c4_mesh -- 10 000 triangles -- is too slow for the brute-force oracle at whole-image sizes: see test_config4_*)
REFERENCE_COMPATIBLE = [k for k in ALL if k != "rect_axes"]
