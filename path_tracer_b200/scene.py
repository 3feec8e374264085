"""Flat scenes (pt_scene of include/pt_abi.h) as numpy structured arrays.

`Scene` is the Python-side twin of the C++ flattener
(path_tracer_b200/include/pt/flatten.hpp): it keeps one array per primitive
kind, the material / texture tables, the image byte pool and the ORDER table,
reads and writes PTSCENE1 files (include/ptscene_io.hpp) and produces the
ctypes `pt_scene` the C-ABI takes.  The add_* helpers mirror the reference's
scene-building API (sphere / xy_rect / triangle / box / constant_medium;
lambertian / metal / dielectric / lightsource / isotropic; solid / checker /
image textures) so tests read like the reference's main.cpp scene script.
"""
import ctypes as C
import struct

import numpy as np

from . import abi

_KINDS = [("order", abi.ORDER_DT), ("spheres", abi.SPHERE_DT), ("rects", abi.RECT_DT),
          ("triangles", abi.TRIANGLE_DT), ("boxes", abi.BOX_DT), ("media", abi.MEDIUM_DT),
          ("materials", abi.MATERIAL_DT), ("textures", abi.TEXTURE_DT)]


class Scene:
    def __init__(self):
        self._lists = {name: [] for name, _ in _KINDS}
        # image_texture::texture_data starts with the fallback texel (texture.hpp:157)
        self.texture_bytes = np.array([0, 0, 1], dtype=np.uint8)
        self._frozen = None

    # ------------------------------------------------------------------ tables
    def _append(self, name, dt, **fields):
        rec = np.zeros((), dtype=dt)
        for k, v in fields.items():
            rec[k] = v
        self._lists[name].append(rec)
        self._frozen = None
        return len(self._lists[name]) - 1

    # textures ---------------------------------------------------------------
    def solid(self, color):
        return self._append("textures", abi.TEXTURE_DT, kind=abi.TEX_SOLID, color0=color)

    def checker(self, odd, even):
        """checker_texture(c1, c2): odd = c1, even = c2 (texture.hpp:38-40)."""
        return self._append("textures", abi.TEXTURE_DT, kind=abi.TEX_CHECKER, color0=odd, color1=even)

    def image(self, rgb8, freq=1.0):
        """image_texture_factory (texture.hpp:97-117): append RGB8 texels (H,W,3) to the pool."""
        rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
        h, w, c = rgb8.shape
        assert c == 3
        offset = self.texture_bytes.size // 3
        self.texture_bytes = np.concatenate([self.texture_bytes, rgb8.reshape(-1)])
        return self._append("textures", abi.TEXTURE_DT, kind=abi.TEX_IMAGE, width=w, height=h, offset=offset,
                            freq=freq)

    def image_view(self, tex_index, freq):
        """A second image_texture over already-loaded texels, different cyclic frequency."""
        t = self._lists["textures"][tex_index]
        return self._append("textures", abi.TEXTURE_DT, kind=abi.TEX_IMAGE, width=t["width"], height=t["height"],
                            offset=t["offset"], freq=freq)

    def _tex(self, t):
        return t if isinstance(t, (int, np.integer)) else self.solid(t)

    # materials ----------------------------------------------------------------
    def lambertian(self, tex_or_color):
        return self._append("materials", abi.MATERIAL_DT, kind=abi.MAT_LAMBERTIAN, texture=self._tex(tex_or_color))

    def metal(self, albedo, fuzz):
        fuzz = float(np.clip(np.float32(fuzz), np.float32(0), np.float32(1)))  # material.hpp:37
        return self._append("materials", abi.MATERIAL_DT, kind=abi.MAT_METAL, texture=-1, albedo=albedo, param=fuzz)

    def dielectric(self, ref_idx, albedo=(1.0, 1.0, 1.0)):
        return self._append("materials", abi.MATERIAL_DT, kind=abi.MAT_DIELECTRIC, texture=-1, albedo=albedo,
                            param=ref_idx)

    def lightsource(self, tex_or_color):
        return self._append("materials", abi.MATERIAL_DT, kind=abi.MAT_LIGHTSOURCE, texture=self._tex(tex_or_color))

    def isotropic(self, tex_or_color):
        return self._append("materials", abi.MATERIAL_DT, kind=abi.MAT_ISOTROPIC, texture=self._tex(tex_or_color))

    # hittables ------------------------------------------------------------------
    def _top(self, kind, index, top_level):
        if top_level:
            self._append("order", abi.ORDER_DT, kind=kind, index=index)
        return index

    def sphere(self, center, radius, material, center1=None, time0=0.0, time1=0.0, top_level=True):
        i = self._append("spheres", abi.SPHERE_DT, center0=center, center1=center if center1 is None else center1,
                         radius=radius, time0=time0, time1=time1, material=material)
        return self._top(abi.HIT_SPHERE, i, top_level)

    def rect(self, a0, a1, b0, b1, k, material, axis=abi.AXIS_XY):
        i = self._append("rects", abi.RECT_DT, a0=a0, a1=a1, b0=b0, b1=b1, k=k, axis=axis, material=material)
        return self._top(abi.HIT_RECT, i, True)

    def triangle(self, v0, v1, v2, material):
        i = self._append("triangles", abi.TRIANGLE_DT, v0=v0, v1=v1, v2=v2, material=material)
        return self._top(abi.HIT_TRIANGLE, i, True)

    def box(self, p0, p1, material, top_level=True):
        i = self._append("boxes", abi.BOX_DT, p0=p0, p1=p1, material=material)
        return self._top(abi.HIT_BOX, i, top_level)

    def medium_sphere(self, center, radius, density, tex_or_color, **kw):
        b = self.sphere(center, radius, self.lambertian((0.5, 0.5, 0.5)), top_level=False, **kw)
        i = self._append("media", abi.MEDIUM_DT, boundary_kind=abi.BOUNDARY_SPHERE, boundary_index=b, density=density,
                         material=self.isotropic(tex_or_color))
        return self._top(abi.HIT_MEDIUM, i, True)

    def medium_box(self, p0, p1, density, tex_or_color):
        b = self.box(p0, p1, self.lambertian((0.5, 0.5, 0.5)), top_level=False)
        i = self._append("media", abi.MEDIUM_DT, boundary_kind=abi.BOUNDARY_BOX, boundary_index=b, density=density,
                         material=self.isotropic(tex_or_color))
        return self._top(abi.HIT_MEDIUM, i, True)

    # ------------------------------------------------------------------ arrays
    def arrays(self):
        if self._frozen is None:
            self._frozen = {}
            for name, dt in _KINDS:
                lst = self._lists[name]
                self._frozen[name] = np.array(lst, dtype=dt) if len(lst) else np.zeros(0, dtype=dt)
        return self._frozen

    def __getattr__(self, name):
        if name in ("order", "spheres", "rects", "triangles", "boxes", "media", "materials", "textures"):
            return self.arrays()[name]
        raise AttributeError(name)

    @property
    def n_hittables(self):
        return len(self._lists["order"])

    def as_c(self):
        """(pt_scene, keepalive).  The arrays must outlive the C call."""
        a = self.arrays()
        tb = np.ascontiguousarray(self.texture_bytes, dtype=np.uint8)
        s = abi.pt_scene()
        keep = [tb]
        for (name, _), (nfield, pfield) in zip(_KINDS, [
                ("n_hittables", "order"), ("n_spheres", "spheres"), ("n_rects", "rects"),
                ("n_triangles", "triangles"), ("n_boxes", "boxes"), ("n_media", "media"),
                ("n_materials", "materials"), ("n_textures", "textures")]):
            arr = np.ascontiguousarray(a[name])
            keep.append(arr)
            setattr(s, nfield, arr.shape[0])
            setattr(s, pfield, arr.ctypes.data if arr.shape[0] else None)
        s.n_texture_bytes = tb.size
        s.texture_bytes = tb.ctypes.data if tb.size else None
        return s, keep

    # ------------------------------------------------------------------ PTSCENE1
    def save(self, path, camera, width, height, spp, depth):
        a = self.arrays()
        with open(path, "wb") as f:
            f.write(b"PTSCENE1")
            f.write(struct.pack("<8I", *[a[name].shape[0] for name, _ in _KINDS]))
            f.write(struct.pack("<Q", int(self.texture_bytes.size)))
            f.write(struct.pack("<4i", width, height, spp, depth))
            f.write(np.asarray(camera, dtype=abi.CAMERA_DT).tobytes())
            for name, _ in _KINDS:
                f.write(np.ascontiguousarray(a[name]).tobytes())
            f.write(np.ascontiguousarray(self.texture_bytes).tobytes())

    @staticmethod
    def load(path):
        """Returns (scene, camera ndarray[CAMERA_DT], (width, height, spp, depth))."""
        import gzip
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            buf = f.read()
        if buf[:8] != b"PTSCENE1":
            raise ValueError("%s: not a PTSCENE1 file" % path)
        counts = struct.unpack_from("<8I", buf, 8)
        (nbytes,) = struct.unpack_from("<Q", buf, 40)
        meta = struct.unpack_from("<4i", buf, 48)
        off = 64
        camera = np.frombuffer(buf, dtype=abi.CAMERA_DT, count=1, offset=off)[0].copy()
        off += 96
        sc = Scene()
        for (name, dt), n in zip(_KINDS, counts):
            sc._lists[name] = list(np.frombuffer(buf, dtype=dt, count=n, offset=off).copy())
            off += n * dt.itemsize
        sc.texture_bytes = np.frombuffer(buf, dtype=np.uint8, count=nbytes, offset=off).copy()
        return sc, camera, meta


def camera_c(camera):
    """numpy CAMERA_DT record -> ctypes pt_camera."""
    c = abi.pt_camera()
    C.memmove(C.byref(c), np.asarray(camera, dtype=abi.CAMERA_DT).tobytes(), 96)
    return c


def make_camera(look_from, look_at, vup, vfov_deg, aspect, aperture, focus_dist, time0=0.0, time1=0.0):
    """The reference camera constructor (camera.hpp:67-87) in float32 numpy, op for op.

    tan() is taken from C's tanf via numpy float32; tests/test_host.py checks this
    against the reference-built camera of the C1 fixture.
    """
    f = np.float32

    def v(x):
        return np.asarray(x, dtype=np.float32)

    def dot(a, b):
        return f(f(f(a[0] * b[0]) + f(a[1] * b[1])) + f(a[2] * b[2]))

    def cross(a, b):
        return np.array([f(f(a[1] * b[2]) - f(a[2] * b[1])), f(f(a[2] * b[0]) - f(a[0] * b[2])),
                         f(f(a[0] * b[1]) - f(a[1] * b[0]))], dtype=np.float32)

    def unit(a):
        return (a / np.sqrt(dot(a, a))).astype(np.float32)

    pi = f(3.1415926535897932385)
    look_from, look_at, vup = v(look_from), v(look_at), v(vup)
    theta = f(f(f(vfov_deg) * pi) / f(180.0))
    h = np.tan(f(theta / f(2)), dtype=np.float32)
    viewport_height = f(f(2.0) * h)
    viewport_width = f(f(aspect) * viewport_height)
    w = unit(look_from - look_at)
    u = unit(cross(vup, w))
    vv = cross(w, u)
    focus = f(focus_dist)
    horizontal = (f(focus * viewport_width) * u).astype(np.float32)
    vertical = (f(focus * viewport_height) * vv).astype(np.float32)
    llc = (((look_from - horizontal / f(2)).astype(np.float32) - vertical / f(2)).astype(np.float32)
           - (focus * w).astype(np.float32)).astype(np.float32)
    cam = np.zeros((), dtype=abi.CAMERA_DT)
    cam["origin"], cam["lower_left_corner"], cam["horizontal"], cam["vertical"] = look_from, llc, horizontal, vertical
    cam["u"], cam["v"], cam["w"] = u, vv, w
    cam["lens_radius"] = f(f(aperture) / f(2))
    cam["time0"], cam["time1"] = time0, time1
    return cam
