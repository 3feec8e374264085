"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracles.

  CPort : oracle/libpt_oracle.so   plain-C restatement (oracle/pt_oracle.c)
  Ref   : oracle/_ref/libptref.so  the unmodified reference headers compiled in place

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` legs import this module.  The product (path_tracer_b200) never does.
"""
import ctypes as C
import os

import numpy as np

from path_tracer_b200 import abi
from path_tracer_b200.scene import camera_c

HERE = os.path.dirname(os.path.abspath(__file__))
CPORT_PATH = os.path.join(HERE, "libpt_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libptref.so")


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("scans", C.c_uint64), ("tests", C.c_uint64 * 5),
                ("moving_sphere_tests", C.c_uint64), ("accepts", C.c_uint64 * 5), ("scatters", C.c_uint64 * 5),
                ("sky", C.c_uint64), ("absorbed", C.c_uint64), ("exhausted", C.c_uint64), ("draws", C.c_uint64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else int(v)
        return d


def full_region(width, height):
    return abi.pt_region(0, 0, width, height, 1)


def rows_region(width, height, first, stride):
    n = (height - first + stride - 1) // stride if first < height else 0
    return abi.pt_region(0, first, width, n, stride)


class _Base:
    prefix = None

    def _kat_common(self):
        L, p = self.lib, self.prefix
        getattr(L, p + "kat_xorshift").argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        getattr(L, p + "kat_float").argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        getattr(L, p + "kat_vec").argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        getattr(L, p + "kat_get_ray").argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        getattr(L, p + "kat_hit_scatter").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        for n in ("kat_xorshift", "kat_float", "kat_vec", "kat_get_ray"):
            getattr(L, p + n).restype = None

    def xorshift(self, seed, n):
        out = np.zeros(n, dtype=np.uint32)
        getattr(self.lib, self.prefix + "kat_xorshift")(seed, n, out.ctypes.data)
        return out

    def floats(self, seed, n):
        out = np.zeros(n, dtype=np.float32)
        getattr(self.lib, self.prefix + "kat_float")(seed, n, out.ctypes.data)
        return out

    def vecs(self, seed, kind, n):
        """kind: 0 unit_vec, 1 in_unit_ball, 2 in_unit_disk, 3 vec_t"""
        out = np.zeros((n, 3), dtype=np.float32)
        getattr(self.lib, self.prefix + "kat_vec")(seed, kind, n, out.ctypes.data)
        return out

    def get_rays(self, camera, seed, st):
        st = np.ascontiguousarray(st, dtype=np.float32)
        out = np.zeros((st.shape[0], 7), dtype=np.float32)
        cam = camera_c(camera)
        getattr(self.lib, self.prefix + "kat_get_ray")(C.byref(cam), seed, st.shape[0], st.ctypes.data,
                                                      out.ctypes.data)
        return out

    def hit_scatter(self, scene, ray7, seed):
        s, keep = scene.as_c()
        ray7 = np.ascontiguousarray(ray7, dtype=np.float32)
        out = np.zeros(26, dtype=np.float32)
        rc = getattr(self.lib, self.prefix + "kat_hit_scatter")(C.byref(s), ray7.ctypes.data, seed, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("hit_scatter failed")
        return out


class CPort(_Base):
    """oracle/pt_oracle.c -- arbitrary sizes, with work counters."""
    prefix = "pt_oracle_"
    kind = "port"

    def __init__(self, path=CPORT_PATH):
        self.lib = C.CDLL(path)
        self.lib.pt_oracle_render_region.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int,
                                                                                       C.c_void_p]
        self.lib.pt_oracle_render_region.restype = C.c_int
        self._kat_common()

    def max_threads(self):
        return self.lib.pt_oracle_max_threads()

    def render_region(self, scene, camera, width, height, spp, depth, region=None, dynamic=True, nthreads=0):
        region = region or full_region(width, height)
        s, keep = scene.as_c()
        cam = camera_c(camera)
        out = np.zeros((region.h, region.w, 3), dtype=np.float32)
        cnt = Counters()
        rc = self.lib.pt_oracle_render_region(width, height, spp, depth, C.addressof(cam), C.addressof(s),
                                              C.addressof(region), out.ctypes.data, region.w * 3, int(dynamic),
                                              nthreads, C.addressof(cnt))
        if rc != 0:
            raise RuntimeError("pt_oracle_render_region failed (%d)" % rc)
        return out, cnt

    def render(self, scene, camera, width, height, spp, depth, **kw):
        return self.render_region(scene, camera, width, height, spp, depth, None, **kw)

    def hit_world_batch(self, scene, rays7, seeds):
        """hit_world (render.hpp:30-51) per ray -> (t [n] float32, vector index of the hit object or -1, generator after)."""
        self.lib.pt_oracle_hit_world_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        s, keep = scene.as_c()
        rays7 = np.ascontiguousarray(rays7, dtype=np.float32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = rays7.shape[0]
        t, idx, rng = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
        rc = self.lib.pt_oracle_hit_world_batch(C.addressof(s), n, rays7.ctypes.data, seeds.ctypes.data, t.ctypes.data, idx.ctypes.data,
                                                rng.ctypes.data)
        if rc != 0:
            raise RuntimeError("pt_oracle_hit_world_batch failed (%d)" % rc)
        return t, idx, rng

    def render_single_task(self, scene, camera, width, height, spp, depth):
        """The reference's USE_SINGLE_TASK mode (render.hpp:113-122): one RNG stream for the whole image."""
        self.lib.pt_oracle_render_single_task.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4
        s, keep = scene.as_c()
        cam = camera_c(camera)
        out = np.zeros((height, width, 3), dtype=np.float32)
        cnt = Counters()
        rc = self.lib.pt_oracle_render_single_task(width, height, spp, depth, C.addressof(cam), C.addressof(s),
                                                   out.ctypes.data, C.addressof(cnt))
        if rc != 0:
            raise RuntimeError("pt_oracle_render_single_task failed (%d)" % rc)
        return out, cnt


class Ref(_Base):
    """oracle/_ref/libptref.so -- the reference's own code; fixed list of template instantiations."""
    prefix = "ptref_"
    kind = "reference"

    def __init__(self, path=REF_PATH):
        self.lib = C.CDLL(path)
        self.lib.ptref_render_region.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int]
        self.lib.ptref_render_full.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int, C.c_int]
        self.lib.ptref_last_error.restype = C.c_char_p
        self.lib.ptref_make_camera.argtypes = [C.c_void_p] * 3 + [C.c_float] * 6 + [C.c_void_p]
        self.lib.ptref_make_camera.restype = None
        self._kat_common()

    @staticmethod
    def available(path=REF_PATH):
        return os.path.exists(path)

    def max_threads(self):
        return self.lib.ptref_max_threads()

    def supported(self, width, height, spp, depth):
        return bool(self.lib.ptref_supported(width, height, spp, depth))

    def configs(self):
        buf = (C.c_int * 400)()
        n = self.lib.ptref_list_configs(buf, 100)
        return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]

    def render_region(self, scene, camera, width, height, spp, depth, region=None, dynamic=True, nthreads=0):
        region = region or full_region(width, height)
        s, keep = scene.as_c()
        cam = camera_c(camera)
        out = np.zeros((region.h, region.w, 3), dtype=np.float32)
        rc = self.lib.ptref_render_region(width, height, spp, depth, C.addressof(cam), C.addressof(s),
                                          C.addressof(region), out.ctypes.data, region.w * 3, int(dynamic), nthreads)
        if rc != 0:
            raise RuntimeError("ptref_render_region: %s" % self.lib.ptref_last_error().decode())
        return out

    def render_full(self, scene, camera, width, height, spp, dynamic=False, nthreads=0):
        """The reference's own render<W,H,S>() entry point (depth 50)."""
        s, keep = scene.as_c()
        cam = camera_c(camera)
        out = np.zeros((height, width, 3), dtype=np.float32)
        rc = self.lib.ptref_render_full(width, height, spp, C.addressof(cam), C.addressof(s), out.ctypes.data,
                                        int(dynamic), nthreads)
        if rc != 0:
            raise RuntimeError("ptref_render_full: %s" % self.lib.ptref_last_error().decode())
        return out

    def make_camera(self, look_from, look_at, vup, vfov_deg, aspect, aperture, focus_dist, t0=0.0, t1=0.0):
        a = [np.ascontiguousarray(v, dtype=np.float32) for v in (look_from, look_at, vup)]
        out = np.zeros((), dtype=abi.CAMERA_DT)
        buf = np.zeros(96, dtype=np.uint8)
        self.lib.ptref_make_camera(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, vfov_deg, aspect, aperture,
                                   focus_dist, t0, t1, buf.ctypes.data)
        return np.frombuffer(buf.tobytes(), dtype=abi.CAMERA_DT)[0].copy()


def compare(a, b):
    """Parity metrics on linear framebuffers: (mean-abs-error, PSNR dB with peak 1.0, fraction bit-identical)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    d = a.astype(np.float64) - b.astype(np.float64)
    mae = float(np.mean(np.abs(d)))
    mse = float(np.mean(d * d))
    psnr = float("inf") if mse == 0 else float(10.0 * np.log10(1.0 / mse))
    same = float(np.mean(np.all(a.view(np.uint32) == b.view(np.uint32), axis=-1)))
    return mae, psnr, same
