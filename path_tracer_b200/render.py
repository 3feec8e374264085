"""ctypes front end of libptb200.so (include/pt_abi.h).

`render()` is the Python spelling of the reference's
    render<width, height, samples>(queue, frame_buf, hittables, cam)
(reference include/render.hpp:141-160): host scene in, host framebuffer out,
blocking.  `DeviceScene` is the device-resident variant used by bench.py and
the one-process-per-GPU launcher (path_tracer_b200/dist.py).

No fallback: if the CUDA library is missing or no GPU is visible, these raise.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .scene import camera_c

LIB_PATH = os.environ.get("PTB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                        "libptb200.so")


class PathTracerError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libptb200: %s (code %d)" % (message, code))
        self.code = code


_lib = None


def lib():
    """Load libptb200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `make lib` (or __graft_entry__.build()); "
                              "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pt_last_error.restype = C.c_char_p
        L.pt_render.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
        L.render.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
        L.pt_render_region.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_int64]
        L.pt_scene_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pt_scene_free.argtypes = [C.c_void_p]
        L.pt_scene_free.restype = None
        L.pt_render_region_device.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
        L.pt_render_resume_device.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.pt_render_resume.argtypes = [C.c_int] * 5 + [C.c_void_p] * 5 + [C.c_int64]
        L.pt_render_single_task.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
        L.pt_scene_read_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.pt_scene_launch_count.argtypes = [C.c_void_p, C.c_void_p]
        L.pt_get_stats.argtypes = [C.c_void_p]
        L.pt_set_num_gpus.argtypes = [C.c_int]
        L.pt_fb_alloc.argtypes = [C.c_int, C.c_size_t, C.c_void_p]
        L.pt_fb_free.argtypes = [C.c_int, C.c_void_p]
        L.pt_fb_export.argtypes = [C.c_void_p, C.c_void_p]
        L.pt_fb_open.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.pt_fb_close.argtypes = [C.c_void_p]
        L.pt_measure_fp32_peak.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        if L.pt_abi_version() != abi.PT_ABI_VERSION:
            raise ImportError("libptb200.so ABI version mismatch")
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PathTracerError(rc, lib().pt_last_error().decode())


def device_count():
    return lib().pt_device_count()


def set_num_gpus(n):
    _check(lib().pt_set_num_gpus(n))


def stats():
    s = abi.pt_stats()
    _check(lib().pt_get_stats(C.byref(s)))
    return {name: getattr(s, name) for name, _ in s._fields_}


def render(scene, camera, width, height, spp, depth=50, out=None):
    """Blocking render of the full image through pt_render (host buffers both ways)."""
    s, keep = scene.as_c()
    cam = camera_c(camera)
    if out is None:
        out = np.empty((height, width, 3), dtype=np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (height, width, 3)
    _check(lib().pt_render(width, height, spp, depth, C.addressof(cam), C.addressof(s), out.ctypes.data))
    return out


def render_region(scene, camera, width, height, spp, depth, region):
    """Blocking render of a pt_region (global seeds) through pt_render_region."""
    s, keep = scene.as_c()
    cam = camera_c(camera)
    out = np.zeros((region.h, region.w, 3), dtype=np.float32)
    _check(lib().pt_render_region(width, height, spp, depth, C.addressof(cam), C.addressof(s), C.addressof(region),
                                  out.ctypes.data, region.w * 3))
    return out


def render_single_task(scene, camera, width, height, spp, depth=50):
    """The reference's USE_SINGLE_TASK executor (render.hpp:113-122) through pt_render_single_task: one generator for the
    whole image, pixels x-major.  Serial by construction: small images only."""
    s, keep = scene.as_c()
    cam = camera_c(camera)
    out = np.empty((height, width, 3), dtype=np.float32)
    _check(lib().pt_render_single_task(width, height, spp, depth, C.addressof(cam), C.addressof(s), out.ctypes.data))
    return out


def render_resume(scene, camera, width, height, spp_from, spp_to, depth, region, state=None):
    """Progressive rendering through pt_render_resume: samples [spp_from, spp_to) of every pixel of `region`.
    -> (framebuffer = sum / spp_to, state [region.h, region.w, 4] to pass to the next call)."""
    s, keep = scene.as_c()
    cam = camera_c(camera)
    if state is None:
        assert spp_from == 0
        state = np.zeros((region.h, region.w, 4), dtype=np.float32)
    state = np.ascontiguousarray(state, dtype=np.float32).copy()
    out = np.zeros((region.h, region.w, 3), dtype=np.float32)
    _check(lib().pt_render_resume(width, height, spp_from, spp_to, depth, C.addressof(cam), C.addressof(s), C.addressof(region),
                                  state.ctypes.data, out.ctypes.data, region.w * 3))
    return out, state


class DeviceScene:
    """A scene resident in HBM on one GPU (pt_scene_upload)."""

    def __init__(self, scene, device=0):
        s, keep = scene.as_c()
        self._h = C.c_void_p()
        _check(lib().pt_scene_upload(C.addressof(s), device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().pt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_region(self, camera, width, height, spp, depth, region, d_out, out_row_pitch, stream=0):
        """Asynchronous launch; d_out is a device pointer (int), stream a cudaStream_t (int)."""
        cam = camera if isinstance(camera, abi.pt_camera) else camera_c(camera)
        _check(lib().pt_render_region_device(self._h, width, height, spp, depth, C.addressof(cam),
                                             C.addressof(region), C.c_void_p(d_out), out_row_pitch,
                                             C.c_void_p(stream)))

    def closest_hit(self, camera, rays7, seeds, mode=0):
        """Test hook (pt_debug_closest_hit): the closest-hit scan of given rays -> (t, vector index or -1, generator after)."""
        cam = camera if isinstance(camera, abi.pt_camera) else camera_c(camera)
        rays7 = np.ascontiguousarray(rays7, dtype=np.float32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = rays7.shape[0]
        t, idx, rng = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
        L = lib()
        L.pt_debug_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.pt_debug_closest_hit(self._h, C.addressof(cam), n, rays7.ctypes.data, seeds.ctypes.data, mode, t.ctypes.data, idx.ctypes.data,
                                      rng.ctypes.data))
        return t, idx, rng

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().pt_scene_launch_count(self._h, C.byref(n)))
        return n.value

    def counters(self, reset=False):
        paths, scans = C.c_uint64(), C.c_uint64()
        _check(lib().pt_scene_read_counters(self._h, C.byref(paths), C.byref(scans), int(reset)))
        return paths.value, scans.value


def measure_fp32_peak(device=0):
    """(TFLOP/s of register-resident FFMA, implied SM clock in MHz)."""
    t, mhz = C.c_double(), C.c_double()
    _check(lib().pt_measure_fp32_peak(device, C.byref(t), C.byref(mhz)))
    return t.value, mhz.value


def rows_region(width, height, first, stride):
    """Rows first, first+stride, ... (the multi-GPU row interleave)."""
    n = (height - first + stride - 1) // stride if first < height else 0
    return abi.pt_region(0, first, width, n, stride)
