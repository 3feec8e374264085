"""ctypes / numpy mirror of include/pt_abi.h (the C-ABI of the hot path).

Nothing here computes anything: these are the POD layouts of pt_camera,
pt_scene and friends so that Python harness code (tests, bench.py) can hand
scenes to libptb200.so exactly like the C++ host wrapper does.  The layouts
are asserted against the sizes the C header implies.
"""
import ctypes as C

import numpy as np

PT_ABI_VERSION = 1

# error codes
PT_OK, PT_ERR_INVALID_ARGUMENT, PT_ERR_NO_DEVICE, PT_ERR_CUDA, PT_ERR_UNSUPPORTED = 0, -1, -2, -3, -4

# texture_t / material_t / hittable_t variant indices (texture.hpp:154,
# material.hpp:133-135, render.hpp:22-23 of the reference)
TEX_CHECKER, TEX_SOLID, TEX_IMAGE = 0, 1, 2
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_LIGHTSOURCE, MAT_ISOTROPIC = 0, 1, 2, 3, 4
HIT_SPHERE, HIT_RECT, HIT_TRIANGLE, HIT_BOX, HIT_MEDIUM = 0, 1, 2, 3, 4
AXIS_XY, AXIS_XZ, AXIS_YZ = 0, 1, 2
BOUNDARY_SPHERE, BOUNDARY_BOX = 0, 1

ORDER_DT = np.dtype([("kind", "<i4"), ("index", "<i4")], align=True)
SPHERE_DT = np.dtype(
    [("center0", "<f4", 3), ("center1", "<f4", 3), ("radius", "<f4"), ("time0", "<f4"), ("time1", "<f4"),
     ("material", "<i4")], align=True)
RECT_DT = np.dtype(
    [("a0", "<f4"), ("a1", "<f4"), ("b0", "<f4"), ("b1", "<f4"), ("k", "<f4"), ("axis", "<i4"),
     ("material", "<i4")], align=True)
TRIANGLE_DT = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3), ("material", "<i4")], align=True)
BOX_DT = np.dtype([("p0", "<f4", 3), ("p1", "<f4", 3), ("material", "<i4")], align=True)
MEDIUM_DT = np.dtype(
    [("boundary_kind", "<i4"), ("boundary_index", "<i4"), ("density", "<f4"), ("material", "<i4")], align=True)
MATERIAL_DT = np.dtype([("kind", "<i4"), ("texture", "<i4"), ("albedo", "<f4", 3), ("param", "<f4")], align=True)
TEXTURE_DT = np.dtype(
    [("kind", "<i4"), ("color0", "<f4", 3), ("color1", "<f4", 3), ("width", "<u4"), ("height", "<u4"),
     ("offset", "<u8"), ("freq", "<f4"), ("_pad", "<u4")], align=True)
CAMERA_DT = np.dtype(
    [("origin", "<f4", 3), ("lower_left_corner", "<f4", 3), ("horizontal", "<f4", 3), ("vertical", "<f4", 3),
     ("u", "<f4", 3), ("v", "<f4", 3), ("w", "<f4", 3), ("lens_radius", "<f4"), ("time0", "<f4"),
     ("time1", "<f4")], align=True)

assert ORDER_DT.itemsize == 8 and SPHERE_DT.itemsize == 40 and RECT_DT.itemsize == 28
assert TRIANGLE_DT.itemsize == 40 and BOX_DT.itemsize == 28 and MEDIUM_DT.itemsize == 16
assert MATERIAL_DT.itemsize == 24 and TEXTURE_DT.itemsize == 56 and CAMERA_DT.itemsize == 96


class pt_camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class pt_scene(C.Structure):
    _fields_ = [
        ("n_hittables", C.c_uint32), ("order", C.c_void_p),
        ("n_spheres", C.c_uint32), ("spheres", C.c_void_p),
        ("n_rects", C.c_uint32), ("rects", C.c_void_p),
        ("n_triangles", C.c_uint32), ("triangles", C.c_void_p),
        ("n_boxes", C.c_uint32), ("boxes", C.c_void_p),
        ("n_media", C.c_uint32), ("media", C.c_void_p),
        ("n_materials", C.c_uint32), ("materials", C.c_void_p),
        ("n_textures", C.c_uint32), ("textures", C.c_void_p),
        ("n_texture_bytes", C.c_uint64), ("texture_bytes", C.c_void_p),
    ]


class pt_region(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("y_stride", C.c_int32)]


class pt_stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("scans", C.c_uint64), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("n_gpus", C.c_uint32)]


assert C.sizeof(pt_camera) == 96 and C.sizeof(pt_scene) == 144 and C.sizeof(pt_region) == 20

# Every symbol include/pt_abi.h declares (tests check the .so exports all of them).
EXPORTED_SYMBOLS = [
    "pt_render", "render", "pt_render_region", "pt_last_error", "pt_abi_version", "pt_device_count",
    "pt_set_num_gpus", "pt_get_num_gpus", "pt_get_stats", "pt_scene_upload", "pt_scene_free",
    "pt_render_region_device", "pt_scene_read_counters", "pt_scene_launch_count", "pt_fb_alloc", "pt_fb_free", "pt_fb_export",
    "pt_fb_open", "pt_fb_close", "pt_measure_fp32_peak", "pt_render_resume", "pt_render_resume_device", "pt_render_single_task",
]
