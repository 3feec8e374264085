/*
 * pt_abi.h -- C-ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary for ONE hot path of triSYCL/path_tracer: the
 * per-pixel path-tracing loop behind
 *
 *     template <int width, int height, int samples>
 *     void render(sycl::queue&, sycl::buffer<color, 2>& frame_buf,
 *                 std::vector<hittable_t>& hittables, camera& cam);
 *                                     (reference include/render.hpp:141-160)
 *
 * whose device body is render_pixel<> (reference include/render.hpp:25-106)
 * launched one work-item per pixel (include/render.hpp:124-136).
 *
 * Everything here is plain C: pointers, sizes, PODs.  No torch, no C++ types.
 * All pointers are owned by the caller; the callee copies what it needs and
 * retains nothing after return (same ownership as the reference, whose SYCL
 * buffers wrap caller memory: render.hpp:146-148).
 *
 * The scene is the reference's std::variant object list flattened into a
 * type-tagged structure of arrays: one array per primitive kind, one material
 * table, one texture table, the image-texture byte pool, and an ORDER table
 * that preserves the original vector order (closest-hit tie-breaking and the
 * RNG draw order of constant_medium depend on it).
 */
#ifndef PT_ABI_H
#define PT_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_ABI_VERSION 1

/* ---- error codes (pt_* functions return 0 on success, <0 on error) ------ */
enum {
  PT_OK = 0,
  PT_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, bad index */
  PT_ERR_NO_DEVICE = -2,        /* no CUDA device / driver: there is NO CPU fallback */
  PT_ERR_CUDA = -3,             /* a CUDA runtime call failed; see pt_last_error() */
  PT_ERR_UNSUPPORTED = -4       /* scene uses something outside the hot path */
};

/* ---- camera: the 96-byte field order of the reference class -------------
 * reference include/camera.hpp:21-46 (origin, lower_left_corner, horizontal,
 * vertical, u, v, w, lens_radius, time0, time1).  Built on the host by the
 * reference constructor arithmetic (camera.hpp:67-87); consumed on the device
 * by get_ray (camera.hpp:93-100). */
typedef struct pt_camera {
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float u[3];
  float v[3];
  float w[3];
  float lens_radius;
  float time0;
  float time1;
} pt_camera;

/* ---- textures: texture_t = variant<checker, solid, image> ---------------
 * reference include/texture.hpp:154 (index order kept). */
enum { PT_TEX_CHECKER = 0, PT_TEX_SOLID = 1, PT_TEX_IMAGE = 2 };

typedef struct pt_texture {
  int32_t kind;
  float color0[3];  /* solid: the colour (texture.hpp:25); checker: `odd`  (texture.hpp:50) */
  float color1[3];  /* checker: `even` (texture.hpp:51) */
  uint32_t width;   /* image (texture.hpp:75-76) */
  uint32_t height;
  uint64_t offset;  /* image: first TEXEL (not byte) in the pool (texture.hpp:78) */
  float freq;       /* image: cyclic_frequency (texture.hpp:81) */
  uint32_t _pad;
} pt_texture;

/* ---- materials: material_t = variant<lambertian, metal, dielectric,
 *                                      lightsource, isotropic>
 * reference include/material.hpp:133-135 (index order kept). */
enum {
  PT_MAT_LAMBERTIAN = 0,
  PT_MAT_METAL = 1,
  PT_MAT_DIELECTRIC = 2,
  PT_MAT_LIGHTSOURCE = 3,
  PT_MAT_ISOTROPIC = 4
};

typedef struct pt_material {
  int32_t kind;
  int32_t texture;  /* lambertian.albedo / lightsource.emit / isotropic.albedo: index into textures[] */
  float albedo[3];  /* metal.albedo (material.hpp:51), dielectric.albedo (material.hpp:94) */
  float param;      /* metal.fuzz, already clamped to [0,1] (material.hpp:37); dielectric.ref_idx */
} pt_material;

/* ---- hittables: hittable_t = variant<sphere, xy_rect, triangle, box,
 *                                      constant_medium>
 * reference include/render.hpp:22-23 (index order kept). */
enum {
  PT_HIT_SPHERE = 0,
  PT_HIT_RECT = 1,
  PT_HIT_TRIANGLE = 2,
  PT_HIT_BOX = 3,
  PT_HIT_MEDIUM = 4
};

/* sphere.hpp:108-117.  A static sphere has time0 == time1 (sphere.hpp:52). */
typedef struct pt_sphere {
  float center0[3];
  float center1[3];
  float radius;
  float time0;
  float time1;
  int32_t material;
} pt_sphere;

/* rectangle.hpp:50,88,126.  axis 0: xy_rect (a=x, b=y, k=z, normal +z);
 * axis 1: xz_rect (a=x, b=z, k=y, normal +y); axis 2: yz_rect (a=y, b=z,
 * k=x, normal +x).  Only xy_rect is a top-level alternative in the reference
 * (render.hpp:22-23); the other two are accepted here as a superset. */
enum { PT_AXIS_XY = 0, PT_AXIS_XZ = 1, PT_AXIS_YZ = 2 };
typedef struct pt_rect {
  float a0, a1, b0, b1, k;
  int32_t axis;
  int32_t material;
} pt_rect;

/* triangle.hpp:10-12 (v0, v1, v2), Moller-Trumbore strategy (triangle.hpp:122). */
typedef struct pt_triangle {
  float v0[3];
  float v1[3];
  float v2[3];
  int32_t material;
} pt_triangle;

/* box.hpp:52-55: six rectangle sides derived from p0 <= p1 (box.hpp:20-25). */
typedef struct pt_box {
  float p0[3];
  float p1[3];
  int32_t material;
} pt_box;

/* constant_medium.hpp:80-82.  The boundary lives in spheres[] / boxes[] at
 * boundary_index but is NOT listed in the order table (it is not a top-level
 * object).  `density` is the constructor argument; neg_inv_density = -1/density
 * is recomputed by the callee exactly as constant_medium.hpp:20 does. */
enum { PT_BOUNDARY_SPHERE = 0, PT_BOUNDARY_BOX = 1 };
typedef struct pt_medium {
  int32_t boundary_kind;
  int32_t boundary_index;
  float density;
  int32_t material; /* the isotropic phase function */
} pt_medium;

/* One entry per element of the reference's std::vector<hittable_t>, in order. */
typedef struct pt_order_entry {
  int32_t kind;   /* PT_HIT_* */
  int32_t index;  /* index into the per-kind array */
} pt_order_entry;

typedef struct pt_scene {
  uint32_t n_hittables;
  const pt_order_entry* order;
  uint32_t n_spheres;
  const pt_sphere* spheres;
  uint32_t n_rects;
  const pt_rect* rects;
  uint32_t n_triangles;
  const pt_triangle* triangles;
  uint32_t n_boxes;
  const pt_box* boxes;
  uint32_t n_media;
  const pt_medium* media;
  uint32_t n_materials;
  const pt_material* materials;
  uint32_t n_textures;
  const pt_texture* textures;
  /* image_texture::texture_data (texture.hpp:71,157): RGB8 texels, starts with
   * the fallback texel {0,0,1}.  May be NULL/0 when no image texture is used. */
  uint64_t n_texture_bytes;
  const uint8_t* texture_bytes;
} pt_scene;

/* A subset of the image: columns [x0, x0+w), rows y0 + k*y_stride for
 * k in [0, h).  Seeds are always the GLOBAL linear id y*width + x
 * (render.hpp:130-132), so any partition reproduces the full render bit for
 * bit.  Row k of the region is written at out + k*out_row_pitch (floats),
 * pixel x at +3*(x - x0). */
typedef struct pt_region {
  int32_t x0, y0, w, h, y_stride;
} pt_region;

/* Work counters of the last render on this thread's device context. */
typedef struct pt_stats {
  uint64_t paths;        /* camera samples traced (render.hpp:95-101 iterations) */
  uint64_t scans;        /* hit_world calls (render.hpp:60) */
  double kernel_ms;      /* device time of the render kernel(s), CUDA events */
  double h2d_ms;         /* scene + camera upload */
  double d2h_ms;         /* framebuffer download */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint32_t kernel_launches;
  uint32_t n_gpus;
} pt_stats;

/* ------------------------------------------------------------------------
 * Blocking host-buffer entry point: what the reference's render<>() becomes.
 * fb = float[height][width][3], 12 B/pixel, row 0 = bottom image row, exactly
 * fb_acc[y][x] of render.hpp:105.  Uses pt_set_num_gpus() devices (default 1):
 * rows are interleaved across GPUs and peers store straight into GPU 0's
 * framebuffer over NVLink.  Not re-entrant.
 * Replaces: render.hpp:141-160 (+ executor, render.hpp:110-138). */
int pt_render(int width, int height, int spp, int depth, const pt_camera* camera,
              const pt_scene* hitables, float* fb);

/* The literal name the north-star spells; thin alias of pt_render. */
int render(int width, int height, int spp, int depth, const pt_camera* camera,
           const pt_scene* hitables, float* fb);

/* Same, for a sub-region; `out` is a HOST buffer of h rows of out_row_pitch floats. */
int pt_render_region(int width, int height, int spp, int depth, const pt_camera* camera,
                     const pt_scene* hitables, const pt_region* region, float* out,
                     int64_t out_row_pitch);

const char* pt_last_error(void);
int pt_abi_version(void);
int pt_device_count(void);
int pt_set_num_gpus(int n);
int pt_get_num_gpus(void);
int pt_get_stats(pt_stats* out);

/* ------------------------------------------------------------------------
 * Device-resident API (scene uploaded once, framebuffer stays in HBM): what
 * bench.py times as `value`, and what the one-process-per-GPU launcher uses. */
typedef struct pt_device_scene pt_device_scene;

/* Flatten + upload to CUDA device `device`.  Replaces the buffer/accessor set-up
 * of render.hpp:146-155 and image_texture::freeze (texture.hpp:126-131). */
int pt_scene_upload(const pt_scene* scene, int device, pt_device_scene** out);
void pt_scene_free(pt_device_scene* scene);

/* Launch on `stream` (a cudaStream_t, 0 = default), asynchronous.  d_out is a
 * DEVICE pointer (may be a peer-mapped pointer on another GPU).
 * Concurrency: a pt_device_scene owns its pixel queue, hand-off ring, probe buffers and counters, so the launches
 * of ONE scene must be ordered (one stream, or events between streams).  DIFFERENT scenes may be launched on
 * different streams of one device at the same time: the render kernel is a cooperative launch (its CTAs wait for
 * each other), so the runtime runs such launches one after the other instead of half-resident side by side --
 * correct, never deadlocked, no overlap to be gained (tests/test_parity_gpu.py::test_concurrent_scenes_on_two_streams). */
int pt_render_region_device(const pt_device_scene* scene, int width, int height, int spp,
                            int depth, const pt_camera* camera, const pt_region* region,
                            float* d_out, int64_t out_row_pitch, void* stream);

/* Progressive rendering (the reference's per-pixel loop state, render.hpp:94-105): trace
 * samples [spp_from, spp_to) of every pixel of the region.  d_state holds 4 floats per pixel of the region
 * (state_row_pitch in PIXELS per region row): the running sum r, g, b and the pixel's xorshift32 state (bit
 * pattern).  It is read when spp_from > 0 and always written; d_out receives sum / spp_to (render.hpp:102), so
 * that 0..a followed by a..b leaves exactly the framebuffer of one launch 0..b. */
int pt_render_resume_device(const pt_device_scene* scene, int width, int height, int spp_from, int spp_to,
                            int depth, const pt_camera* camera, const pt_region* region, float* d_state,
                            int64_t state_row_pitch, float* d_out, int64_t out_row_pitch, void* stream);
/* The same with host buffers, blocking: `state` is float[region.h][region.w][4] (in and out), `out` as in
 * pt_render_region. */
int pt_render_resume(int width, int height, int spp_from, int spp_to, int depth, const pt_camera* camera,
                     const pt_scene* hitables, const pt_region* region, float* state, float* out,
                     int64_t out_row_pitch);

/* The reference built with -DUSE_SINGLE_TASK (render.hpp:113-122, buildparams::use_single_task): ONE LocalPseudoRNG
 * with its default seed for the whole image, pixels visited x-major (for x: for y:), fb[y][x] as in pt_render.
 * Bit-identical to that build.  The mode is strictly serial -- every sample continues the stream where the previous
 * one left it -- so it runs on ONE warp of GPU 0 (a team of lanes splits each closest-hit scan) and is slower than a
 * host core: a drop-in for callers of that mode, not an accelerated path. */
int pt_render_single_task(int width, int height, int spp, int depth, const pt_camera* camera, const pt_scene* hitables,
                          float* fb);

/* Counters accumulated by launches on this device scene since the last reset
 * (paths, scans only; synchronises the device). */
int pt_scene_read_counters(pt_device_scene* scene, uint64_t* paths, uint64_t* scans, int reset);

/* Kernels launched on this device scene since it was uploaded (a render is up to three: the cost
 * probe, the tile sort and the render kernel itself). */
int pt_scene_launch_count(const pt_device_scene* scene, uint64_t* launches);

/* Peer framebuffer sharing between processes (one process per GPU): rank 0
 * allocates the framebuffer with pt_fb_alloc, exports a 64-byte handle, the
 * other ranks open it and pass the mapped pointer as d_out. */
int pt_fb_alloc(int device, size_t bytes, float** d_ptr);
int pt_fb_free(int device, float* d_ptr);
int pt_fb_export(float* d_ptr, unsigned char handle[64]);
int pt_fb_open(int device, const unsigned char handle[64], float** d_ptr);
int pt_fb_close(float* d_ptr);

/* Sustained FP32 FMA rate of `device` measured by a register-resident FFMA
 * kernel (TFLOP/s, 2 flop per FMA); the roofline denominator bench.py quotes. */
int pt_measure_fp32_peak(int device, double* tflops, double* sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* PT_ABI_H */
