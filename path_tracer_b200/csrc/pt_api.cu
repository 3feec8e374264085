// pt_api.cu -- the C-ABI of include/pt_abi.h on top of the sm_100a kernel.
//
// There is no CPU path in this library: without a CUDA device every entry
// point that would render returns PT_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "pt_abi.h"
#include "pt_kernel.h"
#include "pt_pack.h"
#include "pt_packed.h"

using namespace ptb;

namespace {

thread_local std::string g_error;
int g_num_gpus = 1;
int g_team_size_override = 0;
int g_cull_enabled = 1;
int g_lpt_enabled = 1;  // pt_debug_set_lpt
int g_n_express = -1;   // < 0: automatic (pt_debug_set_express)
int g_kernel_kind = 0;  // 0 = wavefront kernel, 1 = lane kernel (pt_debug_set_kernel)
bool g_single_task = false;  // set for the duration of pt_render_single_task()
int g_fb_stage_enabled = 1;  // pt_debug_set_fb_stage: 0 = scalar stores straight into the caller's framebuffer
pt_stats g_stats {};

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return PT_ERR_NO_DEVICE;
  return PT_ERR_CUDA;
}
#define PT_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

constexpr int kCounterSlots = 64;
constexpr unsigned kHeavyCap = 32768;  // slots of the heavy-pixel hand-off ring (a power of two)
static_assert((kHeavyCap & (kHeavyCap - 1)) == 0, "the ring indexes with pos & (cap - 1)");
constexpr int kCounterWords = 24;  // [0] scans, [1] first CTA start (ns), [2] queue ran dry (ns), [3] last warp retired (ns)

}  // namespace

// One uploaded scene: a single device arena holding the scan blob, the side
// tables, materials, textures and the image byte pool (memory laid out once,
// resident in HBM for as long as the handle lives).
struct pt_device_scene {
  int device = 0;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  SceneDesc desc {};
  unsigned long long* queue_heads = nullptr;  // kCounterSlots work-queue heads
  unsigned long long* counters = nullptr;     // [0] scans, [1] paths
  unsigned int* heavy_ctrl = nullptr;
  unsigned int* heavy_ready = nullptr;
  float* heavy_entries = nullptr;
  unsigned int launch_stamp = 0;
  int* lpt_buf = nullptr;  // probe costs + tile order + scratch of the LPT pixel ordering
  size_t lpt_ints = 0;
  FrameTuning* last_tuning = nullptr;  // device: what the last cost probe said (pt_debug_frame_tuning)
  float* fb_stage = nullptr;  // float4 staging image of the last region size (pt_kernel.h: launch_resolve_fb)
  size_t fb_stage_bytes = 0;
  uint2* tree_spill = nullptr;  // scenes with flat trees: the tree lists' continuation in global memory (pt_wave.cu)
  size_t tree_spill_bytes = 0;
  int next_slot = 0;
  unsigned int kernel_launches = 0;  // kernels launched since creation (probe, tile sort, render)
  unsigned long long paths_launched = 0;
  LaunchInfo last_launch {};
  // chunk boxes (pt_pack.h): what they are computed from, and the shutter interval they were last computed for
  PackedScene geo;
  bool cull_valid = false;
  int cull_mode = -1;
  float cull_time0 = 0.f, cull_time1 = 0.f;
};

namespace {

// Device allocations are recycled across uploads (cudaMalloc / cudaFree cost milliseconds and
// serialise the device; with peer mappings alive a cudaFree can take 100+ ms).  Only device memory is
// kept -- never anything of the caller's.
struct CachedBlock {
  int device;
  void* ptr;
  size_t bytes;
};
std::mutex g_cache_mutex;
std::vector<CachedBlock> g_cache;
constexpr size_t kCacheEntries = 8;

cudaError_t cached_malloc(int device, void** out, size_t bytes, size_t* got_bytes) {
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    size_t best = g_cache.size();
    for (size_t i = 0; i < g_cache.size(); ++i)
      if (g_cache[i].device == device && g_cache[i].bytes >= bytes && g_cache[i].bytes <= bytes * 2 + (1u << 20) &&
          (best == g_cache.size() || g_cache[i].bytes < g_cache[best].bytes))
        best = i;
    if (best != g_cache.size()) {
      *out = g_cache[best].ptr, *got_bytes = g_cache[best].bytes;
      g_cache.erase(g_cache.begin() + (long)best);
      return cudaSuccess;
    }
  }
  *got_bytes = bytes;
  return cudaMalloc(out, bytes);
}

void cached_free(int device, void* ptr, size_t bytes) {
  if (!ptr) return;
  void* evict = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cache.push_back(CachedBlock { device, ptr, bytes });
    if (g_cache.size() > kCacheEntries) {
      evict = g_cache.front().ptr;
      device = g_cache.front().device;
      g_cache.erase(g_cache.begin());
    }
  }
  if (evict) {  // free on the block's own device, and leave the caller's current device as it was
    int current = 0;
    cudaGetDevice(&current);
    cudaSetDevice(device);
    cudaFree(evict);
    cudaSetDevice(current);
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T> size_t place(std::vector<unsigned char>& host, const std::vector<T>& v) {
  const size_t off = align_up(host.size(), 256);
  host.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) std::memcpy(host.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

int upload(const pt_scene* scene, int device, pt_device_scene** out, double* h2d_ms, uint64_t* h2d_bytes) {
  if (!scene || !out) return fail(PT_ERR_INVALID_ARGUMENT, "pt_scene_upload: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PT_ERR_INVALID_ARGUMENT, "pt_scene_upload: bad device index");

  PackedScene ps;
  std::string err;
  const int rc = pack_scene(*scene, ps, err);
  if (rc != PT_OK) return fail(rc, err);

  std::vector<unsigned char> host;
  const size_t o_blob = place(host, ps.blob);
  const size_t o_saux = place(host, ps.sphere_aux);
  const size_t o_maux = place(host, ps.moving_aux);
  const size_t o_raux = place(host, ps.rect_aux);
  const size_t o_taux = place(host, ps.tri_aux);
  const size_t o_baux = place(host, ps.box_aux);
  const size_t o_media = place(host, ps.media);
  std::vector<int32_t> keys, object_id;
  uint32_t key_base[6];
  build_key_tables(ps, keys, key_base, object_id);
  const size_t o_keys = place(host, keys);
  const size_t o_objid = place(host, object_id);
  const size_t o_mat = place(host, ps.materials);
  const size_t stage_end = align_up(host.size(), 16);  // what the wavefront kernel keeps in shared memory if it fits
  const size_t o_tex = place(host, ps.textures);
  const size_t o_heads = align_up(host.size(), 256);
  host.resize(o_heads + sizeof(unsigned long long) * (kCounterSlots + kCounterWords), 0);
  const size_t o_hctrl = align_up(host.size(), 256);
  host.resize(o_hctrl + 64, 0);
  const size_t o_bytes = align_up(host.size(), 256);
  const size_t tex_bytes = std::max<size_t>((size_t)scene->n_texture_bytes, 3);
  const size_t o_hready = o_bytes + align_up(tex_bytes, 256);
  const size_t o_hentries = o_hready + align_up(sizeof(unsigned) * kHeavyCap, 256);
  const size_t total = o_hentries + sizeof(float) * kHeavyEntryWords * (size_t)kHeavyCap;

  PT_CUDA(cudaSetDevice(device));
  auto* ds = new pt_device_scene;
  ds->device = device;
  void* arena = nullptr;
  e = cached_malloc(device, &arena, total, &ds->arena_bytes);
  if (e != cudaSuccess) {
    delete ds;
    return cuda_fail(e, "cudaMalloc(scene arena)");
  }
  ds->arena = static_cast<unsigned char*>(arena);
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0), cudaEventCreate(&ev1);
  cudaEventRecord(ev0, 0);
  e = cudaMemcpyAsync(ds->arena, host.data(), host.size(), cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess && scene->n_texture_bytes)
    e = cudaMemcpyAsync(ds->arena + o_bytes, scene->texture_bytes, scene->n_texture_bytes, cudaMemcpyHostToDevice, 0);
  else if (e == cudaSuccess)
    e = cudaMemsetAsync(ds->arena + o_bytes, 0, 3, 0);
  std::vector<unsigned> ring_init(kHeavyCap);  // slot i is ready to be written for position i
  for (unsigned i = 0; i < kHeavyCap; ++i) ring_init[i] = i;
  if (e == cudaSuccess) e = cudaMemcpyAsync(ds->arena + o_hready, ring_init.data(), sizeof(unsigned) * kHeavyCap, cudaMemcpyHostToDevice, 0);
  cudaEventRecord(ev1, 0);
  if (e == cudaSuccess) e = cudaEventSynchronize(ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  cudaEventDestroy(ev0), cudaEventDestroy(ev1);
  if (e != cudaSuccess) {
    cached_free(device, ds->arena, ds->arena_bytes);
    delete ds;
    return cuda_fail(e, "scene upload");
  }
  if (h2d_ms) *h2d_ms += ms;
  if (h2d_bytes) *h2d_bytes += host.size() + scene->n_texture_bytes;

  SceneDesc& d = ds->desc;
  d.blob = ds->arena + o_blob;
  d.blob_bytes = (uint32_t)ps.blob.size();
  d.stage_bytes = (uint32_t)(stage_end - o_blob);
  d.n_groups = ps.n_groups;
  d.off_groups = ps.off_groups, d.off_sphere = ps.off_sphere, d.off_moving = ps.off_moving;
  d.off_rect = ps.off_rect, d.off_triangle = ps.off_triangle, d.off_box = ps.off_box;
  d.off_trees = ps.off_trees, d.off_nodes = ps.off_nodes, d.off_tree_ids = ps.off_tree_ids, d.n_trees = ps.n_trees;
  d.off_planes = ps.off_planes, d.n_planes[0] = ps.n_planes[0], d.n_planes[1] = ps.n_planes[1], d.n_planes[2] = ps.n_planes[2];
  d.flat_extent = ps.flat_extent;
  d.n_objects = ps.n_objects;
  d.sphere_aux = reinterpret_cast<const SphereAux*>(ds->arena + o_saux);
  d.moving_aux = reinterpret_cast<const SphereAux*>(ds->arena + o_maux);
  d.rect_aux = reinterpret_cast<const ObjAux*>(ds->arena + o_raux);
  d.tri_aux = reinterpret_cast<const TriAux*>(ds->arena + o_taux);
  d.box_aux = reinterpret_cast<const ObjAux*>(ds->arena + o_baux);
  d.media = reinterpret_cast<const MediumRec*>(ds->arena + o_media);
  d.keys = reinterpret_cast<const int32_t*>(ds->arena + o_keys);
  for (int k = 0; k < 6; ++k) d.key_base[k] = key_base[k];
  d.materials = ds->arena + o_mat;
  d.textures = ds->arena + o_tex;
  d.texture_bytes = ds->arena + o_bytes;
  d.n_texture_texels = tex_bytes / 3;
  d.n_materials = (uint32_t)ps.materials.size();
  d.n_textures = (uint32_t)ps.textures.size();
  d.object_id = reinterpret_cast<const int32_t*>(ds->arena + o_objid);
  d.n_media_groups = ps.n_media_groups, d.n_flat_groups = ps.n_flat_groups, d.n_late_sphere_groups = ps.n_late_sphere_groups;
  d.off_sphere_box = ps.off_sphere_box, d.off_moving_box = ps.off_moving_box;
  d.n_sphere_chunks = (uint32_t)ps.sphere_chunk_open.size(), d.n_moving_chunks = (uint32_t)ps.moving_chunk_open.size();
  d.cull_bound[0] = d.cull_bound[1] = d.cull_bound[2] = 0.f;  // no culling until update_chunk_boxes()
  ds->geo.sphere_geo.swap(ps.sphere_geo), ds->geo.moving_geo.swap(ps.moving_geo);
  ds->geo.sphere_chunk_open.swap(ps.sphere_chunk_open), ds->geo.moving_chunk_open.swap(ps.moving_chunk_open);
  ds->queue_heads = reinterpret_cast<unsigned long long*>(ds->arena + o_heads);
  ds->counters = ds->queue_heads + kCounterSlots;
  ds->heavy_ctrl = reinterpret_cast<unsigned int*>(ds->arena + o_hctrl);
  ds->heavy_ready = reinterpret_cast<unsigned int*>(ds->arena + o_hready);
  ds->heavy_entries = reinterpret_cast<float*>(ds->arena + o_hentries);
  *out = ds;
  return PT_OK;
}

// The chunk boxes of moving spheres cover their sweep over the camera's shutter interval, so they are
// (re)computed when a render brings a different interval: a few hundred boxes, one small copy ordered
// before the launch on the same stream.  Renders of ONE device scene must therefore not overlap on
// different streams with different shutter intervals.
int update_chunk_boxes(pt_device_scene* scene, const pt_camera* cam, cudaStream_t st) {
  if (scene->cull_valid && scene->cull_mode == g_cull_enabled && std::memcmp(&scene->cull_time0, &cam->time0, 4) == 0 &&
      std::memcmp(&scene->cull_time1, &cam->time1, 4) == 0)
    return PT_OK;
  CullBoxes boxes;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  compute_cull_boxes(scene->geo, g_cull_enabled ? cam->time0 : nan, g_cull_enabled ? cam->time1 : nan, boxes);
  unsigned char* blob = const_cast<unsigned char*>(scene->desc.blob);
  if (!boxes.sphere.empty())
    PT_CUDA(cudaMemcpyAsync(blob + scene->desc.off_sphere_box, boxes.sphere.data(), boxes.sphere.size() * sizeof(float),
                            cudaMemcpyHostToDevice, st));
  if (!boxes.moving.empty())
    PT_CUDA(cudaMemcpyAsync(blob + scene->desc.off_moving_box, boxes.moving.data(), boxes.moving.size() * sizeof(float),
                            cudaMemcpyHostToDevice, st));
  for (int k = 0; k < 3; ++k) scene->desc.cull_bound[k] = boxes.bound[k];
  scene->cull_valid = true, scene->cull_mode = g_cull_enabled;
  scene->cull_time0 = cam->time0, scene->cull_time1 = cam->time1;
  return PT_OK;
}

int check_render_args(int width, int height, int spp, const pt_camera* cam, const pt_region* rg, const void* out) {
  if (!cam || !rg || !out) return fail(PT_ERR_INVALID_ARGUMENT, "render: null argument");
  if (width <= 0 || height <= 0 || spp <= 0) return fail(PT_ERR_INVALID_ARGUMENT, "render: width, height and spp must be positive");
  if (rg->w < 0 || rg->h < 0 || rg->y_stride <= 0 || rg->x0 < 0 || rg->y0 < 0 || rg->x0 + rg->w > width ||
      (rg->h > 0 && rg->y0 + (long long)(rg->h - 1) * rg->y_stride >= height))
    return fail(PT_ERR_INVALID_ARGUMENT, "render: region outside the image");
  return PT_OK;
}

// rows first, first+stride, ... of a `height`-row image
pt_region rows_of(int width, int height, int first, int stride) {
  pt_region r;
  r.x0 = 0, r.y0 = first, r.w = width, r.y_stride = stride;
  r.h = first < height ? (height - first + stride - 1) / stride : 0;
  return r;
}

}  // namespace

extern "C" {

const char* pt_last_error(void) { return g_error.c_str(); }
int pt_abi_version(void) { return PT_ABI_VERSION; }

int pt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int pt_set_num_gpus(int n) {
  if (n < 1) return fail(PT_ERR_INVALID_ARGUMENT, "pt_set_num_gpus: n must be >= 1");
  const int have = pt_device_count();
  if (have == 0) return fail(PT_ERR_NO_DEVICE, "no CUDA device available");
  if (n > have) return fail(PT_ERR_INVALID_ARGUMENT, "pt_set_num_gpus: more GPUs requested than visible");
  g_num_gpus = n;
  return PT_OK;
}
int pt_get_num_gpus(void) { return g_num_gpus; }

int pt_get_stats(pt_stats* out) {
  if (!out) return fail(PT_ERR_INVALID_ARGUMENT, "pt_get_stats: null argument");
  *out = g_stats;
  return PT_OK;
}

int pt_scene_upload(const pt_scene* scene, int device, pt_device_scene** out) {
  return upload(scene, device, out, nullptr, nullptr);
}

void pt_scene_free(pt_device_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();  // nothing may still be using the arena when it is handed to the next upload
  cached_free(s->device, s->arena, s->arena_bytes);
  cached_free(s->device, s->lpt_buf, s->lpt_ints * sizeof(int));
  cached_free(s->device, s->fb_stage, s->fb_stage_bytes);
  cached_free(s->device, s->tree_spill, s->tree_spill_bytes);
  delete s;
}

int pt_render_region_device(const pt_device_scene* cscene, int width, int height, int spp, int depth,
                            const pt_camera* camera, const pt_region* region, float* d_out, int64_t out_row_pitch,
                            void* stream) {
  return pt_render_resume_device(cscene, width, height, 0, spp, depth, camera, region, nullptr, 0, d_out, out_row_pitch, stream);
}

int pt_render_resume_device(const pt_device_scene* cscene, int width, int height, int spp_from, int spp, int depth,
                            const pt_camera* camera, const pt_region* region, float* d_state, int64_t state_row_pitch,
                            float* d_out, int64_t out_row_pitch, void* stream) {
  pt_device_scene* scene = const_cast<pt_device_scene*>(cscene);
  if (!scene) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render_region_device: null scene");
  const int rc = check_render_args(width, height, spp, camera, region, d_out);
  if (rc != PT_OK) return rc;
  if (spp_from < 0 || spp_from >= spp) return fail(PT_ERR_INVALID_ARGUMENT, "render: need 0 <= spp_from < spp_to");
  if (spp_from > 0 && !d_state) return fail(PT_ERR_INVALID_ARGUMENT, "render: resuming needs the per-pixel state");
  if (d_state && depth <= 0) return fail(PT_ERR_INVALID_ARGUMENT, "render: progressive rendering needs depth > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PT_CUDA(cudaSetDevice(scene->device));
  if (region->w == 0 || region->h == 0) return PT_OK;
  if (depth <= 0) {
    // render.hpp:58,91: no bounce allowed -> every sample returns black
    PT_CUDA(cudaMemset2DAsync(d_out, out_row_pitch * sizeof(float), 0, (size_t)region->w * 3 * sizeof(float), region->h, st));
    return PT_OK;
  }
  {
    const int crc = update_chunk_boxes(scene, camera, st);
    if (crc != PT_OK) return crc;
  }
  RenderParams p;
  p.scene = scene->desc;
  p.scene.flat_cull = g_cull_enabled ? 1u : 0u;
  p.cam = *camera;
  p.width = width, p.height = height, p.spp = spp, p.depth = depth;
  p.region = *region;
  p.out = d_out;
  p.out_row_pitch = out_row_pitch;
  p.out_pixel_floats = 3;
  // Vectorised framebuffer stores: finished pixels go to a float4 staging image with one 16-byte store each, and a
  // resolve pass packs them into the caller's rows with coalesced 16-byte stores.  Needs 16-byte aligned rows of a
  // multiple of 4 pixels; anything else keeps the scalar stores.
  const bool staged_fb = g_fb_stage_enabled && !g_single_task && region->w % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 &&
                         (out_row_pitch * (long long)sizeof(float)) % 16 == 0;
  if (staged_fb) {
    const size_t need = (size_t)region->w * region->h * 4 * sizeof(float);
    if (need > scene->fb_stage_bytes) {
      cached_free(scene->device, scene->fb_stage, scene->fb_stage_bytes);
      scene->fb_stage = nullptr, scene->fb_stage_bytes = 0;
      void* buf = nullptr;
      size_t got = 0;
      PT_CUDA(cached_malloc(scene->device, &buf, need, &got));
      scene->fb_stage = static_cast<float*>(buf), scene->fb_stage_bytes = got;
    }
    p.out = scene->fb_stage, p.out_row_pitch = 4ll * region->w, p.out_pixel_floats = 4;
  }
  p.tree_spill = nullptr, p.tree_spill_cap = 0;
  if (scene->desc.n_trees != 0u && g_kernel_kind == 0) {
    // a ray that runs along a mesh crosses hundreds of leaves: what the shared-memory tree lists cannot hold continues here
    const size_t need = (size_t)wave_grid(scene->device) * kTreeSpillLists * kTreeSpillCap * sizeof(uint2);
    if (need > scene->tree_spill_bytes) {
      cached_free(scene->device, scene->tree_spill, scene->tree_spill_bytes);
      scene->tree_spill = nullptr, scene->tree_spill_bytes = 0;
      void* buf = nullptr;
      size_t got = 0;
      PT_CUDA(cached_malloc(scene->device, &buf, need, &got));
      scene->tree_spill = static_cast<uint2*>(buf), scene->tree_spill_bytes = got;
    }
    p.tree_spill = scene->tree_spill, p.tree_spill_cap = kTreeSpillCap;
  }
  p.state = d_state, p.state_row_pitch = state_row_pitch, p.spp_from = spp_from;
  p.counters = scene->counters;
  p.team_size = g_team_size_override;  // 0 = chosen from the pixel count at launch
  p.kernel_kind = g_single_task ? 1 : g_kernel_kind;
  p.pool_cap = 0;
  p.scramble = 1;
  p.n_express = g_n_express;
  p.tuning = nullptr;
  p.express_positions = 0;
  p.order_mode = 0, p.tile_order = nullptr, p.tiles_x = p.tiles_y = 0, p.probe_cost = nullptr, p.n_positions = 0;
  p.heavy.ctrl = scene->heavy_ctrl, p.heavy.ready = scene->heavy_ready, p.heavy.entries = scene->heavy_entries;
  p.heavy.cap = kHeavyCap;
  auto next_queue_head = [&]() {
    const int slot = scene->next_slot;
    scene->next_slot = (slot + 2) % kCounterSlots;  // a pair: everybody's head and the express CTAs'
    return scene->queue_heads + slot;
  };
  PT_CUDA(cudaMemsetAsync(p.counters + 1, 0xff, 2 * sizeof(unsigned long long), st));
  PT_CUDA(cudaMemsetAsync(p.counters + 3, 0, 2 * sizeof(unsigned long long), st));
  PT_CUDA(cudaMemsetAsync(p.counters + 5, 0xff, sizeof(unsigned long long), st));
  PT_CUDA(cudaMemsetAsync(p.counters + 6, 0, 2 * sizeof(unsigned long long), st));  // [6] last CTA out, [7] pixels handed off
  PT_CUDA(cudaMemsetAsync(p.counters + 8, 0, 16 * sizeof(unsigned long long), st));  // hand-off service statistics

  // Longest-processing-time-first pixel order (wavefront kernel, images worth it): a cost probe traces
  // throw-away samples through every second pixel of every second row, the tiles are sorted by probed cost, and
  // the frame starts with the most expensive tiles so that the deepest pixels have the whole frame to finish and
  // the cheapest ones fill its end.  ONE probe sample per 64 samples of the frame (1 .. 8; 1/256 of the frame's work
  // at most): a pixel that is deep in some of its samples only (the mesh: isolated pixels averaging 24 scans a sample
  // where the mean is 2.3) shows in one probe sample with probability 1/2, and every such pixel that the order misses
  // starts late and ends the frame with its whole serial chain.
  const unsigned long long pixels = (unsigned long long)region->w * (unsigned long long)region->h;
  if (g_single_task) p.order_mode = 3;  // render.hpp:113-122: one generator for the whole image, x-major (pt_lane.cu)
  if (g_lpt_enabled && !g_single_task && pixels >= 32768ull && spp - spp_from >= 8) {
    const int pw = (region->w + kProbeStep - 1) / kProbeStep, ph = (region->h + kProbeStep - 1) / kProbeStep;
    const int tiles_x = (region->w + kTile - 1) / kTile, tiles_y = (region->h + kTile - 1) / kTile;
    const size_t need = (size_t)pw * ph + 2 * (size_t)tiles_x * tiles_y + sizeof(FrameTuning) / sizeof(int);
    if (need > scene->lpt_ints) {
      cached_free(scene->device, scene->lpt_buf, scene->lpt_ints * sizeof(int));
      scene->lpt_buf = nullptr, scene->lpt_ints = 0;
      void* buf = nullptr;
      size_t got = 0;
      PT_CUDA(cached_malloc(scene->device, &buf, need * sizeof(int), &got));
      scene->lpt_buf = static_cast<int*>(buf), scene->lpt_ints = got / sizeof(int);
    }
    int* probe_cost = scene->lpt_buf;
    int* tile_order = probe_cost + (size_t)pw * ph;
    int* scratch = tile_order + (size_t)tiles_x * tiles_y;
    FrameTuning* tuning = reinterpret_cast<FrameTuning*>(scratch + (size_t)tiles_x * tiles_y);
    RenderParams probe = p;
#ifdef PT_PROBE_SPP  // (experiments)
    const int probe_spp = PT_PROBE_SPP;
#else
    const int probe_spp = std::min(std::max((spp - spp_from) / 64, 1), 8);
#endif
    probe.order_mode = 2, probe.probe_cost = probe_cost, probe.spp = probe_spp, probe.spp_from = 0, probe.state = nullptr, probe.counters = nullptr;
    probe.kernel_kind = 0;  // the probe always runs on the wavefront kernel
    probe.pixel_counter = next_queue_head();
    probe.heavy.stamp = ++scene->launch_stamp;
    PT_CUDA(cudaMemsetAsync(scene->heavy_ctrl + 2, 0, sizeof(unsigned), st));
    PT_CUDA(cudaMemsetAsync(probe.pixel_counter, 0, 2 * sizeof(unsigned long long), st));
    cudaError_t pe = launch_render(probe, scene->device, 0, st, nullptr);
    if (pe != cudaSuccess) return cuda_fail(pe, "cost probe launch");
    pe = launch_tile_order(probe_cost, region->w, region->h, tiles_x, tiles_y, tile_order, scratch, tuning, wave_grid(scene->device),
                           g_n_express, probe_spp, st);
    if (pe != cudaSuccess) return cuda_fail(pe, "tile order launch");
    p.order_mode = 1, p.tile_order = tile_order, p.tiles_x = tiles_x, p.tiles_y = tiles_y;
    p.tuning = g_kernel_kind == 0 ? tuning : nullptr;
    scene->last_tuning = tuning;
    scene->kernel_launches += 2;
  }
  p.pixel_counter = next_queue_head();
  p.heavy.stamp = ++scene->launch_stamp;
  PT_CUDA(cudaMemsetAsync(scene->heavy_ctrl + 2, 0, sizeof(unsigned), st));
  PT_CUDA(cudaMemsetAsync(p.pixel_counter, 0, 2 * sizeof(unsigned long long), st));
  cudaError_t e = launch_render(p, scene->device, 0, st, &scene->last_launch);
  if (e != cudaSuccess) return cuda_fail(e, "render kernel launch");
  scene->kernel_launches += 1;
  if (staged_fb) {
    e = launch_resolve_fb(scene->fb_stage, region->w, region->h, d_out, out_row_pitch, st);
    if (e != cudaSuccess) return cuda_fail(e, "framebuffer resolve launch");
    scene->kernel_launches += 1;
  }
  scene->paths_launched += (unsigned long long)region->w * region->h * (unsigned long long)(spp - spp_from);
  return PT_OK;
}

int pt_scene_read_counters(pt_device_scene* scene, uint64_t* paths, uint64_t* scans, int reset) {
  if (!scene) return fail(PT_ERR_INVALID_ARGUMENT, "pt_scene_read_counters: null scene");
  PT_CUDA(cudaSetDevice(scene->device));
  PT_CUDA(cudaDeviceSynchronize());
  unsigned long long all[5] = { 0, 0, 0, 0, 0 };
  PT_CUDA(cudaMemcpy(all, scene->counters, sizeof all, cudaMemcpyDeviceToHost));
  const unsigned long long v = all[0];
  if (all[4] != 0) return fail(PT_ERR_CUDA, "render kernel: the heavy-pixel hand-off queue timed out (internal error)");
  if (paths) *paths = scene->paths_launched;
  if (scans) *scans = v;
  if (reset) {
    PT_CUDA(cudaMemset(scene->counters, 0, sizeof(unsigned long long)));
    scene->paths_launched = 0;
  }
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): force the launch team size (0 = automatic).
// Test hook (host only, no GPU needed): the chunk layout and the chunk boxes pack_scene / compute_cull_boxes produce.
// keys: one per sphere element, static elements first (-1 - order index; INT_MIN = padding); boxes: kCullSets sets of
// (static chunks, then moving chunks) x {lo[3], hi[3]}.  Returns the number of floats boxes needs (fills up to cap).
int pt_debug_chunk_layout(const pt_scene* scene, float cam_time0, float cam_time1, int* n_static_elements,
                          int* n_moving_elements, int* keys, int keys_cap, float* boxes, int boxes_cap, float bounds[3]) {
  if (!scene) return fail(PT_ERR_INVALID_ARGUMENT, "pt_debug_chunk_layout: null scene");
  PackedScene ps;
  std::string err;
  const int rc = pack_scene(*scene, ps, err);
  if (rc != PT_OK) return fail(rc, err);
  CullBoxes cb;
  compute_cull_boxes(ps, cam_time0, cam_time1, cb);
  *n_static_elements = (int)ps.sphere_aux.size(), *n_moving_elements = (int)ps.moving_aux.size();
  int k = 0;
  for (const auto& a : ps.sphere_aux)
    if (k < keys_cap) keys[k++] = a.key;
  for (const auto& a : ps.moving_aux)
    if (k < keys_cap) keys[k++] = a.key;
  const size_t ns = ps.sphere_chunk_open.size(), nm = ps.moving_chunk_open.size();
  int at = 0;
  for (int set = 0; set < kCullSets; ++set)
    for (size_t c = 0; c < ns + nm; ++c) {
      const float* b = c < ns ? cb.sphere.data() + ((size_t)set * ns + c) * 8 : cb.moving.data() + ((size_t)set * nm + (c - ns)) * 8;
      for (int j = 0; j < 6; ++j, ++at)
        if (at < boxes_cap) boxes[at] = b[j < 3 ? j : j + 1];
    }
  for (int j = 0; j < 3; ++j) bounds[j] = cb.bound[j];
  return at;
}

int pt_debug_set_cull(int on) {
  g_cull_enabled = on ? 1 : 0;
  return PT_OK;
}
int pt_debug_set_team_size(int t) {
  g_team_size_override = t;
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): 0 = wavefront kernel (default), 1 = lane kernel.
int pt_debug_set_kernel(int kind) {
  g_kernel_kind = kind;
  return PT_OK;
}

// Test hook (not part of pt_abi.h): the closest-hit scan of n host rays {o, d, time} with generator states `seeds`
// (pt_kernel.h: launch_probe_rays); the chunk boxes are those of `camera`'s shutter interval.
int pt_debug_closest_hit(pt_device_scene* scene, const pt_camera* camera, int n, const float* rays7, const uint32_t* seeds, int mode,
                         float* out_t, int32_t* out_index, uint32_t* out_rng) {
  if (!scene || !camera || n < 0 || (n && (!rays7 || !seeds || !out_t || !out_index || !out_rng)))
    return fail(PT_ERR_INVALID_ARGUMENT, "pt_debug_closest_hit: bad argument");
  if (n == 0) return PT_OK;
  PT_CUDA(cudaSetDevice(scene->device));
  const int crc = update_chunk_boxes(scene, camera, nullptr);
  if (crc != PT_OK) return crc;
  unsigned char* buf = nullptr;
  const size_t nn = (size_t)n;
  PT_CUDA(cudaMalloc(&buf, nn * (7 + 1 + 1 + 1 + 1) * 4));
  float* d_rays = reinterpret_cast<float*>(buf);
  uint32_t* d_seeds = reinterpret_cast<uint32_t*>(d_rays + 7 * nn);
  float* d_t = reinterpret_cast<float*>(d_seeds + nn);
  int32_t* d_index = reinterpret_cast<int32_t*>(d_t + nn);
  uint32_t* d_rng = reinterpret_cast<uint32_t*>(d_index + nn);
  SceneDesc desc = scene->desc;
  desc.flat_cull = g_cull_enabled ? 1u : 0u;
  cudaError_t e = cudaMemcpy(d_rays, rays7, nn * 28, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_seeds, seeds, nn * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_probe_rays(desc, n, d_rays, d_seeds, mode, d_t, d_index, d_rng, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out_t, d_t, nn * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_index, d_index, nn * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_rng, d_rng, nn * 4, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (e != cudaSuccess) return cuda_fail(e, "pt_debug_closest_hit");
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): what the last cost probe on this scene said: {heavy rate, express CTAs, mean scans per
// sample x 1000, share of the work in heavy pixels x 1000, deepest probed sample}; zeros when no probe has run.
int pt_debug_frame_tuning(pt_device_scene* scene, int out[5]) {
  for (int k = 0; k < 5; ++k) out[k] = 0;
  if (!scene || !scene->last_tuning) return PT_OK;
  PT_CUDA(cudaSetDevice(scene->device));
  PT_CUDA(cudaDeviceSynchronize());
  FrameTuning t;
  PT_CUDA(cudaMemcpy(&t, scene->last_tuning, sizeof t, cudaMemcpyDeviceToHost));
  out[0] = t.heavy_rate, out[1] = t.n_express, out[2] = t.mean_scans_x1000, out[3] = t.heavy_share_x1000, out[4] = t.max_scans;
  return PT_OK;
}

// Test hook (not part of pt_abi.h): the device's transcendentals on host arrays (pt_kernel.h: launch_probe_math).
int pt_debug_math(int kind, int n, const float* in, float* out) {
  if (n < 0 || (n && (!in || !out))) return fail(PT_ERR_INVALID_ARGUMENT, "pt_debug_math: bad argument");
  if (n == 0) return PT_OK;
  const size_t n_in = (size_t)n * (kind == 4 ? 2 : 1);
  float* d = nullptr;
  PT_CUDA(cudaMalloc(&d, (n_in + (size_t)n) * sizeof(float)));
  cudaError_t e = cudaMemcpy(d, in, n_in * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_probe_math(kind, n, d, d + n_in, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out, d + n_in, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "pt_debug_math");
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): staged + resolved framebuffer stores on / off.
int pt_debug_set_fb_stage(int on) {
  g_fb_stage_enabled = on;
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): switch the LPT pixel ordering (cost probe + tile sort) on / off.
int pt_debug_set_lpt(int on) {
  g_lpt_enabled = on;
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): number of express CTAs of the wavefront kernel (< 0 = automatic).
int pt_debug_set_express(int n) {
  g_n_express = n;
  return PT_OK;
}

// Debug aid (not part of pt_abi.h): timeline of the LAST launch on this scene, in ns:
// out[0] = queue-dry - start, out[1] = last-warp-retired - start.
// out[5..9]: hand-off service: rounds, ray-rounds, total and longest wait in the queue (ns), longest stay (rounds);
// out[10]: (ray, chunk) items that did not fit the item list and were scanned in place
// Tree-list items of the last launch that continued in global memory (RenderParams::tree_spill).
int pt_debug_tree_spilled(pt_device_scene* scene, unsigned long long* out) {
  PT_CUDA(cudaSetDevice(scene->device));
  PT_CUDA(cudaDeviceSynchronize());
  PT_CUDA(cudaMemcpy(out, scene->counters + 10, sizeof *out, cudaMemcpyDeviceToHost));
  return PT_OK;
}
int pt_debug_timeline(pt_device_scene* scene, unsigned long long out[11]) {
  PT_CUDA(cudaSetDevice(scene->device));
  PT_CUDA(cudaDeviceSynchronize());
  unsigned long long v[16];
  PT_CUDA(cudaMemcpy(v, scene->counters, sizeof v, cudaMemcpyDeviceToHost));
  unsigned int ctrl[4];
  PT_CUDA(cudaMemcpy(ctrl, scene->heavy_ctrl, sizeof ctrl, cudaMemcpyDeviceToHost));
  out[0] = v[2] - v[1], out[1] = v[3] - v[1];
  out[2] = v[5] - v[1], out[3] = v[6] - v[1];  // first / last CTA out of regular work
  out[4] = v[7];                               // heavy pixels handed to the express lane
  out[5] = v[8], out[6] = v[9], out[7] = v[12], out[8] = v[13], out[9] = v[14], out[10] = v[15];
  if (std::getenv("PT_LAST_PIXEL")) {  // -DPT_LAST_PIXEL builds
    unsigned long long w = 0;
    PT_CUDA(cudaMemcpy(&w, scene->counters + 16, sizeof w, cudaMemcpyDeviceToHost));
    unsigned long long n_in_order = 0;
    PT_CUDA(cudaMemcpy(&n_in_order, scene->counters + 17, sizeof n_in_order, cudaMemcpyDeviceToHost));
    std::fprintf(stderr, "rays stored that ask for the vector-order scan: %llu\n", n_in_order);
    std::fprintf(stderr, "last pixel: finished %.2f ms after the start, own %llu, service mode %llu, about %llu scans (rounds, if taken over)\n",
                 (double)(w >> 24) / 1e6, (w >> 23) & 1ull, (w >> 22) & 1ull, (w & ((1ull << 22) - 1)) << 2);
  }
  if (const char* env = std::getenv("PT_PHASE_TIMING")) {  // debug builds (-DPT_PHASE_TIMING): cycles per phase
    (void)env;
    std::fprintf(stderr, "desc: groups %u media %u flat %u sphere chunks %u moving chunks %u blob %u B\n", scene->desc.n_groups,
                 scene->desc.n_media_groups, scene->desc.n_flat_groups, scene->desc.n_sphere_chunks, scene->desc.n_moving_chunks,
                 scene->desc.blob_bytes);
    unsigned long long ph[8];
    PT_CUDA(cudaMemcpy(ph, scene->counters + 16, sizeof ph, cudaMemcpyDeviceToHost));
    std::fprintf(stderr, "phase cycles per round: boxes %.0f spheres %.0f flat %.0f sort %.0f shade %.0f; %llu rounds, %.1f rays, %.1f items\n",
                 (double)ph[0] / (double)std::max(ph[5], 1ull), (double)ph[1] / (double)std::max(ph[5], 1ull),
                 (double)ph[2] / (double)std::max(ph[5], 1ull), (double)ph[3] / (double)std::max(ph[5], 1ull),
                 (double)ph[4] / (double)std::max(ph[5], 1ull), ph[5], (double)ph[6] / (double)std::max(ph[5], 1ull), (double)ph[7] / (double)std::max(ph[5], 1ull));
  }
  return PT_OK;
}

int pt_scene_launch_count(const pt_device_scene* scene, uint64_t* launches) {
  if (!scene || !launches) return fail(PT_ERR_INVALID_ARGUMENT, "pt_scene_launch_count: null argument");
  *launches = scene->kernel_launches;
  return PT_OK;
}

// ---- blocking host-buffer path (what render<W,H,S>() of render.hpp:141-160 becomes)
int pt_render_region(int width, int height, int spp, int depth, const pt_camera* camera, const pt_scene* hitables,
                     const pt_region* region, float* out, int64_t out_row_pitch) {
  if (!hitables) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render_region: null scene");
  int rc = check_render_args(width, height, spp, camera, region, out);
  if (rc != PT_OK) return rc;
  pt_stats st {};
  st.n_gpus = 1;
  pt_device_scene* ds = nullptr;
  rc = upload(hitables, 0, &ds, &st.h2d_ms, &st.h2d_bytes);
  if (rc != PT_OK) return rc;
  const size_t row_floats = (size_t)region->w * 3;
  float* d_out = nullptr;
  size_t d_out_bytes = 0;
  void* d_out_v = nullptr;
  cudaError_t e = cached_malloc(0, &d_out_v, std::max<size_t>(row_floats * region->h, 1) * sizeof(float), &d_out_bytes);
  d_out = static_cast<float*>(d_out_v);
  if (e != cudaSuccess) {
    pt_scene_free(ds);
    return cuda_fail(e, "cudaMalloc(framebuffer)");
  }
  cudaEvent_t ev[4];
  for (auto& x : ev) cudaEventCreate(&x);
  cudaEventRecord(ev[0], 0);
  rc = pt_render_region_device(ds, width, height, spp, depth, camera, region, d_out, (int64_t)row_floats, nullptr);
  cudaEventRecord(ev[1], 0);
  if (rc == PT_OK) {
    cudaEventRecord(ev[2], 0);
    e = cudaMemcpy2DAsync(out, out_row_pitch * sizeof(float), d_out, row_floats * sizeof(float),
                          row_floats * sizeof(float), region->h, cudaMemcpyDeviceToHost, 0);
    cudaEventRecord(ev[3], 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) rc = cuda_fail(e, "framebuffer download");
  }
  if (rc == PT_OK) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    st.kernel_ms = ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]);
    st.d2h_ms = ms;
    st.d2h_bytes = row_floats * region->h * sizeof(float);
    st.kernel_launches = ds->kernel_launches;
    uint64_t paths = 0, scans = 0;
    rc = pt_scene_read_counters(ds, &paths, &scans, 0);  // (also where a device-side error flag becomes an error code)
    st.paths = (uint64_t)region->w * region->h * spp, st.scans = scans;
    if (rc == PT_OK) g_stats = st;
  }
  for (auto& x : ev) cudaEventDestroy(x);
  cudaStreamSynchronize(0);
  cached_free(0, d_out, d_out_bytes);
  pt_scene_free(ds);
  return rc;
}

int pt_render_resume(int width, int height, int spp_from, int spp_to, int depth, const pt_camera* camera, const pt_scene* hitables,
                     const pt_region* region, float* state, float* out, int64_t out_row_pitch) {
  if (!hitables || !state) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render_resume: null argument");
  int rc = check_render_args(width, height, spp_to, camera, region, out);
  if (rc != PT_OK) return rc;
  if (region->w == 0 || region->h == 0) return PT_OK;
  pt_device_scene* ds = nullptr;
  rc = upload(hitables, 0, &ds, nullptr, nullptr);
  if (rc != PT_OK) return rc;
  const size_t n_pixels = (size_t)region->w * region->h;
  float *d_state = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_state, n_pixels * 7 * sizeof(float));
  if (e != cudaSuccess) {
    pt_scene_free(ds);
    return cuda_fail(e, "cudaMalloc(progressive state)");
  }
  d_out = d_state + n_pixels * 4;
  if (spp_from > 0) e = cudaMemcpy(d_state, state, n_pixels * 4 * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = pt_render_resume_device(ds, width, height, spp_from, spp_to, depth, camera, region, d_state, region->w, d_out,
                                 (int64_t)region->w * 3, nullptr);
    if (rc == PT_OK) e = cudaMemcpy(state, d_state, n_pixels * 4 * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == PT_OK && e == cudaSuccess)
      e = cudaMemcpy2D(out, out_row_pitch * sizeof(float), d_out, (size_t)region->w * 3 * sizeof(float), (size_t)region->w * 3 * sizeof(float),
                       region->h, cudaMemcpyDeviceToHost);
  }
  if (rc == PT_OK && e != cudaSuccess) rc = cuda_fail(e, "pt_render_resume");
  if (rc == PT_OK) rc = pt_scene_read_counters(ds, nullptr, nullptr, 0);
  cudaFree(d_state);
  pt_scene_free(ds);
  return rc;
}

// The reference built with -DUSE_SINGLE_TASK (render.hpp:113-122).  One warp on GPU 0: see pt_lane.cu.
int pt_render_single_task(int width, int height, int spp, int depth, const pt_camera* camera, const pt_scene* hitables, float* fb) {
  if (!hitables || !camera || !fb) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render_single_task: null argument");
  if (width <= 0 || height <= 0 || spp <= 0) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render_single_task: width, height and spp must be positive");
  const pt_region full = rows_of(width, height, 0, 1);
  g_single_task = true;
  const int rc = pt_render_region(width, height, spp, depth, camera, hitables, &full, fb, (int64_t)width * 3);
  g_single_task = false;
  return rc;
}

int pt_render(int width, int height, int spp, int depth, const pt_camera* camera, const pt_scene* hitables, float* fb) {
  if (!hitables || !camera || !fb) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render: null argument");
  if (width <= 0 || height <= 0 || spp <= 0) return fail(PT_ERR_INVALID_ARGUMENT, "pt_render: width, height and spp must be positive");
  const int n = g_num_gpus;
  if (n == 1) {
    const pt_region full = rows_of(width, height, 0, 1);
    return pt_render_region(width, height, spp, depth, camera, hitables, &full, fb, (int64_t)width * 3);
  }
  // ---- multi-GPU, one process: rows interleaved over the GPUs (row y -> GPU
  // y mod n); every GPU stores its pixels straight into GPU 0's framebuffer
  // through peer-mapped memory over NVLink.  Without peer access the rows are
  // rendered into a local buffer and copied into place afterwards.
  if (pt_device_count() < n) return fail(PT_ERR_NO_DEVICE, "pt_render: fewer CUDA devices than pt_set_num_gpus()");
  pt_stats st {};
  st.n_gpus = (uint32_t)n;
  std::vector<pt_device_scene*> scenes(n, nullptr);
  std::vector<float*> local(n, nullptr);
  std::vector<cudaStream_t> streams(n, nullptr);
  std::vector<cudaEvent_t> ev0(n, nullptr), ev1(n, nullptr);
  std::vector<char> peer_enabled(n, 0);
  float* fb0 = nullptr;
  int rc = PT_OK;
  const size_t row_floats = (size_t)width * 3;
  auto cleanup = [&]() {
    for (int d = 0; d < n; ++d) {
      cudaSetDevice(d);
      if (ev0[d]) cudaEventDestroy(ev0[d]);
      if (ev1[d]) cudaEventDestroy(ev1[d]);
      if (local[d]) cudaFree(local[d]);
      if (streams[d]) cudaStreamDestroy(streams[d]);
      if (scenes[d]) pt_scene_free(scenes[d]);
      if (peer_enabled[d]) cudaDeviceDisablePeerAccess(0);  // (only what this call enabled)
    }
    cudaSetDevice(0);
    if (fb0) cudaFree(fb0);
  };
  cudaSetDevice(0);
  cudaError_t e = cudaMalloc(&fb0, row_floats * height * sizeof(float));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(framebuffer)");
  for (int d = 0; d < n && rc == PT_OK; ++d) {
    rc = upload(hitables, d, &scenes[d], &st.h2d_ms, &st.h2d_bytes);
    if (rc != PT_OK) break;
    cudaSetDevice(d);
    cudaStreamCreateWithFlags(&streams[d], cudaStreamNonBlocking);
    cudaEventCreate(&ev0[d]), cudaEventCreate(&ev1[d]);
  }
  for (int d = 0; d < n && rc == PT_OK; ++d) {
    cudaSetDevice(d);
    const pt_region rg = rows_of(width, height, d, n);
    float* target = fb0 + (size_t)d * row_floats;
    int64_t pitch = (int64_t)n * (int64_t)row_floats;
    if (d != 0) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, d, 0);
      if (can) {
        e = cudaDeviceEnablePeerAccess(0, 0);
        if (e == cudaSuccess) peer_enabled[d] = 1;
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        cudaGetLastError();
      }
      if (!can) {
        e = cudaMalloc(&local[d], std::max<size_t>(row_floats * rg.h, 1) * sizeof(float));
        if (e != cudaSuccess) {
          rc = cuda_fail(e, "cudaMalloc(local rows)");
          break;
        }
        target = local[d], pitch = (int64_t)row_floats;
      }
    }
    cudaEventRecord(ev0[d], streams[d]);
    rc = pt_render_region_device(scenes[d], width, height, spp, depth, camera, &rg, target, pitch, streams[d]);
    cudaEventRecord(ev1[d], streams[d]);
    if (rc == PT_OK && local[d])
      cudaMemcpy2DAsync(fb0 + (size_t)d * row_floats, (size_t)n * row_floats * sizeof(float), local[d],
                        row_floats * sizeof(float), row_floats * sizeof(float), rg.h, cudaMemcpyDefault, streams[d]);
  }
  for (int d = 0; d < n && rc == PT_OK; ++d) {
    cudaSetDevice(d);
    e = cudaStreamSynchronize(streams[d]);
    if (e != cudaSuccess) rc = cuda_fail(e, "multi-GPU render");
    float ms = 0.f;
    if (rc == PT_OK && cudaEventElapsedTime(&ms, ev0[d], ev1[d]) == cudaSuccess) st.kernel_ms = std::max<double>(st.kernel_ms, ms);
    uint64_t scans = 0;
    if (rc == PT_OK) rc = pt_scene_read_counters(scenes[d], nullptr, &scans, 0);
    st.scans += scans;
  }
  if (rc == PT_OK) {
    cudaSetDevice(0);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    e = cudaMemcpyAsync(fb, fb0, row_floats * height * sizeof(float), cudaMemcpyDeviceToHost, 0);
    cudaEventRecord(b, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) rc = cuda_fail(e, "framebuffer download");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    st.d2h_ms = ms, st.d2h_bytes = row_floats * height * sizeof(float);
    cudaEventDestroy(a), cudaEventDestroy(b);
    st.paths = (uint64_t)width * height * spp;
    for (int d = 0; d < n; ++d) st.kernel_launches += scenes[d]->kernel_launches;
    g_stats = st;
  }
  cleanup();
  return rc;
}

int render(int width, int height, int spp, int depth, const pt_camera* camera, const pt_scene* hitables, float* fb) {
  return pt_render(width, height, spp, depth, camera, hitables, fb);
}

// ---- framebuffer sharing between the one-process-per-GPU ranks --------------
int pt_fb_alloc(int device, size_t bytes, float** d_ptr) {
  if (!d_ptr) return fail(PT_ERR_INVALID_ARGUMENT, "pt_fb_alloc: null argument");
  PT_CUDA(cudaSetDevice(device));
  PT_CUDA(cudaMalloc(d_ptr, std::max<size_t>(bytes, 16)));
  return PT_OK;
}
int pt_fb_free(int device, float* d_ptr) {
  PT_CUDA(cudaSetDevice(device));
  PT_CUDA(cudaFree(d_ptr));
  return PT_OK;
}
int pt_fb_export(float* d_ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  PT_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle, &h, 64);
  return PT_OK;
}
int pt_fb_open(int device, const unsigned char handle[64], float** d_ptr) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  PT_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  PT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = static_cast<float*>(p);
  return PT_OK;
}
int pt_fb_close(float* d_ptr) {
  PT_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return PT_OK;
}

}  // extern "C"

// ---- FP32 roofline denominator: register-resident FFMA throughput -----------
namespace {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = __fmaf_rn(x0, a, b), x1 = __fmaf_rn(x1, a, b), x2 = __fmaf_rn(x2, a, b), x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b), x5 = __fmaf_rn(x5, a, b), x6 = __fmaf_rn(x6, a, b), x7 = __fmaf_rn(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

extern "C" int pt_measure_fp32_peak(int device, double* tflops, double* sm_mhz_est) {
  if (!tflops) return fail(PT_ERR_INVALID_ARGUMENT, "pt_measure_fp32_peak: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PT_ERR_NO_DEVICE, "no CUDA device available");
  PT_CUDA(cudaSetDevice(device));
  int sms = 0;
  PT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int grid = sms * 8, block = 256, iters = 4096;
  float* d = nullptr;
  PT_CUDA(cudaMalloc(&d, (size_t)grid * block * sizeof(float)));
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  double best_ms = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, 0);
    ffma_peak_kernel<<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(b, 0);
    cudaError_t e = cudaEventSynchronize(b);
    if (e != cudaSuccess) {
      cudaFree(d);
      return cuda_fail(e, "ffma_peak_kernel");
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0) best_ms = std::min<double>(best_ms, ms);
  }
  cudaEventDestroy(a), cudaEventDestroy(b);
  cudaFree(d);
  const double fmas = (double)grid * block * (double)iters * 16.0 * 8.0;
  *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
  if (sm_mhz_est) *sm_mhz_est = fmas / (best_ms * 1e-3) / ((double)sms * 128.0) / 1e6;
  return PT_OK;
}
