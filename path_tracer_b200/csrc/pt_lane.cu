// pt_lane.cu -- the lane kernel (pt_debug_set_kernel(1)): persistent warps, a pixel per team of lanes.
//
// A warp holds k = 32 / T pixels, each owned by a TEAM of T lanes that carry the pixel's path state
// REPLICATED in registers; the scan is split inside a team and merged with shuffles, bit-identical for
// every T.  The simpler scheduler: the reference point in the profiles, and (closest_hit, pt_prims.cuh)
// the scan of the wavefront kernel's sequential fallback.  Same device functions as pt_wave.cu, same bits.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_abi.h"
#include "pt_device.cuh"
#include "pt_kernel.h"
#include "pt_packed.h"
#include "pt_prims.cuh"
#include "pt_shade.cuh"

namespace ptb {

// ---------------------------------------------------------------- the lane loop
// One warp, k = 32 / team_size pixels at a time, the path state of each pixel replicated in the
// registers of its team (see the header comment); pixels come from the pixel queue.
template <bool kSmem>
PT_DEV void lane_loop(const RenderParams& p, const SceneDesc& sc, const SceneView& sv, int team_size0,
                      unsigned int& n_scans) {
  const pt_camera& cam = p.cam;
  const float fwidth = (float)p.width, fheight = (float)p.height, fspp = (float)p.spp;

  // per-lane path state, replicated across the lanes of a team
  const int lane = (int)(threadIdx.x & 31u);
  int team_size = team_size0;         // lanes per pixel (power of two); grows when the warp is re-packed
  int member = lane & (team_size - 1);
  bool live = false;                  // owns a pixel
  bool need_path = true;              // must start a new camera sample
  bool exhausted_queue = false;       // the pixel queue has run dry
  int px = 0, py = 0;                 // global pixel coordinates
  float* out_px = nullptr;
  float* state_px = nullptr;
  int sample = p.spp;
  int bounce = 0;
  Rng rng { 0u };
  Ray ray { v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f), 0.f };
  V3 att = v3(1.f, 1.f, 1.f);
  V3 acc = v3(0.f, 0.f, 0.f);
  int pix_scans = 0;  // closest-hit scans spent on the current pixel

  for (;;) {
    // ---- (R) re-pack: the queue is dry for this warp and at most half of its teams still own a
    // pixel -> move the survivors into teams twice (or more) as large.
    if (team_size < kSphereChunk) {
      const unsigned can_fetch = __ballot_sync(0xffffffffu, !live && !exhausted_queue);
      const unsigned leaders = __ballot_sync(0xffffffffu, live && member == 0);
      const int k_live = __popc(leaders);
      if (can_fetch == 0u && k_live > 0 && 2 * k_live * team_size <= 32) {
        int new_size = team_size;
        while (2 * k_live * new_size <= 32 && new_size < kSphereChunk) new_size <<= 1;
        const int new_team = lane / new_size;
        const bool keep = new_team < k_live;
        const int src = keep ? (int)__fns(leaders, 0u, new_team + 1) : lane;  // leader lane of the new_team-th live team
#define PT_MOVE(x) x = __shfl_sync(0xffffffffu, x, src)
        PT_MOVE(px), PT_MOVE(py), PT_MOVE(sample), PT_MOVE(bounce), PT_MOVE(rng.s), PT_MOVE(pix_scans);
        PT_MOVE(ray.o.x), PT_MOVE(ray.o.y), PT_MOVE(ray.o.z), PT_MOVE(ray.d.x), PT_MOVE(ray.d.y), PT_MOVE(ray.d.z);
        PT_MOVE(ray.tm), PT_MOVE(att.x), PT_MOVE(att.y), PT_MOVE(att.z), PT_MOVE(acc.x), PT_MOVE(acc.y), PT_MOVE(acc.z);
        unsigned long long optr = (unsigned long long)out_px;
        PT_MOVE(optr);
        out_px = (float*)optr;
        optr = (unsigned long long)state_px;
        PT_MOVE(optr);
        state_px = (float*)optr;
        int np = need_path ? 1 : 0;
        PT_MOVE(np);
        need_path = np != 0;
#undef PT_MOVE
        live = keep;
        exhausted_queue = true;
        team_size = new_size;
        member = lane & (team_size - 1);
      }
    }

    // ---- (A) path regeneration: render.hpp:94-105 sample loop, :130-133 seeding
    if (need_path && live && sample == p.spp) {
      // final_color /= samples; fb[y][x] = final_color (render.hpp:102-105); one writer per team
      if (member == 0) pixel_finish(p, out_px, state_px, acc, rng, fspp);
      live = false;
    }
    {
      const bool wants = need_path && !live && !exhausted_queue;
      if (__any_sync(0xffffffffu, wants)) {
        unsigned long long idx = 0ull;
        if (wants && member == 0) {  // the team leader pulls the next pixel (skipping tile positions outside the region)
          int tx, ty;
          float *tp, *ts;
          do idx = atomicAdd(p.pixel_counter, 1ull);
          while (idx < p.n_positions && !queue_pixel(p, idx, tx, ty, tp, ts));
        }
        idx = __shfl_sync(0xffffffffu, idx, lane - member);
        if (wants) {
          if (idx < p.n_positions) {
            queue_pixel(p, idx, px, py, out_px, state_px);
            pixel_start(p, px, py, state_px, rng, acc, sample);
            pix_scans = 0;
            live = true;
          } else {
            exhausted_queue = true;
            if (p.counters && member == 0) atomicMin(p.counters + 2, globaltimer_ns());  // timeline: queue ran dry
          }
        }
      }
    }
    if (need_path && live) {
      camera_ray(cam, px, py, fwidth, fheight, rng, ray);
      att = v3(1.f, 1.f, 1.f);
      bounce = 0;
      need_path = false;
    }
    if (!__any_sync(0xffffffffu, live)) {
      if (__all_sync(0xffffffffu, exhausted_queue)) break;
      continue;
    }

    // ---- (B) closest hit: render.hpp:60 -> :30-51
    const Best best = closest_hit<kSmem>(sc, sv, ray, rng, live, member, team_size);

    // ---- (C) shade: render.hpp:58-91 (every member of a team computes the same thing)
    if (live) {
      if (member == 0) ++n_scans;
      ++pix_scans;
      V3 contribution;
      if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
        acc = vadd(acc, contribution);
        ++sample;
        need_path = true;
      }
    }
    // A warp that finds itself holding one of the image's deepest pixels stops taking new pixels: as its
    // other pixels finish it is re-packed into ever larger teams, until all 32 lanes scan for the deep
    // pixel and its remaining thousands of bounces take microseconds each instead of a full round.
    if (__any_sync(0xffffffffu, live && pix_scans > kDeepBase + kDeepRate * (sample - p.spp_from))) exhausted_queue = true;
  }

}

// ---------------------------------------------------------------- USE_SINGLE_TASK (render.hpp:113-122)
// The reference's FPGA-style executor: ONE generator with the default seed for the whole image, pixels visited x-major,
// every sample continuing the stream where the previous one left it.  The chain is strictly serial -- width x height x
// spp samples, each consuming a data-dependent number of draws -- so all a GPU can offer is one warp: a team of 16 lanes
// that splits every closest-hit scan (the other 16 lanes idle along in the shuffles).  Bit-identical to the reference
// built with -DUSE_SINGLE_TASK; about 6 us per bounce, i.e. slower than one host core.  It exists so that a caller of
// that mode has a drop-in, not because it is fast.
template <bool kSmem>
PT_DEV void single_task_loop(const RenderParams& p, const SceneDesc& sc, const SceneView& sv, unsigned int& n_scans) {
  const int lane = (int)(threadIdx.x & 31u);
  const bool act = lane < kSphereChunk;
  const int member = lane & (kSphereChunk - 1);
  const float fwidth = (float)p.width, fheight = (float)p.height;
  Rng rng { 2463534242u };  // LocalPseudoRNG's default seed (rtweekend.hpp:35, xorshift.hpp)
  for (int x = 0; x != p.width; ++x)
    for (int y = 0; y != p.height; ++y) {
      V3 acc = v3(0.f, 0.f, 0.f);
      for (int s = 0; s < p.spp; ++s) {  // render.hpp:95-101
        Ray ray;
        camera_ray(p.cam, x, y, fwidth, fheight, rng, ray);
        V3 att = v3(1.f, 1.f, 1.f);
        int bounce = 0;
        for (;;) {
          const Best best = closest_hit<kSmem>(sc, sv, ray, rng, act, member, kSphereChunk);
          if (lane == 0) ++n_scans;
          V3 contribution;
          if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
            acc = vadd(acc, contribution);
            break;
          }
        }
      }
      if (lane == 0) {  // render.hpp:102-105
        float* out = p.out + (long long)y * p.out_row_pitch + 3ll * x;
        const V3 fin = vdivs(acc, (float)p.spp);
        out[0] = fin.x, out[1] = fin.y, out[2] = fin.z;
      }
    }
}

// ---------------------------------------------------------------- the kernel
template <bool kSmem>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocksPerSM) render_kernel(const RenderParams p) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ __align__(8) uint64_t stage_bar;

  const SceneDesc& sc = p.scene;
  const unsigned char* blob_base = sc.blob;
  if constexpr (kSmem) {
    // One TMA bulk copy of the scan blob per CTA; every warp then reads it with
    // broadcast LDS.128 for the rest of the kernel.
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&stage_bar, sc.blob_bytes);
      constexpr uint32_t kPiece = 32768;
      for (uint32_t off = 0; off < sc.blob_bytes; off += kPiece) {
        const uint32_t n = min(kPiece, sc.blob_bytes - off);
        bulk_g2s(smem_blob + off, sc.blob + off, n, &stage_bar);
      }
    }
    mbar_wait(&stage_bar, 0);
    blob_base = smem_blob;
  }
  const SceneView sv = scene_view(sc, blob_base);

  if (p.counters && threadIdx.x == 0 && blockIdx.x == 0) atomicMin(p.counters + 1, globaltimer_ns());
  unsigned int n_scans = 0;
  if (p.order_mode == 3) {  // USE_SINGLE_TASK: one warp of one CTA
    if (blockIdx.x != 0 || threadIdx.x >= 32u) return;
    single_task_loop<kSmem>(p, sc, sv, n_scans);
  } else {
    lane_loop<kSmem>(p, sc, sv, p.team_size, n_scans);
  }

  if (p.counters && (threadIdx.x & 31) == 0) atomicMax(p.counters + 3, globaltimer_ns());  // timeline: warp retired
  // work counters: one atomic per warp
  unsigned int warp_scans = n_scans;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_scans += __shfl_xor_sync(0xffffffffu, warp_scans, o);
  if ((threadIdx.x & 31) == 0 && p.counters) atomicAdd(p.counters, (unsigned long long)warp_scans);
}

// ---------------------------------------------------------------- ray-level test hook
__global__ void __launch_bounds__(128) probe_rays_kernel(const SceneDesc sc, int n, const float* __restrict__ rays, const uint32_t* __restrict__ seeds,
                                                         int mode, float* __restrict__ out_t, int32_t* __restrict__ out_index,
                                                         uint32_t* __restrict__ out_rng) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const SceneView sv = scene_view(sc, sc.blob);
  Ray ray;
  ray.o = v3(rays[7 * i], rays[7 * i + 1], rays[7 * i + 2]);
  ray.d = v3(rays[7 * i + 3], rays[7 * i + 4], rays[7 * i + 5]);
  ray.tm = rays[7 * i + 6];
  Rng rng { seeds[i] };
  const Best b = mode == 0 ? closest_hit<false, true>(sc, sv, ray, rng, true, 0, 1) : closest_hit_in_order<false>(sc, sv, ray, rng);
  const int key = b.id < 0 ? 0 : key_of(sc, b.id);
  out_t[i] = b.t, out_index[i] = b.id < 0 ? -1 : (key < 0 ? -1 - key : key), out_rng[i] = rng.s;
}
cudaError_t launch_probe_rays(const SceneDesc& scene, int n, const float* d_rays7, const uint32_t* d_seeds, int mode, float* d_t,
                              int32_t* d_index, uint32_t* d_rng, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  probe_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(scene, n, d_rays7, d_seeds, mode, d_t, d_index, d_rng);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) probe_math_kernel(int kind, int n, const float* __restrict__ in, float* __restrict__ out) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  float r;
  switch (kind) {
    case 0: r = t_sin(in[i]); break;
    case 1: r = t_cos(in[i]); break;
    case 2: r = t_log(in[i]); break;
    case 3: r = t_asin(in[i]); break;
    case 4: r = t_atan2(in[2 * i], in[2 * i + 1]); break;
    default: r = t_pow5(in[i]); break;
  }
  out[i] = r;
}
cudaError_t launch_probe_math(int kind, int n, const float* d_in, float* d_out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  probe_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(kind, n, d_in, d_out);
  return cudaGetLastError();
}

cudaError_t launch_lane(const RenderParams& p, int device, int grid_override, cudaStream_t stream, LaunchInfo* info) {
  int sms = 0;
  cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) return err;
  RenderParams q = p;
  const unsigned long long pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
  if (p.order_mode == 1)
    q.n_positions = (unsigned long long)p.tiles_x * (unsigned long long)p.tiles_y * (unsigned long long)(kTile * kTile);
  else if (p.order_mode == 3)
    q.n_positions = pixels;  // (USE_SINGLE_TASK: no queue)
  else
    q.n_positions = pixels, q.order_mode = 0, q.scramble = 1ull;  // row-major
  const bool smem = (int)p.scene.blob_bytes <= max_smem_blob_bytes(device);
  const size_t dyn = smem ? p.scene.blob_bytes : 0;
  auto kernel = smem ? render_kernel<true> : render_kernel<false>;
  if (smem) {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (err != cudaSuccess) return err;
  }
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlockThreads, dyn);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > kMaxBlocksPerSM) per_sm = kMaxBlocksPerSM;
  int grid = grid_override > 0 ? grid_override : sms * per_sm;
  if (p.order_mode == 3) grid = 1;
  // Lanes per pixel at launch: one in the normal case; with fewer than kMinPixelsPerTeam pixels per
  // team (a small region, or an image strongly scaled over many GPUs) the teams start larger.
  if (q.team_size <= 0) {
    const unsigned long long lanes = (unsigned long long)grid * kBlockThreads;
    int t = 1;
    while (t < kSphereChunk && pixels * (unsigned long long)t * 2ull < lanes * (unsigned long long)kMinPixelsPerTeam) t <<= 1;
    q.team_size = t;
  }
  if (info) info->grid = grid, info->block = kBlockThreads, info->smem_bytes = (int)dyn, info->blocks_per_sm = per_sm, info->staged = smem, info->team_size = q.team_size;
  kernel<<<grid, kBlockThreads, dyn, stream>>>(q);
  return cudaGetLastError();
}

}  // namespace ptb
