// pt_kernel.h -- launch interface of the render kernel (pt_kernel.cu).
#ifndef PT_KERNEL_H
#define PT_KERNEL_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_abi.h"
#include "pt_packed.h"

namespace ptb {

#ifndef PT_BLOCK_THREADS
#define PT_BLOCK_THREADS 128
#endif
#ifndef PT_MIN_BLOCKS_PER_SM
#define PT_MIN_BLOCKS_PER_SM 5
#endif
constexpr int kBlockThreads = PT_BLOCK_THREADS;
constexpr int kMinBlocksPerSM = PT_MIN_BLOCKS_PER_SM;
constexpr int kMaxBlocksPerSM = PT_MIN_BLOCKS_PER_SM;  // persistent grid = SMs x this

// Global hand-off queue for HEAVY pixels (wavefront kernel): a bounded multi-producer / multi-consumer ring,
// filled by CTAs whose pixels turn out to be deep, drained by the express CTAs and by every CTA whose own pixels
// have run out.  ctrl[0] = head and ctrl[1] = tail are free-running positions that live on across launches (a
// launch leaves the ring empty and every slot ready for its next lap); ctrl[2] = CTAs whose regular work is
// finished (they hand nothing off any more), reset per launch.
constexpr int kHeavyEntryWords = 20;
struct HeavyQueue {
  unsigned int* ctrl;
  unsigned int* ready;   // per slot: the position the slot is ready for (pos: write, pos + 1: read); starts at the slot's index
  float* entries;        // kHeavyEntryWords words per entry
  unsigned int cap;      // slots, a power of two
  unsigned int stamp;    // (launch number; informational)
};

// What the cost probe says about the frame, computed on the device by the tile-order kernel and read by the render
// kernel at its start (no host round trip): how many scans a sample takes on average, how much of the work sits in
// pixels far above that, and the scheduling knobs derived from the two.
struct FrameTuning {
  int heavy_rate;        // a pixel is HEAVY (handed to the express CTAs) above kHeavyBase + heavy_rate * samples scans
  int n_express;         // CTAs that only serve the hand-off queue
  int mean_scans_x1000;  // probed scans per sample, x 1000
  int heavy_share_x1000; // share of the probed work in pixels above 4 x the mean, x 1000
  int max_scans;         // deepest probed sample
  int pad[3];
};

struct RenderParams {
  SceneDesc scene;
  pt_camera cam;
  int width, height, spp, depth;
  pt_region region;
  float* out;               // device (or peer-mapped) pointer: the caller's packed float3 rows, or the float4 staging rows
  long long out_row_pitch;  // floats
  int out_pixel_floats;     // 3: `out` is the caller's framebuffer (scalar stores); 4: the staging buffer (one 16-byte store per pixel)
  // progressive rendering (pt_render_resume*): this launch traces samples [spp_from, spp) of every pixel; the per-pixel
  // state {sum r, g, b, RNG bits} is read when spp_from > 0 and written back when `state` is set
  float* state;             // 4 floats per pixel of the region, or null
  long long state_row_pitch;  // pixels
  int spp_from;
  unsigned long long* pixel_counter;  // work-queue heads ([0] everybody's, [1] the express CTAs'), zeroed before launch
  unsigned long long express_positions;  // leading queue positions reserved for the express CTAs (set by the launcher)
  unsigned long long* counters;       // [0] += closest-hit scans (may be null)
  int team_size;                      // lane kernel: lanes per pixel at launch (power of two, 1..32; 0 = automatic)
  int kernel_kind;                    // 0 = wavefront kernel (default), 1 = lane kernel
  int pool_cap;                       // wavefront kernel: pixels a CTA may hold (set by the launcher)
  unsigned int staged_bytes;          // wavefront kernel: bytes of the arena kept in shared memory (set by the launcher)
  unsigned int tree_list_bytes;       // wavefront kernel: dynamic shared memory behind them for the tree lists (set by the launcher)
  uint2* tree_spill;                  // wavefront kernel: global-memory continuation of the tree lists, kTreeSpillLists x tree_spill_cap items per CTA (null: none)
  unsigned int tree_spill_cap;
  int n_express;                      // wavefront kernel: CTAs that only serve the hand-off queue (< 0 = automatic)
  const FrameTuning* tuning;          // wavefront kernel: the cost probe's verdict (null: no probe ran; the defaults apply)
  // pixel order of the wavefront kernel (queue position -> pixel)
  int order_mode;                     // 0 = scrambled, 1 = tiles in `tile_order` (heaviest first), 2 = cost probe grid
  const int* tile_order;              // order_mode 1: tile ids (kTile x kTile pixels) in processing order
  int tiles_x, tiles_y;
  int* probe_cost;                    // order_mode 2: scans of one throw-away sample per probed pixel
  unsigned long long n_positions;     // queue length (set by the launcher)
  unsigned long long scramble;        // pixel-order multiplier (coprime with the pixel count), set by the launcher
  HeavyQueue heavy;
};

struct LaunchInfo {
  int grid, block, smem_bytes, blocks_per_sm, team_size;
  bool staged;  // scan blob staged in shared memory (else streamed from L2)
};

constexpr int kTile = 8;   // LPT ordering granularity (pixels)
constexpr int kTreeSpillLists = 3;       // node items of tree pass 0, of pass 1, leaf items
constexpr unsigned int kTreeSpillCap = 32768;  // items per list and CTA beyond shared memory before a thread walks its subtree in place
#ifndef PT_PROBE_STEP
#define PT_PROBE_STEP 2
#endif
constexpr int kProbeStep = PT_PROBE_STEP;  // the cost probe traces every kProbeStep-th pixel of every kProbeStep-th row

// Sort the region's tiles by probed cost, heaviest first (one small kernel).  `scratch` holds n_tiles ints.
cudaError_t launch_tile_order(const int* probe_cost, int region_w, int region_h, int tiles_x, int tiles_y,
                              int* tile_order, int* scratch, FrameTuning* tuning, int grid, int n_express_forced, int probe_spp, cudaStream_t stream);
int wave_grid(int device);  // CTAs of a wavefront launch on this device

// The staged framebuffer {r, g, b, -} x (w x h) -> the caller's packed float3 rows, whole 16-byte stores, coalesced
// (w must be a multiple of 4 and `out` rows 16-byte aligned).
cudaError_t launch_resolve_fb(const float* stage, int w, int h, float* out, long long out_row_pitch, cudaStream_t stream);

// Test hook: the closest-hit scan (render.hpp:30-51) of n given rays, one thread per ray, each with its own generator
// state: t, the hit object's index in the scene's vector (-1: none), the generator afterwards.  mode 0 = the product's
// scan (chunk boxes, flat trees, grazing index; vector order for rays that can meet a NaN), 1 = vector order for all.
cudaError_t launch_probe_rays(const SceneDesc& scene, int n, const float* d_rays7, const uint32_t* d_seeds, int mode, float* d_t,
                              int32_t* d_index, uint32_t* d_rng, cudaStream_t stream);

// Test hook: out[i] = f(in[i]) with the device's transcendentals (pt_device.cuh): kind 0 sin, 1 cos, 2 log, 3 asin,
// 4 atan2(in[2 i], in[2 i + 1]), 5 pow(x, 5).
cudaError_t launch_probe_math(int kind, int n, const float* d_in, float* d_out, cudaStream_t stream);

int max_smem_blob_bytes(int device);
cudaError_t launch_render(const RenderParams& p, int device, int grid_override, cudaStream_t stream,
                          LaunchInfo* info);
// the lane kernel's launch (pt_lane.cu); launch_render() dispatches to it for kernel_kind == 1
cudaError_t launch_lane(const RenderParams& p, int device, int grid_override, cudaStream_t stream, LaunchInfo* info);

}  // namespace ptb
#endif
