// pt_wave.cu -- the default render kernel: a bulk-synchronous wavefront inside one CTA per SM
// (reference include/render.hpp:25-106 and everything it calls, re-scheduled for sm_100a).
//
// Execution model (B200-first, not a translation of the SYCL kernel):
//   * A pixel's `spp` samples are inherently serial -- its xorshift32 stream is consumed in a
//     data-dependent way (render.hpp:95-101) -- so the unit of work is a PIXEL, pulled from a global
//     atomic queue; a finished pixel is replaced at once ("path regeneration").  Parallelism is
//     pixels x objects.
//   * The scene's scan blob (pt_packed.h) and its side tables are staged into shared memory with
//     cp.async.bulk (TMA) once per CTA.  Objects come in k-d ordered CHUNKS behind conservative bounding
//     boxes; a ray only looks at the chunks whose box it crosses (pt_prims.cuh: a proof, the result is
//     bit-identical with and without it).
//   * The winner of a scan is (minimum t, then maximum key), which reproduces the sequential scan's tie
//     behaviour for ANY visiting order (pt_packed.h): that is what allows skipping chunks, splitting a
//     scan over lanes or work items, and merging with shuffles or one 64-bit atomicMin.
//   * One 896-thread CTA per SM, the path state of 896 pixels in a structure-of-arrays pool in shared
//     memory, phases BOXES / ITEMS / LATE / SHADE separated by __syncthreads(); the closest-hit scan runs
//     as (ray, chunk) work items spread over the whole CTA; shading runs in warps of one material kind.
//     Deep pixels (paths bouncing dozens of times inside glass hold ten times the average work and would
//     sit on a serial critical path) are traced by EXPRESS CTAs in short rounds, from the head of a
//     longest-processing-time-first pixel order and from a global hand-off queue.
// All arithmetic follows the operation order of the reference; see pt_device.cuh for the numerics contract.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "pt_abi.h"
#include "pt_device.cuh"
#include "pt_kernel.h"
#include "pt_packed.h"
#include "pt_prims.cuh"
#include "pt_shade.cuh"

namespace ptb {

namespace {
// ---------------------------------------------------------------- the wavefront kernel
// Bulk-synchronous wavefront inside one CTA per SM.  The path state of up to kWavePool pixels lives
// in a structure-of-arrays RAY POOL in shared memory instead of in the registers of fixed lanes, and
// the CTA alternates between phases separated by __syncthreads():
//
//   BOXES   one thread per ray: which sphere chunks does the ray cross (chunk culling, above)?  Every
//           (ray, chunk) pair becomes a work ITEM in a CTA-wide list.
//   SPHERES one thread per ITEM: the 16 spheres of the chunk against the ray; a hit is merged into the
//           ray's winner with one 64-bit shared-memory atomicMin on {t, original object index} -- exactly
//           the reference's rule for spheres (smallest t, then the earlier object, sphere.hpp:77,93).
//           The work per item is the same whatever the ray, so the lanes of a warp stay busy although
//           their rays cross different numbers of chunks, and a CTA with few rays left still has
//           (rays x chunks) items to spread over its threads.
//   LATE    one thread per ray: the groups from the first constant_medium on (flat objects and media in object
//           order against the running closest hit), then what has to happen next -- background, or the kind of
//           the hit material -- and the ray joins that kind's list ("compact divergent material work").
//   SHADE   one warp per unit of up to 32 rays of ONE kind, so the lanes of a warp run the same material
//           code; finished paths start their pixel's next sample (the RNG stream of a pixel is strictly
//           serial) or write the pixel and pull a new one from the global pixel queue.
// A scene with spheres BEHIND a constant_medium in the object list is scanned sequentially per ray in
// BOXES instead (the medium needs the running closest hit of everything before it, and what follows
// needs the medium's).
//
// HAND-OFF QUEUE.  A pixel's samples are serial and a full round takes tens of microseconds, so the
// few pixels that hold ten times the average work (paths bouncing dozens of times inside glass:
// 3 000 scans where the mean is 260) would sit on a critical path longer than the whole frame, and
// at the end of the frame every CTA would drain its own leftovers alone.  A pixel whose scan rate
// marks it as HEAVY is therefore handed, with its complete path state, to a global queue.  It is
// taken over by a CTA that runs SHORT rounds because it keeps only kExpressPool rays in flight: one
// of a few EXPRESS CTAs that do nothing else, or any CTA whose own pixels have run out -- which also
// balances the end of the frame across the whole GPU.  With few rays the phases switch to finer
// work units (a ray's boxes in blocks of 8, a chunk's spheres in quarters).
//
// Results are bit-identical to the lane kernel: the same device functions are called on the same
// per-pixel state, only the assignment of work to lanes differs.
#ifndef PT_WAVE_THREADS
#define PT_WAVE_THREADS 896
#endif
#ifndef PT_WAVE_BLOCKS_PER_SM
#define PT_WAVE_BLOCKS_PER_SM 1
#endif
#ifndef PT_WAVE_ROUNDS
#define PT_WAVE_ROUNDS 1
#endif
#ifndef PT_HEAVY_RATE
#define PT_HEAVY_RATE 10
#endif
#ifndef PT_EXPRESS_POOL
#define PT_EXPRESS_POOL 64
#endif
#ifndef PT_WAVE_ITEMS
#define PT_WAVE_ITEMS 4096
#endif
constexpr int kWaveThreads = PT_WAVE_THREADS;
constexpr int kWavePool = PT_WAVE_ROUNDS * kWaveThreads;  // pixels (rays) a CTA keeps in flight: whole scan passes
constexpr int kWaveKinds = 6;                              // 0 = background, 1 + PT_MAT_* otherwise
constexpr int kHeavyRate = PT_HEAVY_RATE;                  // heavy: more than kHeavyBase + rate * samples scans so far
constexpr int kHeavyBase = 64;
constexpr int kExpressPool = PT_EXPRESS_POOL;              // rays in flight in a CTA that serves the hand-off queue
constexpr int kWaveItems = PT_WAVE_ITEMS;                  // (ray, chunk) items per round; the overflow is scanned in place
constexpr int kWaveItemsStatic = kWaveItems * 3 / 8;       // ... of static spheres (from the front of the list)
constexpr int kWaveItemsMoving = kWaveItems - kWaveItemsStatic;  // ... of moving spheres (from the back)
constexpr unsigned long long kNoHit64 = 0x7f800000ffffffffull;  // {t = +inf, no object}
#ifndef PT_FINE_RAYS
#define PT_FINE_RAYS 160
#endif
constexpr int kFineRays = PT_FINE_RAYS;  // at most this many rays in the round: finer work units (short rounds)
#ifndef PT_FINE_BOXES
#define PT_FINE_BOXES 8
#endif
constexpr int kFineBoxes = PT_FINE_BOXES;            // ... BOXES: a ray's chunks in blocks of this many
constexpr int kFineQuarter = 4;          // ... SPHERES: a chunk's spheres in runs of this many
#ifndef PT_QUEUE_PER_EXPRESS
#define PT_QUEUE_PER_EXPRESS 16
#endif
// The tree functions below exist in the kTrees = true instantiations only (scenes without flat trees never pay for them).
// Inlined there: out of line they cost the mesh frame 9 % (382 -> 351 ms at 64 spp) -- a call saves and restores the
// caller's registers through local memory, and the mesh kernel's L1 is busy with nodes and triangles.
#ifndef PT_TREE_FN
#define PT_TREE_FN __forceinline__
#endif
constexpr int kQueuePerExpress = PT_QUEUE_PER_EXPRESS;  // waiting pixels per express CTA beyond which nobody hands off
constexpr int kHandoffPause = 8;
constexpr int kMaxBoxBlocks = 64;
constexpr int kMaxFlats = 256;
constexpr int kMaxTreeGroups = 8;
#ifndef PT_TREE_UNROLL
#define PT_TREE_UNROLL 4
#endif
#ifndef PT_LEAF_SUB
#define PT_LEAF_SUB 1
#endif
constexpr int kTreeUnroll = PT_TREE_UNROLL;  // node boxes tested per trip of a tree expansion's loop
static_assert(kWavePool <= 1024, "an item packs the pool slot into 10 bits");

// What the out-of-line tree code needs of the kernel's context.  It lives in shared memory because the arguments of a
// __noinline__ function beyond a few words travel through LOCAL memory: with the key table, the tree pointers, the group
// and the ray passed by value the mesh frame moved 4.4 G local-memory sectors, and half of its stall samples waited on them.
struct WaveTreeCtx {
  const SceneDesc* sc;        // the CTA's scene descriptor (shared memory)
  const unsigned char* blob;  // base of the scan blob as this CTA sees it (staged or global)
  uint2* lists;               // the tree lists
  uint2* spill;               // ... and this CTA's continuation of each in global memory (null: none)
  int spill_cap;
  unsigned long long* counters;
};

struct WavePool {
  unsigned long long best64[kWavePool];  // SPHERES: {float bits of t, original object index} of the ray's winner
  uint2 items[kWaveItems];               // {slot | chunk << 10, f bits}: static-sphere items from the front, moving from the back
  float ox[kWavePool], oy[kWavePool], oz[kWavePool], dx[kWavePool], dy[kWavePool], dz[kWavePool], tm[kWavePool];
  float hit_t[kWavePool];
  int hit_id[kWavePool];
  float att_x[kWavePool], att_y[kWavePool], att_z[kWavePool];
  float acc_x[kWavePool], acc_y[kWavePool], acc_z[kWavePool];
  uint32_t rng[kWavePool];
  uint32_t pix[kWavePool];  // position of the pixel in the work queue
  int sample[kWavePool];
  int bounce[kWavePool];
  int scans[kWavePool];               // closest-hit scans spent on the current pixel (< 0: taken over, never handed off again)
  unsigned short list_a[kWavePool];   // rays to scan (unordered)
  unsigned short list_k[kWaveKinds][kWavePool];  // the same rays by kind (what happens next), filled by LATE
  unsigned short free_list[kWavePool];  // hand-off service: pool slots without a pixel
  int counts[8];
  int n_next;      // length of list_a being built
  int n_own;       // of those, pixels this CTA pulled from the pixel queue itself
  int n_items_s, n_items_m;  // static / moving items reserved this round (may exceed what fits)
  int cap_items_s;           // the static spheres' share of `items` this round (the moving spheres have the rest)
  int n_tgroups;   // flat groups with a tree in front of the first constant_medium (expanded breadth first: wave_tree_expand)
  int tgroups[kMaxTreeGroups];
  unsigned long long express_positions;  // leading positions of the LPT order that belong to the express CTAs
  int heavy_rate;  // hand-off threshold of this frame (FrameTuning)
  int queue_limit;    // hand-offs stop while the ring holds this many pixels (back-pressure: reserve_heavy)
  int handoff_pause;  // ... and are not tried again for this many rounds
  int n_express;   // express CTAs of this frame
  int tree_passes; // tree passes per round: the deepest of those trees has this many levels above its leaves
  WaveTreeCtx tctx;
  int tl_n[3], tl_off[3], tl_cap[3];  // the tree lists (in dynamic shared memory behind the staged scene): node items of pass 0 and 1, leaf items
  int free_count;
  int pixel_dry;   // the pixel queue has run dry
  int in_order;    // a ray stored since the last BOXES phase needs the scan in vector order (store_ray)
  int service;     // hand-off service: 0 = keep polling, 1 = every producer is done and the queue is empty
  int blocks_ok;   // the sphere chunks fit the block table (else no short rounds)
  int n_blocks;    // fine BOXES: blocks of <= kFineBoxes chunks over all sphere groups (0: too many, coarse only)
  int4 blocks[kMaxBoxBlocks];  // {group, first chunk, chunks, 0}
  int n_flats;     // rectangles, triangles and boxes in FRONT of the first constant_medium: tested one thread per (ray, object)
  int first_late_group;  // the first constant_medium's group (n_groups if none): from here on the scan is sequential per ray
  int2 flats[kMaxFlats];  // {group type, element}
};

PT_DEV int material_of(const SceneDesc& sc, int id) {
  const int idx = id & (int)kIdMask;
  switch (id >> kIdShift) {
    case G_SPHERE: return sc.sphere_aux[idx].material;
    case G_MOVING_SPHERE: return sc.moving_aux[idx].material;
    case G_RECT: return sc.rect_aux[idx].material;
    case G_TRIANGLE: return sc.tri_aux[idx].material;
    case G_BOX: return sc.box_aux[idx].material;
    default: return sc.media[idx].material;
  }
}

// A ray's winner as one 64-bit word whose UNSIGNED ORDER is the winner rule of pt_packed.h (smaller t, then larger
// key): {float bits of t (t >= 0), 0x7fffffff - key}.  Spheres (key = -1 - object) give codes 0x80000000 + object,
// the others (key = object) 0x7fffffff - object.  A NaN t orders behind +inf and is never taken.
PT_DEV unsigned long long pack_winner(float t, int key) {
  return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(0x7fffffffu - (uint32_t)key);
}
PT_DEV unsigned long long pack_sphere_winner(const SphereAux* aux, const Best& b) {
  return pack_winner(b.t, aux[b.id & (int)kIdMask].key);
}
PT_DEV Best unpack_winner(const SceneDesc& sc, unsigned long long v) {
  Best best;
  const uint32_t code = (uint32_t)v;
  best.t = __uint_as_float((uint32_t)(v >> 32));
  best.id = code == 0xffffffffu ? -1 : sc.object_id[code >= 0x80000000u ? code - 0x80000000u : 0x7fffffffu - code];
  return best;
}

}  // namespace

namespace {
// One (ray, leaf) item of a flat group with a tree -- {slot | tree group << 10, leaf | grazing-index flag << 31} --
// for the elements first, first + step, ... of the leaf.
template <bool kSmem, typename Keys>
PT_DEV void flat_item_scan(const Keys& sc, const FlatTrees& ft, const Group& g, uint32_t leaf_word, const Ray& ray, int first, int step,
                           Best& b) {
  const int leaf = (int)(leaf_word & 0x7fffffffu);
  if (leaf_word >> 31) {
    const Tree gt = ft.trees[g.gtree];
    scan_graze_leaf<kSmem>(sc, ft.data, ft.ids + gt.leaf_ids + leaf * kFlatChunk, first, step, ray, graze_threshold(ray), b);
  } else {
    const int base = g.begin + leaf * kFlatChunk;
    scan_flat_range<kSmem>(sc, ft.data, g.type, base + first, min(base + kFlatChunk, g.begin + g.count), step, ray, b);
  }
}
PT_DEV Ray pool_ray(const WavePool* W, int slot) {
  Ray ray;
  ray.o = v3(W->ox[slot], W->oy[slot], W->oz[slot]);
  ray.d = v3(W->dx[slot], W->dy[slot], W->dz[slot]);
  ray.tm = W->tm[slot];
  return ray;
}
// ITEMS: one (ray, leaf) item, merged into the ray's winner.
template <bool kSmem>
__device__ PT_TREE_FN void wave_run_flat(WavePool* W, uint2 it, int first, int step) {
  const SceneDesc& sc = *W->tctx.sc;
  const SceneView sv = scene_view(sc, W->tctx.blob);
  const Group g = sv.groups()[W->tgroups[(it.x >> 10) & 7u]];
  const int slot = (int)(it.x & 1023u);
  const Ray ray = pool_ray(W, slot);
  Best b { kInf, -1 };
  flat_item_scan<kSmem>(sc, flat_trees(sc, sv, g.type), g, it.y, ray, first, step, b);
  if (b.id >= 0) atomicMin(&W->best64[slot], pack_winner(b.t, key_of(sc, b.id)));
}
// TREE EXPANSION (BOXES and the tree passes): test the nodes [first, first + count) of level `level` of one tree of
// one flat group against one ray.  A crossed leaf becomes a (ray, leaf) item for ITEMS, a crossed inner node a (ray,
// node) item for the next tree pass -- the trees are walked BREADTH FIRST, one level per pass, every pass a flat list of
// items of the same size spread over the whole CTA, like the (ray, chunk) items of the spheres: a depth-first walk
// per ray leaves the lanes of a warp in loops of different lengths (measured on the 10 002-triangle mesh: 9.8 of 32
// lanes busy, and the CTA waiting at the barrier for its deepest walk).  What does not fit a list is walked depth
// first here and now, and folded into the ray's winner `v`.
//   node item  {slot | tree group << 10 | grazing index << 13 | level << 14, node}
//   leaf item  {slot | tree group << 10, leaf | grazing index << 31}
template <bool kSmem>
PT_DEV unsigned long long tree_expand(WavePool* W, const SceneDesc& sc, const FlatTrees& ft, const Group& g, const Tree& t, uint32_t slot_tg,
                                      uint32_t graze, int level, int first, int count, int out_list, const Ray& ray,
                                      unsigned long long v) {
  const FlatRay fr = make_flat_ray(ft.extent, ray);
  if (graze && !fr.ok) return v;  // (a ray that is not culled visits every leaf of the box tree already)
  const float4* boxes = ft.nodes + tree_off(t, level);
  uint2* const lists = W->tctx.lists;
  auto test = [&](const float4* box) {
    return graze ? graze_node_bits<kSmem>(box, ray.d, fr.taud) : flat_node_bits<kSmem>(box, fr, ray, kInf);
  };
  auto leaf_here = [&](int leaf) {  // overflow: the leaf's elements, in place
    if (W->tctx.counters) atomicAdd(W->tctx.counters + 15, 1ull);  // stats: items scanned in place
    Best b { kInf, -1 };
    flat_item_scan<kSmem>(sc, ft, g, (uint32_t)leaf | (graze << 31), ray, 0, 1, b);
    if (b.id >= 0) {
      const unsigned long long w64 = pack_winner(b.t, key_of(sc, b.id));
      if (w64 < v) v = w64;
    }
  };
  const int dest = level == 0 ? 2 : out_list;
#pragma unroll 1
  for (int c = first; c < first + count; c += 32) {
    const int k = min(32, first + count - c);
    uint32_t m = 0;
#pragma unroll kTreeUnroll  // (the boxes come from L2 when the scene is too large to stage: several loads in flight)
    for (int j = 0; j < k; ++j) m = __funnelshift_l(test(boxes + 2 * (c + j)), m, 1);
    if (m == 0u) continue;
    int at = atomicAdd(&W->tl_n[dest], __popc(m));
#pragma unroll 1
    while (m) {
      const int top = 31 - __clz((int)m);
      m &= ~(1u << top);
      const int node = c + (k - 1 - top);
      if (at < W->tl_cap[dest] + W->tctx.spill_cap) {
        const uint2 item = level == 0 ? make_uint2(slot_tg, (uint32_t)node | (graze << 31))
                                      : make_uint2(slot_tg | (graze << 13) | ((uint32_t)level << 14), (uint32_t)node);
        if (at < W->tl_cap[dest])
          lists[W->tl_off[dest] + at] = item;
        else  // (a ray along the mesh crosses hundreds of leaves: the list continues in global memory)
          W->tctx.spill[dest * W->tctx.spill_cap + (at - W->tl_cap[dest])] = item;
      } else if (level == 0) {
        leaf_here(node);
      } else {
        const int c0 = node * kTreeFan;
        tree_walk(t, ft.nodes, level - 1, c0, min(kTreeFan, tree_n(t, level - 1) - c0), test, leaf_here);
      }
      ++at;
    }
  }
  return v;
}
// (Scalar arguments only, the context in shared memory -- WaveTreeCtx -- so that -DPT_TREE_FN=__noinline__ stays cheap.)
// BOXES: the top level of the trees [graze_from, graze_to) (0: boxes, 1: grazing index) of tree group `tg` for the ray in `slot`.
template <bool kSmem>
__device__ PT_TREE_FN unsigned long long wave_tree_roots(WavePool* W, int slot, int tg, uint32_t graze_from, uint32_t graze_to, unsigned long long v) {
  const SceneDesc& sc = *W->tctx.sc;
  const SceneView sv = scene_view(sc, W->tctx.blob);
  const Group g = sv.groups()[W->tgroups[tg]];
  const FlatTrees ft = flat_trees(sc, sv, g.type);
  const Ray ray = pool_ray(W, slot);
#pragma unroll 1
  for (uint32_t graze = graze_from; graze < graze_to; ++graze) {
    const int ti = graze ? g.gtree : g.tree;
    if (ti < 0) break;
    const Tree t = ft.trees[ti];
    v = tree_expand<kSmem>(W, sc, ft, g, t, (uint32_t)slot | ((uint32_t)tg << 10), graze, t.levels - 1, 0, tree_n(t, t.levels - 1), 0, ray, v);
  }
  return v;
}
// A tree pass: the children sub, sub + 1, ... (`width` of them) of one (ray, node) item.
template <bool kSmem>
__device__ PT_TREE_FN void wave_tree_node_item(WavePool* W, uint2 it, int pass, int sub, int width) {
  const SceneDesc& sc = *W->tctx.sc;
  const SceneView sv = scene_view(sc, W->tctx.blob);
  const int slot = (int)(it.x & 1023u);
  const Group g = sv.groups()[W->tgroups[(it.x >> 10) & 7u]];
  const uint32_t graze = (it.x >> 13) & 1u;
  const int level = (int)(it.x >> 14);
  const FlatTrees ft = flat_trees(sc, sv, g.type);
  const Tree t = ft.trees[graze ? g.gtree : g.tree];
  const int c0 = (int)it.y * kTreeFan + sub;
  const int count = min(width, tree_n(t, level - 1) - c0);
  if (count <= 0) return;
  const unsigned long long v = tree_expand<kSmem>(W, sc, ft, g, t, it.x & 0x1fffu, graze, level - 1, c0, count, pass + 1, pool_ray(W, slot), kNoHit64);
  if (v != kNoHit64) atomicMin(&W->best64[slot], v);
}
}  // namespace

// The whole scan of one ray by one thread, in group order (closest_hit, pt_prims.cuh): scenes with spheres behind a
// constant_medium, and rounds in which a ray can meet a NaN.  Out of line (see wave_build_tables).
template <bool kSmem, bool kTrees>
__device__ __noinline__ void wave_sequential_scan(WavePool* W, const SceneDesc* scp, const unsigned char* blob_base, int slot) {
  const SceneDesc& sc = *scp;
  const SceneView sv = scene_view(sc, blob_base);
  Ray ray;
  ray.o = v3(W->ox[slot], W->oy[slot], W->oz[slot]);
  ray.d = v3(W->dx[slot], W->dy[slot], W->dz[slot]);
  ray.tm = W->tm[slot];
  Rng rng { W->rng[slot] };
  const Best best = closest_hit<kSmem, kTrees>(sc, sv, ray, rng, true, 0, 1);
  W->hit_t[slot] = best.t, W->hit_id[slot] = best.id;
  W->rng[slot] = rng.s;  // a constant_medium may have drawn from it (constant_medium.hpp:65)
}

// The CTA's tables, built once by one thread: the sphere chunks in blocks for the short rounds, the (ray, flat object)
// units, the flat groups with a tree and the shares of the tree lists.  Out of line, like every RARE path of this
// kernel: code that runs once per launch -- or never, for the scene at hand -- must not cost the hot loops registers
// (measured: a dormant branch inlined into the kernel moved the default scene's frame by 3 %).
template <bool kSmem, bool kTrees>
__device__ __noinline__ void wave_build_tables(WavePool* W, const SceneDesc* scp, const Group* groups, const Tree* trees, unsigned int tree_list_bytes) {
  const SceneDesc& sc = *scp;
  const int n_groups = (int)sc.n_groups;
  int nb = 0;
  for (int gi = 0; gi < n_groups && nb >= 0; ++gi) {
    const Group g = groups[gi];
    if (g.type != G_SPHERE && g.type != G_MOVING_SPHERE) continue;
    const int c_end = (g.begin + g.count) / kSphereChunk;
    for (int cb = g.begin / kSphereChunk; cb < c_end; cb += kFineBoxes) {
      if (nb == kMaxBoxBlocks) {
        nb = -1;
        break;
      }
      W->blocks[nb++] = make_int4(gi, cb, min(kFineBoxes, c_end - cb), 0);
    }
  }
  W->n_blocks = nb < 0 ? 0 : nb, W->blocks_ok = nb >= 0;
  // the flat objects in front of the first medium (as many as fit; the rest stays with the sequential part)
  int nf = 0, nt = 0, late = n_groups;
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = groups[gi];
    const bool flat = g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX;
    const bool tree = kTrees && flat && has_tree(sc, g);
    if (g.type == G_MEDIUM || (flat && !tree && nf + 6 * g.count > kMaxFlats) || (tree && nt == kMaxTreeGroups)) {
      late = gi;
      break;
    }
    if (tree) {  // a group with a tree: traversed per ray in BOXES, its leaves become items
      W->tgroups[nt++] = gi;
      continue;
    }
    if (g.type == G_RECT || g.type == G_TRIANGLE)
      for (int i = 0; i < g.count; ++i) W->flats[nf++] = make_int2(g.type, g.begin + i);
    if (g.type == G_BOX)  // a box is six independent sides (box.hpp:20-25): the closest side is the box's hit
      for (int i = 0; i < g.count; ++i)
        for (int side = 0; side < 6; ++side) W->flats[nf++] = make_int2(G_BOX | (side << 8), g.begin + i);
  }
  W->n_flats = nf, W->first_late_group = late, W->n_tgroups = nt;
  // the tree lists share what the launch left of the dynamic shared memory: node items of pass 0 / pass 1 / leaf items
  int passes = 0;
  for (int k = 0; k < nt; ++k) {
    const Group g = groups[W->tgroups[k]];
    passes = max(passes, trees[g.tree].levels - 1);
    if (g.gtree >= 0) passes = max(passes, trees[g.gtree].levels - 1);
  }
  W->tree_passes = passes;
  const int cap = (int)(tree_list_bytes / sizeof(uint2));
  const int c0 = passes >= 1 ? cap / (passes == 1 ? 3 : 4) : 0, c1 = passes >= 2 ? cap / 4 : 0;
  W->tl_off[0] = 0, W->tl_cap[0] = c0, W->tl_off[1] = c0, W->tl_cap[1] = c1, W->tl_off[2] = c0 + c1, W->tl_cap[2] = cap - c0 - c1;
  W->tl_n[0] = W->tl_n[1] = W->tl_n[2] = 0;
}

// kTrees = false: the scene has no flat group with a tree (or culling is off) -- the tree paths are compiled out, so
// that scenes without them (spheres and a handful of flats: the default scene) keep their registers and their
// instruction-cache footprint.
template <bool kSmem, bool kTrees>
__global__ void __launch_bounds__(kWaveThreads, PT_WAVE_BLOCKS_PER_SM) render_wave_kernel(const RenderParams p) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ __align__(8) uint64_t stage_bar;
  __shared__ SceneDesc staged_scene;  // the scene descriptor with the staged tables' pointers redirected to shared memory

  // dynamic shared memory: the ray pool first (at a compile-time offset, so that its accesses need no address arithmetic),
  // then the staged part of the arena
  constexpr uint32_t kPoolBytes = ((uint32_t)sizeof(WavePool) + 127u) & ~127u;
  unsigned char* const smem_stage = smem_blob + kPoolBytes;
  const unsigned char* blob_base = p.scene.blob;
  const uint32_t staged = kSmem ? p.staged_bytes : 0u;  // the scan blob, and the side tables behind it when they fit too
  if constexpr (kSmem) {
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&stage_bar, staged);
      constexpr uint32_t kPiece = 32768;
      for (uint32_t off = 0; off < staged; off += kPiece)
        bulk_g2s(smem_stage + off, p.scene.blob + off, min(kPiece, staged - off), &stage_bar);
    }
    blob_base = smem_stage;
  }
  if (threadIdx.x == 0) {
    staged_scene = p.scene;
    const unsigned char* g0 = p.scene.blob;
    auto redirect = [&](auto& ptr) {
      const size_t off = (size_t)(reinterpret_cast<const unsigned char*>(ptr) - g0);
      if (kSmem && off < (size_t)staged) ptr = reinterpret_cast<decltype(ptr + 0)>(smem_stage + off);
    };
    redirect(staged_scene.sphere_aux), redirect(staged_scene.moving_aux), redirect(staged_scene.rect_aux);
    redirect(staged_scene.tri_aux), redirect(staged_scene.box_aux), redirect(staged_scene.media);
    redirect(staged_scene.keys), redirect(staged_scene.object_id);
    const unsigned char* mats = static_cast<const unsigned char*>(staged_scene.materials);
    redirect(mats);
    staged_scene.materials = mats;
  }
  __syncthreads();
  if constexpr (kSmem) mbar_wait(&stage_bar, 0);
  const SceneDesc& sc = staged_scene;
  WavePool& W = *reinterpret_cast<WavePool*>(smem_blob);
  const SceneView sv = scene_view(sc, blob_base);

  if (p.counters && threadIdx.x == 0 && blockIdx.x == 0) atomicMin(p.counters + 1, globaltimer_ns());
  // (only what is cheap to keep: everything else that is loop invariant is re-formed where it is used -- the kernel runs at
  // its register limit, and a value that is live across all phases costs more than the instruction that makes it)
  const int tid = (int)threadIdx.x, lane = tid & 31;
  const pt_camera& cam = p.cam;
  const HeavyQueue& hq = p.heavy;
  // the frame's scheduling knobs: the cost probe's verdict when one ran (read from device memory: no host round trip),
  // else the launcher's defaults.  (The probe itself runs with n_express = 0 and no tuning.)
  // (kept in shared memory, not in registers that would be live across the whole kernel)
  if (threadIdx.x == 0) {
    const int n_express = p.tuning ? min(__ldg(&p.tuning->n_express), (int)gridDim.x - 1) : p.n_express;
    W.n_express = n_express;
    W.heavy_rate = p.tuning ? __ldg(&p.tuning->heavy_rate) : kHeavyRate;
    W.queue_limit = max(n_express, 1) * kQueuePerExpress, W.handoff_pause = 0;
    W.express_positions = p.order_mode == 1 ? min((unsigned long long)n_express * (unsigned long long)kExpressPool, p.n_positions) : 0ull;
  }
  __syncthreads();
  const bool express = (int)blockIdx.x < W.n_express;  // this CTA only serves the hand-off queue
  const bool sequential_scan = sc.n_late_sphere_groups != 0u;
  unsigned int n_scans = 0;

  // Pull the next pixel of the queue; false (and the CTA-wide flag set) when the queue is dry.
  // The first p.express_positions positions of the queue (the most expensive tiles of the LPT order) belong to the
  // express CTAs, which trace them in short rounds from the start; everybody else begins behind them.
  auto next_pixel = [&](uint32_t& pixq, Rng& rng, int& px, int& py, V3& acc0, int& sample0) -> bool {
    for (;;) {
      if (*reinterpret_cast<volatile int*>(&W.pixel_dry)) return false;  // (a set-once flag; a stale 0 only costs one more atomic)
      unsigned long long pos;
      if (express) {
        pos = atomicAdd(p.pixel_counter + 1, 1ull);
        if (pos >= W.express_positions) {
          atomicExch(&W.pixel_dry, 1);
          return false;
        }
      } else {
        pos = W.express_positions + atomicAdd(p.pixel_counter, 1ull);
        if (pos >= p.n_positions) {
          atomicExch(&W.pixel_dry, 1);
          if (p.counters) atomicMin(p.counters + 2, globaltimer_ns());  // timeline: queue ran dry
          return false;
        }
      }
      float *unused, *state_px;
      if (!queue_pixel(p, pos, px, py, unused, state_px)) continue;  // a tile position outside the region
      pixq = (uint32_t)pos;
      pixel_start(p, px, py, state_px, rng, acc0, sample0);
      return true;
    }
  };
  // The hand-off queue is a bounded multi-producer / multi-consumer RING (pt_kernel.h): position `pos` lives in slot
  // pos mod cap, whose sequence word says what the slot is ready for -- pos: to be written, pos + 1: to be read, pos +
  // cap: the next lap.  Nobody ever waits: a full ring keeps the pixel where it is, an empty one (or an entry still
  // being written) is polled again next time.
  // Claim the next entry: its position (read it, then release_heavy()), or false when there is none right now.
  auto take_heavy = [&](unsigned int& pos_out) -> bool {
    unsigned int pos = ld_volatile_u32(hq.ctrl + 0);
    if (pos == ld_volatile_u32(hq.ctrl + 1)) return false;  // empty (the common answer: two independent loads, one round trip)
    for (int attempt = 0; attempt < 8; ++attempt) {
      const int dif = (int)(ld_volatile_u32(hq.ready + (pos & (hq.cap - 1u))) - (pos + 1u));
      if (dif == 0) {
        const unsigned int seen = atomicCAS(hq.ctrl + 0, pos, pos + 1u);
        if (seen == pos) {
          __threadfence();
          pos_out = pos;
          return true;
        }
        pos = seen;
      } else if (dif < 0) {
        return false;
      } else {
        pos = ld_volatile_u32(hq.ctrl + 0);
      }
    }
    return false;
  };
  auto release_heavy = [&](unsigned int pos) {
    __threadfence();
    *reinterpret_cast<volatile unsigned int*>(hq.ready + (pos & (hq.cap - 1u))) = pos + hq.cap;
  };
  // Reserve a slot for a pixel that leaves: its position (write the entry, then publish_heavy()), or false: ring full.
  // BACK-PRESSURE: a pixel is only handed off while the ring is short.  A pixel parked in a long queue makes no progress
  // at all, and what waits when the pixel queue runs dry is the frame's tail (measured on the mesh, 256 spp: 4 916 pixels
  // handed to 9 express CTAs waited 1.1 s on average and the frame ended 0.66 s after the queue was dry); a pixel that
  // stays advances one bounce per round of its own CTA, and is offered again a few rounds later.
  auto reserve_heavy = [&](unsigned int& pos_out) -> bool {
    if (*reinterpret_cast<volatile int*>(&W.handoff_pause) > 0) return false;
    unsigned int pos = ld_volatile_u32(hq.ctrl + 1);
    if (pos - ld_volatile_u32(hq.ctrl + 0) >= (unsigned int)W.queue_limit) {
      W.handoff_pause = kHandoffPause;
      return false;
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
      const int dif = (int)(ld_volatile_u32(hq.ready + (pos & (hq.cap - 1u))) - pos);
      if (dif == 0) {
        const unsigned int seen = atomicCAS(hq.ctrl + 1, pos, pos + 1u);
        if (seen == pos) {
          pos_out = pos;
          return true;
        }
        pos = seen;
      } else if (dif < 0) {
        return false;
      } else {
        pos = ld_volatile_u32(hq.ctrl + 1);
      }
    }
    return false;
  };
  // Append the live slots of this warp to the next scan list (one shared-memory atomic per warp).
  auto append = [&](bool alive, bool own, int slot) {
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    const unsigned mo = __ballot_sync(0xffffffffu, alive && own);
    if (m != 0u) {
      int base = 0;
      if (lane == 0) {
        base = atomicAdd(&W.n_next, __popc(m));
        if (mo != 0u) atomicAdd(&W.n_own, __popc(mo));
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      if (alive) W.list_a[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
    }
  };
  auto store_ray = [&](int slot, const Ray& ray, V3 att, V3 acc, Rng rng, int bounce, int sample) {
    W.ox[slot] = ray.o.x, W.oy[slot] = ray.o.y, W.oz[slot] = ray.o.z;
    W.dx[slot] = ray.d.x, W.dy[slot] = ray.d.y, W.dz[slot] = ray.d.z, W.tm[slot] = ray.tm;
    W.att_x[slot] = att.x, W.att_y[slot] = att.y, W.att_z[slot] = att.z;
    W.acc_x[slot] = acc.x, W.acc_y[slot] = acc.y, W.acc_z[slot] = acc.z;
    W.rng[slot] = rng.s, W.bounce[slot] = bounce, W.sample[slot] = sample;
    W.best64[slot] = kNoHit64;
    // A ray that can meet a NaN (pt_prims.cuh, "the scan in vector order") switches the CTA's NEXT round to the per-ray
    // sequential scan -- closest_hit(), which knows what to do with it -- instead of burdening the phases with a test per
    // ray: such rays are one in many millions.
#ifndef PT_X1
    if (needs_in_order(sc, blob_base, ray)) {
      W.in_order = 1;
#ifdef PT_LAST_PIXEL  // (experiments) which rays ask for the vector-order scan
      if (p.counters && atomicAdd(p.counters + 17, 1ull) < 12ull)
        printf("in-order ray: o %g %g %g d %g %g %g bounce %d sample %d\n", ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, bounce, sample);
#endif
    }
#endif
  };
  auto load_ray = [&](int slot) -> Ray {
    Ray ray;
    ray.o = v3(W.ox[slot], W.oy[slot], W.oz[slot]);
    ray.d = v3(W.dx[slot], W.dy[slot], W.dz[slot]);
    ray.tm = W.tm[slot];
    return ray;
  };
  // BOXES for one ray and the chunks [cb, cb + nb) of one sphere group: every crossed chunk becomes an item;
  // what does not fit into the item list is scanned here and now (`inl`).
  auto emit_items = [&](int slot, const Ray& ray, const CullRay& cr, const float4* boxes, bool moving, int cb, int nb, float f,
                        float a, Best& inl) {
    uint32_t hits = chunk_hits<kSmem>(boxes, cb, nb, cr, kInf);
    if (hits == 0u) return;
    const int cnt = __popc(hits);
    int at = atomicAdd(moving ? &W.n_items_m : &W.n_items_s, cnt);
    const int cap = moving ? kWaveItems - W.cap_items_s : W.cap_items_s;  // this kind's share of the list this round
    // static items grow from the front, moving ones from the back, each within its fixed share of the list
    while (hits) {
      const int top = 31 - __clz((int)hits);
      hits &= ~(1u << top);
      const int chunk = cb + (nb - 1 - top);
      if (at < cap) {
        W.items[moving ? kWaveItems - 1 - at : at] = make_uint2((uint32_t)slot | ((uint32_t)chunk << 10), __float_as_uint(f));
      } else {
        if (p.counters) atomicAdd(p.counters + 15, 1ull);  // stats: items scanned in place (tests check that it happens)
        if (moving)
          scan_chunk<kSmem, true, kSphereChunk, 1>(sc, sv.moving(), sc.moving_aux, chunk, (lane & (kSphereChunk - 1)), ray, a, filter_a(a), f, G_MOVING_SPHERE, inl);
        else
          scan_chunk<kSmem, false, kSphereChunk, 1>(sc, sv.sphere(), sc.sphere_aux, chunk, (lane & (kSphereChunk - 1)), ray, a, filter_a(a), 0.f, G_SPHERE, inl);
      }
      ++at;
    }
  };

  // (ray, leaf) items of the flat groups with a tree: flat_item_scan / wave_emit_flat above the kernel
  uint2* const tree_lists = reinterpret_cast<uint2*>(smem_stage + ((staged + 15u) & ~15u));
  auto run_flat_item = [&](const uint2 it, int first, int step) { wave_run_flat<kSmem>(&W, it, first, step); };
  // BOXES: the top level of both trees (boxes, grazing index) of one flat group for one ray
  auto emit_flat_items = [&](int slot, int tg, uint32_t graze_from, uint32_t graze_to, unsigned long long& v) {
    v = wave_tree_roots<kSmem>(&W, slot, tg, graze_from, graze_to, v);
  };
  // a tree pass: the children of one (ray, node) item
  auto expand_node_item = [&](const uint2 it, int pass, int sub, int width) { wave_tree_node_item<kSmem>(&W, it, pass, sub, width); };

  int mode = 0;  // 0: this CTA's share of the pixel queue (none for an express CTA); 1: hand-off service
  if (tid == 0) {
    wave_build_tables<kSmem, kTrees>(&W, &sc, sv.groups(), sv.trees(), p.tree_list_bytes);
    if (kTrees)
      W.tctx = WaveTreeCtx { &sc, blob_base, tree_lists, p.tree_spill ? p.tree_spill + (size_t)blockIdx.x * kTreeSpillLists * p.tree_spill_cap : nullptr,
                             p.tree_spill ? (int)p.tree_spill_cap : 0, p.counters };
  }
  if (tid < 8) W.counts[tid] = 0;
  if (tid == 0) W.cap_items_s = kWaveItemsStatic;
  if (tid == 0) W.n_next = 0, W.n_own = 0, W.n_items_s = 0, W.n_items_m = 0, W.free_count = 0, W.pixel_dry = 0, W.service = 0, W.in_order = 0;
  __syncthreads();

  if (!express) {  // (an express CTA goes straight to the hand-off service, whose first source is its reserved tiles)
    // ---- start: every pool slot (up to this CTA's fair share of the image) takes a pixel
    const int cap = p.pool_cap;
    for (int s0 = (tid >> 5) * 32; s0 < kWavePool; s0 += kWaveThreads) {
      const int slot = s0 + lane;
      bool alive = false;
      if (slot < cap) {
        uint32_t pixq;
        Rng rng;
        int px, py, sample0;
        V3 acc0;
        if (next_pixel(pixq, rng, px, py, acc0, sample0)) {
          Ray ray;
          camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
          store_ray(slot, ray, v3(1.f, 1.f, 1.f), acc0, rng, 0, sample0);
          W.pix[slot] = pixq, W.scans[slot] = express ? -1 : 0;  // (an express CTA never hands a pixel off)
          alive = true;
        }
      }
      append(alive, !express, slot);
    }
    __syncthreads();
  }

  // ---- DEEP WARPS (-DPT_DEEP_WARPS; an experiment, compiled out by default).  The deepest pixels of an image are
  // serial chains of thousands of bounces (a pixel's samples are strictly serial, render.hpp:94-101), so what bounds
  // the frame is the LATENCY of one bounce: 17 000 cycles in a short round of the phase machine (four CTA barriers,
  // every phase waiting for its slowest thread).  With the macro the warps of an express CTA leave the phase machine:
  // each traces ONE pixel at a time, all 32 lanes holding the pixel's state and splitting every scan
  // (closest_hit_warp), with shuffles and ballots instead of barriers and shared-memory lists.  Same device functions,
  // same bits (the GPU suite passes with it) -- but MEASURED SLOWER: 28 such warps per SM share the issue slots, a
  // bounce costs each of them ~2 000 warp instructions for ONE ray, and the express CTAs' throughput falls below what
  // the hand-off queue brings (default scene: 38.8 ms per frame against 29.1 ms, pixels waiting 13 ms in the queue; with
  // 24 / 32 express CTAs 35.4 / 34.5 ms).  A short round moves ~35 rays in 17 000 cycles: the same throughput per SM, and
  // it is throughput the express CTAs lack, not only latency.
#ifdef PT_DEEP_WARPS
  const bool deep = express && p.order_mode != 2;
#else
  const bool deep = false;
#endif
  bool go_deep = deep;  // (a regular CTA joins them when its own pixels have run out: -DPT_SERVICE_DEEP)

  for (unsigned int round = 0; !deep; ++round) {
    if (mode == 1) {
      // ---- hand-off service: fill the free pool slots from the global queue (polled every 8th round: a poll is
      // two round trips to L2, and these rounds are the frame's critical path)
      const int room = W.free_count, waiting = W.n_next;
      if (room > 0 && ((round & 7u) == 0u || waiting == 0)) {
        __syncthreads();  // everybody has read the two (and decided alike) before anybody changes them
        bool got = false;
        unsigned int got_pos = 0u;
        if (tid < room) {
          // an express CTA's first source is the head of the LPT order (reserved for it), then the hand-off queue -- which
          // it serves from the start, whenever it has room
          uint32_t pixq;
          Rng rng;
          int px, py, sample0;
          V3 acc0;
          if (express && next_pixel(pixq, rng, px, py, acc0, sample0)) {
            const int slot = (int)W.free_list[atomicSub(&W.free_count, 1) - 1];
            Ray ray;
            camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
            store_ray(slot, ray, v3(1.f, 1.f, 1.f), acc0, rng, 0, sample0);
            W.pix[slot] = pixq, W.scans[slot] = -1;  // (never handed off again)
            W.list_a[atomicAdd(&W.n_next, 1)] = (unsigned short)slot;
          } else {
            got = take_heavy(got_pos);
          }
        }
        if (got) {
          const int slot = (int)W.free_list[atomicSub(&W.free_count, 1) - 1];
          const float* e = hq.entries + (size_t)(got_pos & (hq.cap - 1u)) * kHeavyEntryWords;
          Ray ray;
          ray.o = v3(__ldcg(e + 4), __ldcg(e + 5), __ldcg(e + 6));
          ray.d = v3(__ldcg(e + 7), __ldcg(e + 8), __ldcg(e + 9));
          ray.tm = __ldcg(e + 10);
          store_ray(slot, ray, v3(__ldcg(e + 11), __ldcg(e + 12), __ldcg(e + 13)), v3(__ldcg(e + 14), __ldcg(e + 15), __ldcg(e + 16)),
                    Rng { __float_as_uint(__ldcg(e + 1)) }, __float_as_int(__ldcg(e + 3)), __float_as_int(__ldcg(e + 2)));
          W.pix[slot] = __float_as_uint(__ldcg(e + 0)), W.scans[slot] = -1;
          W.list_a[atomicAdd(&W.n_next, 1)] = (unsigned short)slot;
          if (p.counters) {  // stats: how long did the pixel wait in the queue
            const unsigned long long waited = (uint32_t)((uint32_t)globaltimer_ns() - __float_as_uint(__ldcg(e + 17)));
            atomicAdd(p.counters + 12, waited), atomicMax(p.counters + 13, waited);
          }
          release_heavy(got_pos);
        }
        __syncthreads();
      }
    }
    const int n = W.n_next;  // rays to trace this round
    if (n == 0) {
      if (mode == 0) {
        // ---- no regular work (left): say so, then serve the hand-off queue until the whole GPU is finished
        if (p.order_mode == 2) break;  // the cost probe hands nothing off
        mode = 1;
        for (int s = tid; s < kExpressPool; s += kWaveThreads) W.free_list[s] = (unsigned short)s;
        if (tid == 0) {
          W.free_count = kExpressPool;
          __threadfence();
          atomicAdd(hq.ctrl + 2, 1u);
          if (p.counters && !express) atomicMin(p.counters + 5, globaltimer_ns()), atomicMax(p.counters + 6, globaltimer_ns());
        }
        __syncthreads();
#ifdef PT_SERVICE_DEEP
        go_deep = true;
        break;
#else
        continue;
#endif
      }
      // idle: every producer is done (the launch is cooperative: every CTA is resident and will report) and every entry
      // has been claimed => nothing will ever arrive again
      if (tid == 0)
        W.service = (ld_volatile_u32(hq.ctrl + 2) >= gridDim.x && ld_volatile_u32(hq.ctrl + 0) == ld_volatile_u32(hq.ctrl + 1)) ? 1 : 0;
      __syncthreads();
      if (W.service == 1) break;
      __nanosleep(300);
      continue;  // (the next write of W.service is behind the barrier at the loop top)
    }
    const int n_tgroups = kTrees ? W.n_tgroups : 0;  // flat groups with a tree in front of the first medium
    const int units_per_ray = W.n_blocks + 2 * n_tgroups;  // short rounds: a ray's sphere chunks in blocks, its flat trees (boxes, grazing index) one by one
    const bool fine = n <= kFineRays && W.blocks_ok && units_per_ray > 0;
#ifdef PT_PHASE_TIMING
#ifdef PT_PHASE_EXPRESS_ONLY  // time the express CTAs' short rounds only (the frame's critical path)
#define PT_PHASE_WHO (express && mode == 1)
#elif defined(PT_PHASE_TAIL_ONLY)  // ... or the regular CTAs' rounds after the pixel queue has run dry (the frame's tail)
#define PT_PHASE_WHO (!express && mode == 0 && *reinterpret_cast<volatile int*>(&W.pixel_dry) != 0)
#else
#define PT_PHASE_WHO true
#endif
    long long pt_t0 = clock64();
#define PT_PHASE(k)                                                                                   \
  if (tid == 0 && p.counters && PT_PHASE_WHO) {                                                                       \
    const long long now = clock64();                                                                  \
    atomicAdd(p.counters + 16 + (k), (unsigned long long)(now - pt_t0));                              \
    pt_t0 = now;                                                                                      \
  }
#else
#define PT_PHASE(k)
#endif
    if (mode == 1 && tid == 0 && p.counters) atomicAdd(p.counters + 8, 1ull), atomicAdd(p.counters + 9, (unsigned long long)n);  // stats

    // ---- BOXES (or, with spheres behind a medium -- or a ray that can meet a NaN in this round -- the whole sequential scan)
#ifdef PT_X2
    const bool seq_round = sequential_scan;
#else
    const bool seq_round = sequential_scan || W.in_order != 0;  // (W.in_order: stable since the barrier before `n` was read)
#endif
    if (seq_round) {
      for (int e = tid; e < n; e += kWaveThreads) {
        wave_sequential_scan<kSmem, kTrees>(&W, &sc, blob_base, (int)W.list_a[e]);
      }
    } else {
      // one work unit = one ray x (all its chunks | a block of kFineBoxes chunks)
      const int n_units = fine ? n * units_per_ray : n;
      for (int w = tid; w < n_units; w += kWaveThreads) {
        const int e = fine ? w / units_per_ray : w;
        const int slot = (int)W.list_a[e];
        const Ray ray = load_ray(slot);
        const int unit = fine ? w - e * units_per_ray : 0;
        CullRay cr;
        const int cull_set = make_cull_ray(sc, ray, cr);
        const float a = vdot(ray.d, ray.d);  // sphere.hpp:69
        Best inl { kInf, -1 };
        unsigned long long v = kNoHit64;
        if (kTrees && fine && unit >= W.n_blocks) {  // short rounds: one flat tree of the ray
          const int tu = unit - W.n_blocks;
          emit_flat_items(slot, tu >> 1, (uint32_t)(tu & 1), (uint32_t)(tu & 1) + 1u, v);
          if (v != kNoHit64) atomicMin(&W.best64[slot], v);
          continue;
        }
        const int4 blk = fine ? W.blocks[unit] : make_int4(0, 0, 0, 0);
        for (int gi = fine ? blk.x : 0; gi < (fine ? blk.x + 1 : (int)sc.n_groups); ++gi) {
          const Group g = sv.groups()[gi];
          if (g.type != G_SPHERE && g.type != G_MOVING_SPHERE) continue;
          const bool moving = g.type == G_MOVING_SPHERE;
          const float4* boxes = moving ? sv.moving_box() + cull_set * 2 * (int)sc.n_moving_chunks
                                       : sv.sphere_box() + cull_set * 2 * (int)sc.n_sphere_chunks;
          const float f = moving ? fdiv(fsub(ray.tm, g.time0), g.den) : 0.f;
          // the group's outsized spheres (a ground sphere ...) are tested right here, one by one: their chunks are never
          // culled and mostly padding (the same sphere for every lane: broadcast loads)
          // (short rounds leave them to SPHERES as items of their never-culled chunks: a shorter chain here)
          const int open_chunks = fine ? 0 : (g.n_open + kSphereChunk - 1) / kSphereChunk;
          if (g.n_open > 0 && !fine) {
            const float af = filter_a(a);
            for (int i = g.begin; i < g.begin + g.n_open; ++i) {
              float cx, cy, cz, r2f;
              if (moving) {
                const float4* ps = sv.moving() + moving_slot(i);
                if ((int)sphere_filter_bits<kSmem, true>(ps, f, ray, af) < 0) {
                  sphere_center<kSmem, true>(ps, f, cx, cy, cz, r2f);
                  sphere_roots_scan(sc, inl, ray, a, cx, cy, cz, exact_r2(sc.moving_aux, i), make_id(G_MOVING_SPHERE, i));
                }
              } else {
                const float4* ps = sv.sphere() + sphere_slot(i);
                if ((int)sphere_filter_bits<kSmem, false>(ps, f, ray, af) < 0) {
                  sphere_center<kSmem, false>(ps, f, cx, cy, cz, r2f);
                  sphere_roots_scan(sc, inl, ray, a, cx, cy, cz, exact_r2(sc.sphere_aux, i), make_id(G_SPHERE, i));
                }
              }
            }
          }
          const int c_group = g.begin / kSphereChunk + open_chunks;  // the first chunk with a box
          const int c_first = fine ? max(blk.y, c_group) : c_group;
          const int c_end = fine ? blk.y + blk.z : (g.begin + g.count) / kSphereChunk;
          if (inl.id >= 0) {
            const unsigned long long w64 = pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, inl);
            if (w64 < v) v = w64;
            inl.t = kInf, inl.id = -1;
          }
          for (int cb = c_first; cb < c_end; cb += 32) {
            emit_items(slot, ray, cr, boxes, moving, cb, min(32, c_end - cb), f, a, inl);
            if (inl.id >= 0) {  // scanned in place: fold into the ray's winner
              const unsigned long long w64 = pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, inl);
              if (w64 < v) v = w64;
              inl.t = kInf, inl.id = -1;
            }
          }
        }
        if (!fine && W.n_flats != 0) {
          // many rays: the flat objects in front of the first medium one after the other, here (few rays: one thread per
          // (ray, object) in SPHERES)
          Best fb { kInf, -1 };
          for (int gi = 0; gi < W.first_late_group; ++gi) {
            const Group g = sv.groups()[gi];
            if ((g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) && !(kTrees && has_tree(sc, g))) scan_flat_group<kSmem>(sc, sv, g, ray, g.begin, 1, fb);
          }
          if (fb.id >= 0) {
            const unsigned long long w64 = pack_winner(fb.t, key_of(sc, fb.id));
            if (w64 < v) v = w64;
          }
        }
        if (!fine && n_tgroups != 0)
          for (int tg = 0; tg < n_tgroups; ++tg) emit_flat_items(slot, tg, 0u, 2u, v);
        if (v != kNoHit64) atomicMin(&W.best64[slot], v);
      }
    }
    __syncthreads();
    if (tid == 0) W.in_order = 0;  // (everybody has read it; SHADE and the next intake set it for the next round)
    if (kTrees && !seq_round) {
      // ---- TREE PASSES: one level of the flat groups' trees per pass, one thread per (ray, node) item
      for (int pass = 0; pass < W.tree_passes; ++pass) {
        const int n_items = min(W.tl_n[pass], W.tl_cap[pass] + W.tctx.spill_cap);
        auto item_at = [&](int i) { return i < W.tl_cap[pass] ? tree_lists[W.tl_off[pass] + i] : W.tctx.spill[pass * W.tctx.spill_cap + (i - W.tl_cap[pass])]; };
        // a node item's 16 children go to as many threads as the pass can keep busy: one thread per item in a full round,
        // up to one per child in a short one (whose cost is the per-thread chain: 16 boxes from L2, four at a time;
        // measured on the mesh: 29 of the 41 us of an express CTA's round were BOXES and the tree passes).  Never finer than
        // needed: one thread per child in FULL rounds costs 1.5 x (every thread repeats the item's set-up).
        int per = 1;
        while (per < kTreeFan && 2 * per * n_items <= kWaveThreads) per *= 2;
        const int width = kTreeFan / per;
        for (int w = tid; w < n_items * per; w += kWaveThreads) expand_node_item(item_at(w / per), pass, (w & (per - 1)) * width, width);
        __syncthreads();
      }
    }
    PT_PHASE(0)

    // ---- SPHERES: one thread per (ray, chunk) item, or per quarter of one
    if (!seq_round) {
      const int n_s = min(W.n_items_s, W.cap_items_s), n_m = min(W.n_items_m, kWaveItems - W.cap_items_s), n_f = kTrees ? min(W.tl_n[2], W.tl_cap[2] + W.tctx.spill_cap) : 0;
      auto leaf_item = [&](int i) { return i < W.tl_cap[2] ? tree_lists[W.tl_off[2] + i] : W.tctx.spill[2 * W.tctx.spill_cap + (i - W.tl_cap[2])]; };
#ifdef PT_PHASE_TIMING
      if (tid == 0 && p.counters) atomicAdd(p.counters + 23, (unsigned long long)(n_s + n_m));
#endif
      if (!fine) {
        // one index space for both kinds (a thread's items follow each other without a pass boundary in between); the
        // moving items start at a warp boundary so that a warp runs one kind's code
        const int m_base = (n_s + 31) & ~31;
        const int fl_base = (m_base + n_m + 31) & ~31;  // (ray, leaf) items of the flat trees: one thread per leaf
#if PT_LEAF_SUB  // one thread per (ray, leaf, element), like the short rounds
        if (kTrees)
          for (int w = tid; w < kFlatChunk * n_f; w += kWaveThreads) run_flat_item(leaf_item(w / kFlatChunk), w % kFlatChunk, kFlatChunk);
        for (int i = tid; i < m_base + n_m; i += kWaveThreads) {
          if (i >= n_s && i < m_base) continue;
#else
        for (int i = tid; i < fl_base + n_f; i += kWaveThreads) {
          if ((i >= n_s && i < m_base) || (i >= m_base + n_m && i < fl_base)) continue;
          if (kTrees && i >= fl_base) {
            run_flat_item(leaf_item(i - fl_base), 0, 1);
            continue;
          }
#endif
          const bool moving = i >= m_base;
          const uint2 it = W.items[moving ? kWaveItems - 1 - (i - m_base) : i];
          const int slot = (int)(it.x & 1023u);
          const Ray ray = load_ray(slot);
          const float a = vdot(ray.d, ray.d);
          Best b { kInf, -1 };
          if (moving)
            scan_chunk<kSmem, true, kSphereChunk, 1>(sc, sv.moving(), sc.moving_aux, (int)(it.x >> 10), (lane & (kSphereChunk - 1)), ray, a, filter_a(a),
                                                     __uint_as_float(it.y), G_MOVING_SPHERE, b);
          else
            scan_chunk<kSmem, false, kSphereChunk, 1>(sc, sv.sphere(), sc.sphere_aux, (int)(it.x >> 10), (lane & (kSphereChunk - 1)), ray, a, filter_a(a), 0.f,
                                                      G_SPHERE, b);
          if (b.id >= 0) atomicMin(&W.best64[slot], pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, b));
        }
      } else {
        // Short rounds, ONE unit per thread where possible: a quarter of an item (its kParts threads are neighbouring lanes:
        // lane & 15 = 4 m + q takes slots 4 m + q + 4 k, like a team of 4), or one (ray, flat object) pair for the
        // rectangles, triangles and box sides in front of the first medium (object-major: a warp tests one object).
        constexpr int kParts = kSphereChunk / kFineQuarter;
        static_assert(kWaveThreads % kParts == 0 && kParts == 4, "an item's threads must be lanes 4 m .. 4 m + 3");
        const int n_quarters = kParts * (n_s + n_m);
        const int f_base = (n_quarters + 31) & ~31;
        const int n_flats = W.n_flats;
        const int fl_base = (f_base + n * n_flats + 31) & ~31;  // flat-tree items: one thread per (leaf, element)
        for (int w = tid; w < fl_base + kFlatChunk * n_f; w += kWaveThreads) {
          if (kTrees && w >= fl_base) {
            run_flat_item(leaf_item((w - fl_base) / kFlatChunk), (w - fl_base) % kFlatChunk, kFlatChunk);
          } else if (w < n_quarters) {
            const int i = w / kParts;
            const bool moving = i >= n_s;
            const uint2 it = W.items[moving ? kWaveItems - 1 - (i - n_s) : i];
            const int slot = (int)(it.x & 1023u);
            const Ray ray = load_ray(slot);
            const float a = vdot(ray.d, ray.d);
            Best b { kInf, -1 };
            if (moving)
              scan_chunk<kSmem, true, kFineQuarter, kParts>(sc, sv.moving(), sc.moving_aux, (int)(it.x >> 10), (lane & (kSphereChunk - 1)), ray, a, filter_a(a),
                                                            __uint_as_float(it.y), G_MOVING_SPHERE, b);
            else
              scan_chunk<kSmem, false, kFineQuarter, kParts>(sc, sv.sphere(), sc.sphere_aux, (int)(it.x >> 10), (lane & (kSphereChunk - 1)), ray, a, filter_a(a),
                                                             0.f, G_SPHERE, b);
            if (b.id >= 0) atomicMin(&W.best64[slot], pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, b));
          } else if (w >= f_base && w < f_base + n * n_flats) {
            const int j = (w - f_base) / n;
            const int2 fo = W.flats[j];
            const int slot = (int)W.list_a[(w - f_base) - j * n];
            const Ray ray = load_ray(slot);
            Best b { kInf, -1 };
            if ((fo.x & 255) == G_BOX) {
              const float4 p0 = ld4<kSmem>(sv.box() + 2 * fo.y);
              const float4 p1 = ld4<kSmem>(sv.box() + 2 * fo.y + 1);
              float t, ra, rb;
              if (box_side_hit_t(ray, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), fo.x >> 8, kTMin, kInf, t, ra, rb))
                b.t = t, b.id = make_id(G_BOX, fo.y);
            } else {
              Group g {};
              g.type = fo.x, g.begin = fo.y, g.count = 1;
              scan_flat_group<kSmem>(sc, sv, g, ray, fo.y, 1, b);
            }
            if (b.id >= 0) atomicMin(&W.best64[slot], pack_winner(b.t, key_of(sc, b.id)));
          }
        }
      }
      __syncthreads();
    }
    PT_PHASE(1)

    // ---- LATE: the groups from the first constant_medium on; what happens next to the ray
    if (kTrees && tid == 0 && p.counters) {  // stats: tree-list items of this round that went to global memory (tests check that it happens)
      int spilled = 0;
      for (int k = 0; k < 3; ++k) spilled += max(0, min(W.tl_n[k], W.tl_cap[k] + W.tctx.spill_cap) - W.tl_cap[k]);
      if (spilled) atomicAdd(p.counters + 10, (unsigned long long)spilled);
    }
#ifndef PT_FIXED_ITEM_SPLIT
    // The shares of the item list follow the demand.  3/8 : 5/8 is the default scene's optimum and stays while both
    // kinds fit their share; when this round's rays asked for more of one kind (the demand is counted beyond what
    // fitted), the next round's list is split by that demand -- the room left over in halves, a shortage in proportion.
    // (The fixed shares gave an RTIOW scene without moving spheres 1 536 places for ~1 700 items a round: the rest was
    // scanned in place by the thread that found it, and the frame took 19 % longer.)
    if (tid == 0) {
      const int ds = W.n_items_s, dm = W.n_items_m;
      W.cap_items_s = ds <= kWaveItemsStatic && dm <= kWaveItemsMoving ? kWaveItemsStatic
                      : ds + dm <= kWaveItems                          ? ds + ((kWaveItems - ds - dm) >> 1)
                                                                        : (int)((unsigned)kWaveItems * (unsigned)min(ds, 1 << 18) / (unsigned)(min(ds, 1 << 18) + dm));
    }
#endif
    if (tid == 0) W.n_next = 0, W.n_own = 0, W.n_items_s = 0, W.n_items_m = 0, W.tl_n[0] = 0, W.tl_n[1] = 0, W.tl_n[2] = 0;  // (SHADE builds the next round's list)
    if (tid == 0 && W.handoff_pause > 0) --W.handoff_pause;
    for (int e = tid; e < n; e += kWaveThreads) {
      const int slot = (int)W.list_a[e];
      Best best;
      if (seq_round) {
        best.t = W.hit_t[slot], best.id = W.hit_id[slot];
      } else {
        best = unpack_winner(sc, W.best64[slot]);
        const int late = W.first_late_group;
        if (late < (int)sc.n_groups) {
          // from the first constant_medium on, in group order against the running closest hit; a medium commits
          // unconditionally and may draw from the pixel's stream (constant_medium.hpp:52-65)
          const Ray ray = load_ray(slot);
          Rng rng { W.rng[slot] };
          for (int gi = late; gi < (int)sc.n_groups; ++gi) {
            const Group g = sv.groups()[gi];
            if (g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) {
              if (kTrees && has_tree(sc, g))
                best = scan_flat_tree<kSmem>(key_table(sc), flat_trees(sc, sv, g.type), g, ray, 0, 1, best);
              else
                scan_flat_group<kSmem>(sc, sv, g, ray, g.begin, 1, best);
            } else if (g.type == G_MEDIUM) {
              float t;
              if (medium_hit_t(sc.media[g.begin], ray, kTMin, best.t, rng, t, sc.flat_cull != 0u)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
            }
          }
          W.rng[slot] = rng.s;
        }
        W.hit_t[slot] = best.t, W.hit_id[slot] = best.id;
      }
      W.scans[slot] += W.scans[slot] >= 0 ? 1 : -1;  // (a taken-over pixel counts downwards: -1 - rounds in the service)
      int kind = 0;
      if (best.id >= 0) kind = 1 + reinterpret_cast<const pt_material*>(sc.materials)[material_of(sc, best.id)].kind;
      W.list_k[kind][atomicAdd(&W.counts[kind], 1)] = (unsigned short)slot;  // "sorted" by kind as a side effect
      ++n_scans;
    }
    __syncthreads();
    PT_PHASE(2)
    PT_PHASE(3)

    // ---- SHADE: one warp per unit of up to 32 rays of ONE kind (no divergence on the material)
    int unit_base[kWaveKinds + 1], kind_count[kWaveKinds];  // 32-ray shading units in front of each kind
    {
      int units = 0;
#pragma unroll
      for (int k = 0; k < kWaveKinds; ++k) {
        kind_count[k] = W.counts[k];
        unit_base[k] = units, units += (kind_count[k] + 31) >> 5;
      }
      unit_base[kWaveKinds] = units;
    }
    const int heavy_rate = W.heavy_rate;
    for (int u = tid >> 5; u < unit_base[kWaveKinds]; u += kWaveThreads / 32) {
      int kind = 0, e = 0, e_end = 0;
#pragma unroll
      for (int k = 0; k < kWaveKinds; ++k)
        if (u >= unit_base[k] && u < unit_base[k + 1]) kind = k, e = ((u - unit_base[k]) << 5) + lane, e_end = kind_count[k];
      const bool act = e < e_end;
      const int slot = act ? (int)W.list_k[kind][e] : 0;
      bool alive = false, own = false;
      if (act) {
        Ray ray = load_ray(slot);
        const Best best { W.hit_t[slot], W.hit_id[slot] };
        V3 att = v3(W.att_x[slot], W.att_y[slot], W.att_z[slot]);
        V3 acc = v3(W.acc_x[slot], W.acc_y[slot], W.acc_z[slot]);
        Rng rng { W.rng[slot] };
        int bounce = W.bounce[slot], sample = W.sample[slot];
        uint32_t pixq = W.pix[slot];
        const int scans = W.scans[slot];
        own = scans >= 0;
        V3 contribution;
        bool new_pixel = false;
        alive = true;
        if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
          // the path ended: render.hpp:100-105
          acc = vadd(acc, contribution);
          int px, py;
          float *out_px, *state_px;
          queue_pixel(p, pixq, px, py, out_px, state_px);
          if (++sample == p.spp) {
            if (p.order_mode == 2)
              p.probe_cost[pixq] = scans;  // cost probe: how deep did one sample of this pixel go
            else
              pixel_finish(p, out_px, state_px, acc, rng, (float)p.spp);
#ifdef PT_LAST_PIXEL  // (experiments) who finishes last: {ns since the start, own, service mode, |scans| / 4}
            if (p.counters)
              atomicMax(p.counters + 16, ((globaltimer_ns() - p.counters[1]) << 24) | ((unsigned long long)own << 23) | ((unsigned long long)mode << 22) |
                                             (unsigned long long)min((scans >= 0 ? scans : -1 - scans) >> 2, (1 << 22) - 1));
#endif
            new_pixel = true;
          } else {
            camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
            att = v3(1.f, 1.f, 1.f);
            bounce = 0;
          }
        }
        // a heavy pixel leaves for a CTA that runs short rounds, with its complete state
        if (!new_pixel && own && p.order_mode != 2 && scans > kHeavyBase + heavy_rate * (sample - p.spp_from)) {
          unsigned int i = 0u;
          if (reserve_heavy(i)) {
            float* q = hq.entries + (size_t)(i & (hq.cap - 1u)) * kHeavyEntryWords;
            __stcg(q + 0, __uint_as_float(pixq)), __stcg(q + 1, __uint_as_float(rng.s));
            __stcg(q + 2, __int_as_float(sample)), __stcg(q + 3, __int_as_float(bounce));
            __stcg(q + 4, ray.o.x), __stcg(q + 5, ray.o.y), __stcg(q + 6, ray.o.z);
            __stcg(q + 7, ray.d.x), __stcg(q + 8, ray.d.y), __stcg(q + 9, ray.d.z), __stcg(q + 10, ray.tm);
            __stcg(q + 11, att.x), __stcg(q + 12, att.y), __stcg(q + 13, att.z);
            __stcg(q + 14, acc.x), __stcg(q + 15, acc.y), __stcg(q + 16, acc.z);
            __stcg(q + 17, __uint_as_float((uint32_t)globaltimer_ns()));
            __threadfence();
            *reinterpret_cast<volatile unsigned int*>(hq.ready + (i & (hq.cap - 1u))) = i + 1u;
            if (p.counters) atomicAdd(p.counters + 7, 1ull);  // stats: pixels handed off
            new_pixel = true;
          }
        }
        if (new_pixel) {
          if (!own && p.counters) atomicMax(p.counters + 14, (unsigned long long)(-1 - scans));  // stats: longest stay in the service
          int px, py;
          alive = mode == 0 && next_pixel(pixq, rng, px, py, acc, sample);
          if (alive) {
            camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
            att = v3(1.f, 1.f, 1.f);
            bounce = 0;
            W.pix[slot] = pixq, W.scans[slot] = express ? -1 : 0;
            own = !express;
          } else if (mode == 1) {
            W.free_list[atomicAdd(&W.free_count, 1)] = (unsigned short)slot;  // refilled from the hand-off queue
          }
        }
        if (alive) store_ray(slot, ray, att, acc, rng, bounce, sample);
      }
      append(alive, own, slot);
    }
    __syncthreads();
    if (tid < 8) W.counts[tid] = 0;  // (everybody has read them; LATE of the next round is two barriers away)
    PT_PHASE(4)
#ifdef PT_PHASE_TIMING
    if (tid == 0 && p.counters && PT_PHASE_WHO) atomicAdd(p.counters + 21, 1ull), atomicAdd(p.counters + 22, (unsigned long long)n);
#endif
  }

#ifdef PT_DEEP_WARPS
  if (go_deep) {
    if (deep && tid == 0) {  // an express CTA hands nothing off: it counts as "finished with its regular work" from the start
      __threadfence();
      atomicAdd(hq.ctrl + 2, 1u);
    }
    bool have = false;
    uint32_t pixq = 0u;
    int px = 0, py = 0, sample = 0, bounce = 0;
    Rng rng { 0u };
    Ray ray { v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f), 0.f };
    V3 att = v3(1.f, 1.f, 1.f), acc = v3(0.f, 0.f, 0.f);
    for (;;) {
      if (!have) {
        // (1) the reserved head of the LPT order
        uint32_t got = 0xffffffffu;
        if (lane == 0) {
          uint32_t q;
          Rng r0;
          int x, y, s0;
          V3 a0;
          if (next_pixel(q, r0, x, y, a0, s0)) got = q;
        }
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got != 0xffffffffu) {
          pixq = got;
          float *unused, *state_px;
          queue_pixel(p, pixq, px, py, unused, state_px);
          pixel_start(p, px, py, state_px, rng, acc, sample);
          camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
          att = v3(1.f, 1.f, 1.f), bounce = 0, have = true;
        } else {
          // (2) the hand-off queue: lane 0 claims an entry, the lanes read one word each
          unsigned int hpos = 0u;
          int ok = 0;
          if (lane == 0) ok = take_heavy(hpos) ? 1 : 0;
          ok = __shfl_sync(0xffffffffu, ok, 0), hpos = __shfl_sync(0xffffffffu, hpos, 0);
          if (ok) {
            const float* e = hq.entries + (size_t)(hpos & (hq.cap - 1u)) * kHeavyEntryWords;
            const float wv = lane < kHeavyEntryWords ? __ldcg(e + lane) : 0.f;
#define PT_WORD(k) __shfl_sync(0xffffffffu, wv, (k))
            pixq = __float_as_uint(PT_WORD(0)), rng.s = __float_as_uint(PT_WORD(1));
            sample = __float_as_int(PT_WORD(2)), bounce = __float_as_int(PT_WORD(3));
            ray.o = v3(PT_WORD(4), PT_WORD(5), PT_WORD(6)), ray.d = v3(PT_WORD(7), PT_WORD(8), PT_WORD(9)), ray.tm = PT_WORD(10);
            att = v3(PT_WORD(11), PT_WORD(12), PT_WORD(13)), acc = v3(PT_WORD(14), PT_WORD(15), PT_WORD(16));
            const uint32_t queued_at = __float_as_uint(PT_WORD(17));  // (every lane takes part in every shuffle)
            if (lane == 0) {
              if (p.counters) {  // stats: how long did the pixel wait in the queue
                const unsigned long long waited = (uint32_t)((uint32_t)globaltimer_ns() - queued_at);
                atomicAdd(p.counters + 12, waited), atomicMax(p.counters + 13, waited);
              }
              release_heavy(hpos);
            }
#undef PT_WORD
            float *unused, *state_px;
            queue_pixel(p, pixq, px, py, unused, state_px);
            have = true;
          } else {
            // nothing right now: finished when every CTA has reported and every entry has been claimed
            int done = 0;
            if (lane == 0) done = (ld_volatile_u32(hq.ctrl + 2) >= gridDim.x && ld_volatile_u32(hq.ctrl + 0) == ld_volatile_u32(hq.ctrl + 1)) ? 1 : 0;
            if (__shfl_sync(0xffffffffu, done, 0)) break;
            __nanosleep(400);
            continue;
          }
        }
      }
      const Best best = closest_hit_warp<kSmem, kTrees>(sc, sv, ray, rng, lane);
      if (lane == 0) ++n_scans;
      V3 contribution;
      if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
        acc = vadd(acc, contribution);  // the path ended: render.hpp:100-105
        if (++sample == p.spp) {
          if (lane == 0) {
            int x, y;
            float *out_px, *state_px;
            queue_pixel(p, pixq, x, y, out_px, state_px);
            pixel_finish(p, out_px, state_px, acc, rng, (float)p.spp);
          }
          have = false;
        } else {
          camera_ray(cam, px, py, (float)p.width, (float)p.height, rng, ray);
          att = v3(1.f, 1.f, 1.f), bounce = 0;
        }
      }
    }
  }

#endif
  if (p.counters && lane == 0) atomicMax(p.counters + 3, globaltimer_ns());  // timeline: warp retired
  unsigned int warp_scans = n_scans;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_scans += __shfl_xor_sync(0xffffffffu, warp_scans, o);
  if (lane == 0 && p.counters) atomicAdd(p.counters, (unsigned long long)warp_scans);
}

// ---------------------------------------------------------------- LPT tile order
// How many of `grid` CTAs only serve the hand-off queue.  The express CTAs exist for LATENCY: the deepest pixel of the
// frame is a serial chain of  chain = spp x (scans per sample of a deep pixel ~ 0.6 x the deepest probed sample) x 9 us
// (a short round), while the frame as a whole takes about  frame = spp x mean scans x pixels / (grid x 30 scans / us).
// Where the chain is as long as the frame (the default scene: 800x480, ratio 1.2) 17 of 148 CTAs is the measured optimum;
// where the frame is many chains long they only need to keep the queue from piling up, and every CTA they do not
// take renders pixels.  Measured optima (tools/express_sweep.py): ratio 1.19 -> 17, 0.22 -> 8, 0.14 -> 2..4,
// 0.055 -> 4..8, 0.02 -> 2..4 of 148; the rule below is 17/148 x sqrt(ratio), at least 2.  Without a probe: 17 of 148.
__host__ __device__ inline int express_ctas(int grid, float chain_over_frame) {
  if (grid < 64) return 0;
  const float r = chain_over_frame < 0.f || chain_over_frame > 1.f ? 1.f : chain_over_frame;
  const int n = (int)(0.1149f * (float)grid * sqrtf(r) + 0.5f);
  return n < 2 ? 2 : n;
}

// One block: bin the tiles by probed cost (sum of the probes inside the tile), then list them from the
// most expensive bin to the cheapest (counting sort; the order inside a bin does not matter).
constexpr int kCostBins = 1024;
__global__ void __launch_bounds__(1024) tile_order_kernel(const int* __restrict__ probe_cost, int region_w, int region_h,
                                                           int tiles_x, int tiles_y, int* __restrict__ tile_order,
                                                           int* __restrict__ tile_bin, FrameTuning* __restrict__ tuning, int grid,
                                                           int n_express_forced, int probe_spp) {
  __shared__ int hist[kCostBins];
  __shared__ int start[kCostBins];
  __shared__ unsigned long long total, heavy;
  __shared__ int deepest;
  if (threadIdx.x == 0) total = 0ull, heavy = 0ull, deepest = 0;
  const int n_tiles = tiles_x * tiles_y;
  const int pw = (region_w + kProbeStep - 1) / kProbeStep, ph = (region_h + kProbeStep - 1) / kProbeStep;
  for (int b = threadIdx.x; b < kCostBins; b += blockDim.x) hist[b] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const int tx = t % tiles_x, ty = t / tiles_x;
    int cost = 0;
    for (int j = 0; j < kTile / kProbeStep; ++j)
      for (int i = 0; i < kTile / kProbeStep; ++i) {
        const int qx = tx * (kTile / kProbeStep) + i, qy = ty * (kTile / kProbeStep) + j;
        if (qx < pw && qy < ph) cost += probe_cost[qy * pw + qx];
      }
    const int bin = min(cost, kCostBins - 1);
    tile_bin[t] = bin;
    atomicAdd(&hist[bin], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = kCostBins - 1; b >= 0; --b) start[b] = run, run += hist[b];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) tile_order[atomicAdd(&start[tile_bin[t]], 1)] = t;
  // ---- what the probe says about the frame (FrameTuning, pt_kernel.h)
  const int n_probes = pw * ph;
  unsigned long long mine = 0ull;
  int my_max = 0;
  for (int i = threadIdx.x; i < n_probes; i += blockDim.x) mine += (unsigned long long)probe_cost[i], my_max = max(my_max, probe_cost[i]);
  atomicAdd(&total, mine), atomicMax(&deepest, my_max);
  __syncthreads();
  const unsigned long long sum = total;
  mine = 0ull;
  for (int i = threadIdx.x; i < n_probes; i += blockDim.x)
    if (4ull * sum < (unsigned long long)probe_cost[i] * (unsigned long long)n_probes) mine += (unsigned long long)probe_cost[i];  // cost > 4 x mean
  atomicAdd(&heavy, mine);
  __syncthreads();
  if (threadIdx.x == 0 && tuning) {
    FrameTuning t;
    const unsigned long long samples = (unsigned long long)max(n_probes, 1) * (unsigned long long)max(probe_spp, 1);
    t.mean_scans_x1000 = (int)(1000ull * sum / samples);
    t.heavy_share_x1000 = (int)(1000ull * heavy / (sum ? sum : 1ull));
    // the deepest probed sample -- or, from a probe of several samples, what a pixel as deep as the deepest probed PIXEL
    // would show in one (a deep pixel averages 0.6 of its deepest sample: the default scene's 31.5 of 50)
    const int deepest_sample = probe_spp > 1 ? (int)((float)deepest / (0.6f * (float)probe_spp)) : deepest;
    t.max_scans = deepest_sample;
    // a pixel is heavy when it spends more than four times the frame's average per sample (the default scene: 2.6 scans
    // per sample, rate 10); never below the default
    t.heavy_rate = max(kHeavyRate, (int)((4ull * sum + samples / 2ull) / samples));
#ifdef PT_FORCE_HEAVY_RATE  // (experiments)
    t.heavy_rate = PT_FORCE_HEAVY_RATE;
#endif
    const float chain_over_frame = 162.f * (float)grid * (float)deepest_sample / (fmaxf(1e-3f * (float)t.mean_scans_x1000, 1e-3f) * (float)region_w * (float)region_h);
    t.n_express = n_express_forced >= 0 ? n_express_forced : express_ctas(grid, chain_over_frame);
    t.pad[0] = t.pad[1] = t.pad[2] = 0;
    *tuning = t;
  }
}

cudaError_t launch_tile_order(const int* probe_cost, int region_w, int region_h, int tiles_x, int tiles_y,
                              int* tile_order, int* scratch, FrameTuning* tuning, int grid, int n_express_forced, int probe_spp, cudaStream_t stream) {
  tile_order_kernel<<<1, 1024, 0, stream>>>(probe_cost, region_w, region_h, tiles_x, tiles_y, tile_order, scratch, tuning, grid,
                                            n_express_forced, probe_spp);
  return cudaGetLastError();
}
int wave_grid(int device) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * PT_WAVE_BLOCKS_PER_SM;
}

// ---------------------------------------------------------------- framebuffer resolve
// Pixels finish one at a time, in the LPT tile order, and a packed float3 pixel is only 4-byte aligned: the render
// kernels therefore write each finished pixel with ONE 16-byte store into a float4 staging image, and this pass
// turns four staged pixels into three 16-byte stores of the caller's rows (render.hpp:102-105's fb[y][x] = ...),
// a warp writing 1 536 contiguous bytes -- to local HBM or, at N > 1, straight into rank 0's peer-mapped framebuffer.
__global__ void __launch_bounds__(256) resolve_fb_kernel(const float4* __restrict__ stage, int quads_per_row, int h, float* __restrict__ out,
                                                         long long out_row_pitch) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)quads_per_row * h) return;
  const int row = (int)(idx / quads_per_row), q = (int)(idx - (long long)row * quads_per_row);
  const float4* s = stage + ((long long)row * quads_per_row + q) * 4;
  const float4 a = s[0], b = s[1], c = s[2], d = s[3];
  float4* o = reinterpret_cast<float4*>(out + (long long)row * out_row_pitch + 12ll * q);
  o[0] = make_float4(a.x, a.y, a.z, b.x);
  o[1] = make_float4(b.y, b.z, c.x, c.y);
  o[2] = make_float4(c.z, d.x, d.y, d.z);
}
cudaError_t launch_resolve_fb(const float* stage, int w, int h, float* out, long long out_row_pitch, cudaStream_t stream) {
  const long long n = (long long)(w / 4) * h;
  if (n <= 0) return cudaSuccess;
  resolve_fb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(stage), w / 4, h, out, out_row_pitch);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- launch
int max_smem_blob_bytes(int device) {
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  return optin - 1024;
}

cudaError_t launch_render(const RenderParams& p, int device, int grid_override, cudaStream_t stream,
                          LaunchInfo* info) {
  int sms = 0;
  cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) return err;
  RenderParams q = p;
  const unsigned long long pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
  if (p.kernel_kind == 0) {
    // ---- wavefront kernel: one CTA per SM, ray pool + (when it fits) the scan blob in shared memory
    // shared memory: the ray pool, and in front of it the scan blob with the side tables (else the blob alone, else nothing)
    // and behind them, for scenes with flat trees, the lists of the breadth-first tree passes.  The lists continue in global
    // memory (RenderParams::tree_spill), so they are kept SMALL: a scene that does not fit shared memory streams its nodes
    // and triangles through L1, and every KB of shared memory is a KB less of it (the 10 002-triangle mesh at 64 spp: 426 ms
    // with 98 KB of lists, 387 with 64, 380 with 32, 385 with 16).
    const size_t pool_bytes = (sizeof(WavePool) + 127u) / 128u * 128u;
    const bool trees = p.scene.n_trees != 0u && p.scene.flat_cull != 0u;
#ifndef PT_TREE_LIST_MAX_KB
#define PT_TREE_LIST_MAX_KB 32
#endif
#ifndef PT_TREE_LIST_MIN_KB
#define PT_TREE_LIST_MIN_KB 32
#endif
    constexpr long long kMinTreeListBytes = PT_TREE_LIST_MIN_KB << 10, kMaxTreeListBytes = PT_TREE_LIST_MAX_KB << 10;
    const long long room = (long long)max_smem_blob_bytes(device) - (long long)sizeof(SceneDesc) - (long long)pool_bytes;
    const long long for_scene = room - (trees ? kMinTreeListBytes : 0);
    q.staged_bytes = (long long)p.scene.stage_bytes <= for_scene ? p.scene.stage_bytes : (long long)p.scene.blob_bytes <= for_scene ? p.scene.blob_bytes : 0u;
    const size_t staged_al = ((size_t)q.staged_bytes + 15u) & ~(size_t)15u;
    q.tree_list_bytes = trees ? (unsigned int)(std::min<long long>(room - (long long)staged_al, kMaxTreeListBytes) & ~15ll) : 0u;
    const bool smem = q.staged_bytes != 0u;
    const size_t dyn = pool_bytes + staged_al + q.tree_list_bytes;
    auto kernel = smem ? (trees ? render_wave_kernel<true, true> : render_wave_kernel<true, false>)
                       : (trees ? render_wave_kernel<false, true> : render_wave_kernel<false, false>);
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (err != cudaSuccess) return err;
    // every CTA must be resident at once: the express warps wait for all CTAs to report
    int resident = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kWaveThreads, dyn);
    if (err != cudaSuccess) return err;
    if (resident < 1) return cudaErrorLaunchOutOfResources;
    if (resident > PT_WAVE_BLOCKS_PER_SM) resident = PT_WAVE_BLOCKS_PER_SM;
    int grid = grid_override > 0 ? grid_override : sms * resident;
    if (grid > sms * resident) grid = sms * resident;
    const unsigned long long share = (pixels + (unsigned long long)grid - 1ull) / (unsigned long long)grid;
    q.pool_cap = (int)(share < 32ull ? 32ull : (share > (unsigned long long)kWavePool ? (unsigned long long)kWavePool : share));
    // a few CTAs only serve the hand-off queue (short rounds for the deepest pixels of the image)
    q.n_express = p.n_express >= 0 ? p.n_express : express_ctas(grid, -1.f);  // (a frame with a cost probe brings its own: FrameTuning)
    if (q.n_express >= grid) q.n_express = grid - 1;
    if (p.order_mode == 2) {
      q.n_express = 0;
      q.n_positions = (unsigned long long)((p.region.w + kProbeStep - 1) / kProbeStep) *
                      (unsigned long long)((p.region.h + kProbeStep - 1) / kProbeStep);
    } else if (p.order_mode == 1) {
      q.n_positions = (unsigned long long)p.tiles_x * (unsigned long long)p.tiles_y * (unsigned long long)(kTile * kTile);
    } else {
      q.n_positions = pixels;
    }
    // with the LPT order the express CTAs start on the most expensive tiles, one pixel per pool slot
    q.express_positions = 0;
    if (p.order_mode == 1) {
      q.express_positions = (unsigned long long)q.n_express * (unsigned long long)kExpressPool;
      if (q.express_positions > q.n_positions) q.express_positions = q.n_positions;
    }
    {
      const unsigned long long sh = (q.n_positions + (unsigned long long)grid - 1ull) / (unsigned long long)grid;
      q.pool_cap = (int)(sh < 32ull ? 32ull : (sh > (unsigned long long)kWavePool ? (unsigned long long)kWavePool : sh));
    }
#ifdef PT_TREE_POOL  // (experiments.  While the tree lists lived in shared memory alone, fewer rays in flight -- 640 -- kept
    // a round's items inside them; since they continue in global memory the full pool is faster again: 341 vs 351 ms on the mesh)
    if (trees && q.pool_cap > PT_TREE_POOL) q.pool_cap = PT_TREE_POOL;
#endif
    // pixel-order permutation pos -> (pos * scramble) mod pixels: a multiplier near pixels / golden ratio,
    // made coprime with the pixel count so that it is a bijection
    unsigned long long mul = (unsigned long long)((double)pixels * 0.6180339887498949) | 1ull;
    auto gcd = [](unsigned long long a, unsigned long long b) {
      while (b) {
        const unsigned long long t = a % b;
        a = b, b = t;
      }
      return a;
    };
    while (pixels > 1 && gcd(mul % pixels, pixels) != 1ull) mul += 2ull;
    q.scramble = pixels > 1 ? mul % pixels : 1ull;
    if (q.scramble == 0ull) q.scramble = 1ull;
    if (info) info->grid = grid, info->block = kWaveThreads, info->smem_bytes = (int)dyn, info->blocks_per_sm = 1, info->staged = smem, info->team_size = 0;
    // COOPERATIVE launch: the CTAs wait for each other (the service loop ends when every CTA has reported), so the grid
    // must be co-resident; launched this way the runtime guarantees it -- or fails the launch -- whatever else is
    // running on the device (another stream's render, another process).
#ifdef PT_NO_COOP  // (experiments only)
    kernel<<<grid, kWaveThreads, dyn, stream>>>(q);
    return cudaGetLastError();
#else
    void* args[] = { (void*)&q };
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3((unsigned)kWaveThreads), args, dyn, stream);
#endif
  }
  return launch_lane(p, device, grid_override, stream, info);
}

}  // namespace ptb
