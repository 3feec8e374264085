// pt_packed.h -- device-side scene layout, shared by the host packer
// (pt_pack.cpp) and the kernels (pt_kernel.cu).
//
// The reference keeps the scene as an array of 624-byte std::variant objects
// and walks it sequentially per ray (render.hpp:30-51).  Here the variant list
// becomes
//   * a "scan blob": float4-packed, per-kind structure-of-arrays holding ONLY
//     what the closest-hit scan needs, ordered as a list of GROUPS.  The blob
//     is staged into shared memory once per CTA (cp.async.bulk) and every warp
//     streams it with broadcast LDS.128;
//   * small per-object side tables in global memory (material, original index
//     key, radius, normals) that are touched once per accepted hit.
//
// Groups: objects are re-ordered by kind inside SEGMENTS delimited by
// constant_medium objects (the only primitive whose result depends on the
// running closest-t and which draws RNG, constant_medium.hpp:52-65).  Within a
// segment the winner is picked by (minimum t, then maximum KEY), which
// reproduces the sequential scan's tie behaviour exactly for any visiting
// order (DESIGN.md "closest-hit order semantics"):
//   key = -1 - original_index   for spheres  (earlier sphere keeps a tie,
//                                             sphere.hpp:77,93 use strict <)
//   key = original_index        otherwise    (later object takes a tie,
//                                             rectangle.hpp:36, triangle.hpp:91)
#ifndef PT_PACKED_H
#define PT_PACKED_H

#include <stdint.h>

#ifdef __CUDACC__
#define PT_HD __host__ __device__
#else
#define PT_HD
#endif

namespace ptb {

enum GroupType : int {
  G_SPHERE = 0,         // static spheres: 1 float4 {cx, cy, cz, r*r}, in CHUNKS (below)
  G_MOVING_SPHERE = 1,  // 2 float4 {c0x, c0y, c0z, r*r} {c1x-c0x, c1y-c0y, c1z-c0z, 0}; one (time0,time1) class
  G_RECT = 2,           // 2 float4 {a0, a1, b0, b1} {k, axis, 0, 0}
  G_TRIANGLE = 3,       // 3 float4 {v0, 0} {v1-v0, 0} {v2-v0, 0}
  G_BOX = 4,            // 2 float4 {p0, 0} {p1, 0}
  G_MEDIUM = 5          // `begin` = index into the media table; closes a segment
};

// Sphere groups are stored in CHUNKS of kSphereChunk spatially close spheres (k-d ordered by the packer,
// padded with spheres that can never be hit).  Every chunk has a bounding box per origin class
// (kCullSets of them, see "chunk culling" in pt_kernel.cu); a ray only scans the chunks whose box it
// crosses, and lanes of a warp scan DIFFERENT chunks at the same time.  So that those loads do not
// collide on shared-memory banks each lane starts at its own sphere of the chunk ((lane & 15) + j),
// and so that this rotation needs no wrap-around arithmetic the 16 entries are stored TWICE in a row:
//   static chunk c : float4[32] at 32 c      (entry s and its copy at 16 + s)
//   moving chunk c : float4[64] at 64 c      ({c0, r*r} at s and 16 + s, {c1 - c0} at 32 + s and 48 + s)
constexpr int kSphereChunk = 16;
constexpr int kCullSets = 4;  // box sets per origin class; the last one is "no culling" (infinite boxes)
PT_HD inline int sphere_slot(int i) { return ((i >> 4) << 5) | (i & 15); }  // float4 index of static sphere i
PT_HD inline int moving_slot(int i) { return ((i >> 4) << 6) | (i & 15); }  // {c0, r*r}; {c1 - c0} is 32 further

// Rectangles, triangles and boxes ("flat" groups) with at least kFlatTreeMin elements are k-d ordered too, in
// leaves of kFlatChunk consecutive elements under a TREE of bounding boxes (kTreeFan children per inner
// node, at most kTreeLevels box levels, the top level a plain list): a ray only tests the leaves whose
// box -- grown by a margin relative to its distance from the ray origin -- it crosses ("flat culling" in
// pt_prims.cuh; the proof is in DESIGN.md).  Triangle groups carry a second tree over the triangles'
// scaled NORMALS (the "grazing index"): the reference's Moller-Trumbore arithmetic is noise for a ray that
// lies almost in a triangle's plane, and such (ray, triangle) pairs -- for which no geometric margin holds --
// are found there and tested exactly.
constexpr int kFlatChunk = 8;
constexpr int kTreeFan = 16;
constexpr int kTreeLevels = 3;
constexpr int kFlatTreeMin = 33;

struct Tree {  // 32 bytes
  int32_t levels;    // box levels in use (1 .. kTreeLevels); level 0 = leaves
  int32_t off[3];    // float4 index of level l's boxes in the node array ({lo} {hi} per node)
  int32_t n[3];      // nodes at level l; node j of level l + 1 covers nodes 16 j .. 16 j + 15 of level l
  int32_t leaf_ids;  // grazing index only: float4 index of the leaves' triangle lists (kFlatChunk x {g, element index; -1 = none} per leaf)
};

struct Group {  // 32 bytes
  int32_t type;
  int32_t begin;  // first element, in elements of this kind's array
  int32_t count;  // padded count for sphere groups
  float time0;    // G_MOVING_SPHERE: the class's time0
  float den;      // G_MOVING_SPHERE: time1 - time0 (sphere.hpp:55)
  int32_t n_open; // sphere groups: the first n_open elements are OUTSIZED spheres in chunks of their own that are never culled
  int32_t tree;   // flat groups: index of the group's box tree (leaf c = elements begin + 8 c ...), -1 = scanned directly
  int32_t gtree;  // triangle groups with a tree: index of the grazing index, -1 = none
};

// Unified object id carried by the scan: kind in the top bits.
constexpr int kIdShift = 27;
constexpr uint32_t kIdMask = (1u << kIdShift) - 1u;
PT_HD inline int make_id(int type, int index) { return (type << kIdShift) | index; }

struct SphereAux {  // 32 bytes; index = element index in the static / moving array
  float radius;     // sphere.hpp:81 divides by it
  float time0, den; // moving only
  int32_t material;
  int32_t key;
  int32_t pad[3];
};

struct ObjAux {  // 8 bytes: rects, boxes
  int32_t material;
  int32_t key;
};

struct TriAux {  // 32 bytes
  float nx, ny, nz;  // cross(edge1, edge2), unnormalised (triangle.hpp:96)
  int32_t material;
  int32_t key;
  int32_t pad[3];
};

struct MediumRec {  // constant_medium.hpp:80-82 with its boundary inlined
  int32_t boundary_kind;
  float neg_inv_density;  // -1 / density (constant_medium.hpp:20)
  int32_t material;
  int32_t key;
  // sphere boundary (sphere.hpp:108-113)
  float c0[3];
  float dv[3];  // center1 - center0
  float radius, r2, time0, den;
  int32_t moving;
  // box boundary (box.hpp:20-25)
  float p0[3];
  float p1[3];
  int32_t pad;
};

// Everything the kernel needs to find its data; passed by value as a kernel
// parameter.  Blob offsets are in BYTES from the blob start, 16-byte aligned.
struct SceneDesc {
  const unsigned char* blob;  // global copy of the scan blob
  uint32_t blob_bytes;        // multiple of 16
  uint32_t stage_bytes;       // blob + the side tables that follow it in the arena (aux, media, keys, object ids, materials)
  uint32_t n_groups;
  uint32_t off_groups, off_sphere, off_moving, off_rect, off_triangle, off_box;
  uint32_t off_trees, off_nodes, off_tree_ids, n_trees;  // flat groups' trees: Tree[], float4 node boxes, float4 grazing-index leaves
  // the coordinates k of every axis-aligned plane that carries a rectangle, a box side or a side of a medium's boundary box:
  // n_planes[c] sorted floats for component c (x, y, z), one after the other.  A ray with d_c == 0 that starts ON such a
  // plane has t = 0/0 there (pt_prims.cuh: needs_in_order)
  uint32_t off_planes, n_planes[3];
  uint32_t flat_cull;         // 0: ignore the trees (every flat object is tested: pt_debug_set_cull(0))
  float flat_extent;          // max |coordinate| over the flat objects under a tree (the slab test's rounding allowance)
  uint32_t n_objects;         // reference n_hittables (for work accounting)
  // chunk boxes: float4 {lo} {hi} per chunk, [kCullSets][n_chunks][2] per sphere kind; rewritten when the
  // camera's shutter interval changes (the boxes of moving spheres cover their sweep over it)
  uint32_t off_sphere_box, off_moving_box, n_sphere_chunks, n_moving_chunks;
  float cull_bound[3];        // set s serves origins with max |coordinate| <= cull_bound[s]
  uint32_t n_media_groups;    // constant_medium objects
  uint32_t n_late_sphere_groups;  // sphere groups BEHIND a constant_medium (the wavefront kernel then scans sequentially per ray)
  uint32_t n_flat_groups;     // rect / triangle / box groups
  const int32_t* object_id;   // original object index -> scan id (the wavefront kernel's winner table)
  const SphereAux* sphere_aux;
  const SphereAux* moving_aux;
  const ObjAux* rect_aux;
  const TriAux* tri_aux;
  const ObjAux* box_aux;
  const MediumRec* media;
  const int32_t* keys;    // tie-break keys of every object: keys[key_base[type] + index]
  uint32_t key_base[6];
  const void* materials;  // pt_material[]
  const void* textures;   // pt_texture[]
  const unsigned char* texture_bytes;
  uint64_t n_texture_texels;
  uint32_t n_materials, n_textures;
};

}  // namespace ptb
#endif
