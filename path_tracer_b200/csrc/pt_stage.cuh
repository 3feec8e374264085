// pt_stage.cuh -- shared-memory staging (cp.async.bulk + mbarrier) and the kernels' view of the scan blob.
#ifndef PT_STAGE_CUH
#define PT_STAGE_CUH
#include <stdint.h>

#include "pt_device.cuh"
#include "pt_packed.h"

namespace ptb {
namespace {

PT_DEV unsigned int ld_volatile_u32(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

#ifdef __CUDACC__
PT_DEV unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- staging
PT_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PT_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PT_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
PT_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
PT_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#endif  // __CUDACC__

// ---------------------------------------------------------------- scene view
// The kernels' view of the scan blob (in shared or global memory).  The arrays of the hot loops are plain pointers (the
// compiler keeps them in registers across the phases; forming them from the descriptor's offsets at every use cost the
// default scene 4 %, 32-bit offsets from a common base 1 % -- measured); the flat groups' trees, used by few scenes,
// are formed where they are used.
struct SceneView {
  const Group* groups_;
  const float4 *sphere_, *moving_, *rect_, *triangle_, *box_, *sphere_box_, *moving_box_;
  const unsigned char* base;
  const SceneDesc* sc;
  PT_DEV const Group* groups() const { return groups_; }
  PT_DEV const float4* sphere() const { return sphere_; }
  PT_DEV const float4* moving() const { return moving_; }
  PT_DEV const float4* rect() const { return rect_; }
  PT_DEV const float4* triangle() const { return triangle_; }
  PT_DEV const float4* box() const { return box_; }
  PT_DEV const float4* sphere_box() const { return sphere_box_; }  // chunk boxes, set 0 first
  PT_DEV const float4* moving_box() const { return moving_box_; }
  PT_DEV const Tree* trees() const { return reinterpret_cast<const Tree*>(base + sc->off_trees); }  // flat groups' box trees and grazing indices
  PT_DEV const float4* nodes() const { return reinterpret_cast<const float4*>(base + sc->off_nodes); }  // their boxes
  PT_DEV const float4* tree_ids() const { return reinterpret_cast<const float4*>(base + sc->off_tree_ids); }  // grazing index: {g, triangle} lists
};
PT_DEV SceneView scene_view(const SceneDesc& sc, const unsigned char* blob_base) {
  SceneView sv;
  sv.groups_ = reinterpret_cast<const Group*>(blob_base + sc.off_groups);
  sv.sphere_ = reinterpret_cast<const float4*>(blob_base + sc.off_sphere);
  sv.moving_ = reinterpret_cast<const float4*>(blob_base + sc.off_moving);
  sv.rect_ = reinterpret_cast<const float4*>(blob_base + sc.off_rect);
  sv.triangle_ = reinterpret_cast<const float4*>(blob_base + sc.off_triangle);
  sv.box_ = reinterpret_cast<const float4*>(blob_base + sc.off_box);
  sv.sphere_box_ = reinterpret_cast<const float4*>(blob_base + sc.off_sphere_box);
  sv.moving_box_ = reinterpret_cast<const float4*>(blob_base + sc.off_moving_box);
  sv.base = blob_base, sv.sc = &sc;
  return sv;
}

template <bool kSmem> PT_DEV float4 ld4(const float4* p) {
  if constexpr (kSmem)
    return *p;
  else
    return __ldg(p);
}

}  // namespace
}  // namespace ptb
#endif
