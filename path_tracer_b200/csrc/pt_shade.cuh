// pt_shade.cuh -- what happens after the closest-hit scan: the hit record of the winner, textures,
// materials (material.hpp), the camera sample (camera.hpp:93-100) and the pixel queue.
#ifndef PT_SHADE_CUH
#define PT_SHADE_CUH
#include <stdint.h>

#include "pt_abi.h"
#include "pt_device.cuh"
#include "pt_kernel.h"
#include "pt_packed.h"
#include "pt_prims.cuh"

namespace ptb {
namespace {

// ---------------------------------------------------------------- shading
struct HitRec {  // hitable.hpp:8-18
  V3 p, normal;
  bool front_face;
  float u, v;
  bool has_uv;
};

// hitable.hpp:20-23
PT_DEV void set_face_normal(HitRec& rec, const Ray& r, V3 outward) {
  rec.front_face = vdot(r.d, outward) < 0.f;
  rec.normal = rec.front_face ? outward : vsub(v3(0.f, 0.f, 0.f), outward);
}

// sphere.hpp:13-24
PT_DEV void mercator(V3 p, float& u, float& v) {
  const float phi = t_atan2(p.z, p.x);
  const float theta = t_asin(p.y);
  u = fsub(1.f, fdiv(fadd(phi, kPi), fmul(2.f, kPi)));
  v = fdiv(fadd(theta, fdiv(kPi, 2.f)), kPi);
}

// texture.hpp:25 / 42-49 / 135-151
PT_DEV V3 texture_value(const SceneDesc& sc, int tex, const HitRec& rec) {
  const pt_texture* t = reinterpret_cast<const pt_texture*>(sc.textures) + tex;
  const int kind = t->kind;
  if (kind == PT_TEX_SOLID) return vld(t->color0);
  if (kind == PT_TEX_CHECKER) {
    const float sines = fmul(fmul(t_sin(fmul(10.f, rec.p.x)), t_sin(fmul(10.f, rec.p.y))), t_sin(fmul(10.f, rec.p.z)));
    return (sines < 0.f) ? vld(t->color0) : vld(t->color1);
  }
  const unsigned long long width = t->width, height = t->height;
  const float fu = fmul(t_fmod1(fmul(rec.u, t->freq)), (float)(width - 1ull));
  const float fv = fmul(fsub(1.f, t_fmod1(fmul(rec.v, t->freq))), (float)(height - 1ull));
  unsigned long long i = (unsigned long long)fu;  // truncation, texture.hpp:139-143
  unsigned long long j = (unsigned long long)fv;
  unsigned long long pix = j * width + i + t->offset;
  if (pix >= sc.n_texture_texels) pix = sc.n_texture_texels - 1ull;  // the reference would read out of bounds
  const unsigned char* td = sc.texture_bytes + pix * 3ull;
  const float scale = fdiv(1.f, 255.f);
  return v3(fmul((float)td[0], scale), fmul((float)td[1], scale), fmul((float)td[2], scale));
}

// material.hpp:62-66
PT_DEV float reflectance(float cosine, float ref_idx) {
  float r0 = fdiv(fsub(1.f, ref_idx), fadd(1.f, ref_idx));
  r0 = fmul(r0, r0);
  return fadd(r0, fmul(fsub(1.f, r0), t_pow5(fsub(1.f, cosine))));
}

// Rebuild the hit_record of the scan winner (the reference fills it inside
// hit(); only the accepted one survives, render.hpp:44-47).
PT_DEV int build_record(const SceneDesc& sc, const SceneView& sv, const Ray& r, const Best& best, HitRec& rec,
                        bool smem) {
  const int type = best.id >> kIdShift;
  const int idx = best.id & (int)kIdMask;
  rec.p = ray_at(r, best.t);
  rec.u = 0.f, rec.v = 0.f, rec.has_uv = false;
  switch (type) {
    case G_SPHERE:
    case G_MOVING_SPHERE: {  // sphere.hpp:78-88
      V3 center;
      const SphereAux* aux;
      if (type == G_SPHERE) {
        const float4 s = smem ? sv.sphere()[sphere_slot(idx)] : __ldg(sv.sphere() + sphere_slot(idx));
        center = v3(s.x, s.y, s.z);
        aux = sc.sphere_aux + idx;
      } else {
        const float4 s = smem ? sv.moving()[moving_slot(idx)] : __ldg(sv.moving() + moving_slot(idx));
        const float4 v = smem ? sv.moving()[moving_slot(idx) + 2 * kSphereChunk] : __ldg(sv.moving() + moving_slot(idx) + 2 * kSphereChunk);
        aux = sc.moving_aux + idx;
        center = moving_center(v3(s.x, s.y, s.z), v3(v.x, v.y, v.z), fdiv(fsub(r.tm, aux->time0), aux->den));
      }
      const V3 outward = vdivs(vsub(rec.p, center), aux->radius);
      set_face_normal(rec, r, outward);
      rec.has_uv = true;  // mercator(rec.normal) evaluated lazily, only for image textures
      return aux->material;
    }
    case G_RECT: {  // rectangle.hpp:42-47
      const float4 q0 = smem ? sv.rect()[2 * idx] : __ldg(sv.rect() + 2 * idx);
      const float4 q1 = smem ? sv.rect()[2 * idx + 1] : __ldg(sv.rect() + 2 * idx + 1);
      const int axis = __float_as_int(q1.y);
      const AxisSel s = axis_select(r, axis);
      const float a = fadd(s.oa, fmul(best.t, s.da));
      const float b = fadd(s.ob, fmul(best.t, s.db));
      rec.u = fdiv(fsub(a, q0.x), fsub(q0.y, q0.x));
      rec.v = fdiv(fsub(b, q0.z), fsub(q0.w, q0.z));
      const V3 n = axis == PT_AXIS_XY ? v3(0.f, 0.f, 1.f) : axis == PT_AXIS_XZ ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
      set_face_normal(rec, r, n);
      return sc.rect_aux[idx].material;
    }
    case G_TRIANGLE: {  // triangle.hpp:94-98 (u, v are not written by the reference)
      const TriAux* aux = sc.tri_aux + idx;
      set_face_normal(rec, r, v3(aux->nx, aux->ny, aux->nz));
      return aux->material;
    }
    case G_BOX: {  // box.hpp:29-50: replay the six sides to find the winning one
      const float4 p0 = smem ? sv.box()[2 * idx] : __ldg(sv.box() + 2 * idx);
      const float4 p1 = smem ? sv.box()[2 * idx + 1] : __ldg(sv.box() + 2 * idx + 1);
      float t, a, b;
      const V3 lo = v3(p0.x, p0.y, p0.z), hi = v3(p1.x, p1.y, p1.z);
      const int side = box_hit_t(r, lo, hi, kTMin, kInf, t, a, b);
      float a0, a1, b0, b1;
      V3 n;
      if (side < 2) {
        a0 = lo.x, a1 = hi.x, b0 = lo.y, b1 = hi.y, n = v3(0.f, 0.f, 1.f);
      } else if (side < 4) {
        a0 = lo.x, a1 = hi.x, b0 = lo.z, b1 = hi.z, n = v3(0.f, 1.f, 0.f);
      } else {
        a0 = lo.y, a1 = hi.y, b0 = lo.z, b1 = hi.z, n = v3(1.f, 0.f, 0.f);
      }
      rec.u = fdiv(fsub(a, a0), fsub(a1, a0));
      rec.v = fdiv(fsub(b, b0), fsub(b1, b0));
      set_face_normal(rec, r, n);
      return sc.box_aux[idx].material;
    }
    default: {  // constant_medium.hpp:72-76
      rec.normal = v3(1.f, 0.f, 0.f);
      rec.front_face = true;
      return sc.media[idx].material;
    }
  }
}

PT_DEV V3 textured(const SceneDesc& sc, int tex, HitRec& rec) {
  const pt_texture* t = reinterpret_cast<const pt_texture*>(sc.textures) + tex;
  if (t->kind == PT_TEX_IMAGE && rec.has_uv) {
    mercator(rec.normal, rec.u, rec.v);  // sphere.hpp:88
    rec.has_uv = false;
  }
  return texture_value(sc, tex, rec);
}

// render.hpp:96-99 + camera.hpp:93-100: one camera sample for pixel (px, py); 5 RNG draws
PT_DEV void camera_ray(const pt_camera& cam, int px, int py, float fwidth, float fheight, Rng& rng, Ray& ray) {
  const float u = fdiv(fadd((float)px, rng_float(rng)), fwidth);
  const float v = fdiv(fadd((float)py, rng_float(rng)), fheight);
  float dx, dy;
  rng_in_unit_disk(rng, dx, dy);
  const V3 rd = v3(fmul(cam.lens_radius, dx), fmul(cam.lens_radius, dy), fmul(cam.lens_radius, 0.f));
  const V3 cu = vld(cam.u), cv = vld(cam.v);
  const V3 offset = vadd(v3(fmul(cu.x, rd.x), fmul(cu.y, rd.x), fmul(cu.z, rd.x)),
                         v3(fmul(cv.x, rd.y), fmul(cv.y, rd.y), fmul(cv.z, rd.y)));
  const V3 origin = vld(cam.origin);
  ray.o = vadd(origin, offset);
  ray.d = vsub(vsub(vadd(vadd(vld(cam.lower_left_corner), vscale(u, vld(cam.horizontal))),
                         vscale(v, vld(cam.vertical))),
                    origin),
               offset);
  ray.tm = rng_range(rng, cam.time0, cam.time1);
}

// One iteration of get_color's depth loop after the closest-hit scan (render.hpp:58-91): sky,
// emission or scatter.  Returns true when the path ends; `contribution` is what it adds to the pixel.
PT_DEV bool shade(const SceneDesc& sc, const SceneView& sv, int depth, bool smem, const Best& best, Ray& ray,
                  Rng& rng, V3& att, int& bounce, V3& contribution) {
  contribution = v3(0.f, 0.f, 0.f);
  if (best.id < 0) {
    // background gradient, render.hpp:83-87
    const V3 ud = unit_vector(ray.d);
    const float hit_pt = fmul(0.5f, fadd(ud.y, 1.0f));
    const float w0 = fsub(1.0f, hit_pt);
    const V3 c = vadd(v3(fmul(w0, 1.0f), fmul(w0, 1.0f), fmul(w0, 1.0f)),
                      v3(fmul(hit_pt, 0.5f), fmul(hit_pt, 0.7f), fmul(hit_pt, 1.0f)));
    contribution = vmul(att, c);
    return true;
  }
  HitRec rec;
  const int mat_index = build_record(sc, sv, ray, best, rec, smem);
  const pt_material* m = reinterpret_cast<const pt_material*>(sc.materials) + mat_index;
  const int kind = m->kind;
  bool scattered_ok = true;
  Ray scattered;
  scattered.o = rec.p;
  scattered.tm = ray.tm;
  if (kind == PT_MAT_LAMBERTIAN) {  // material.hpp:18-28
    scattered.d = vadd(rec.normal, rng_unit_vec(rng));
    att = vmul(att, textured(sc, m->texture, rec));
  } else if (kind == PT_MAT_METAL) {  // material.hpp:39-48
    const V3 reflected = reflect(unit_vector(ray.d), rec.normal);
    scattered.d = vadd(reflected, vscale(m->param, rng_in_unit_ball(rng)));
    att = vmul(att, vld(m->albedo));
    scattered_ok = vdot(scattered.d, rec.normal) > 0.f;
  } else if (kind == PT_MAT_DIELECTRIC) {  // material.hpp:68-88
    att = vmul(att, vld(m->albedo));
    const float ref_idx = m->param;
    const float refraction_ratio = rec.front_face ? fdiv(1.0f, ref_idx) : ref_idx;
    const V3 unit_direction = unit_vector(ray.d);
    const float cos_theta = fminf(-vdot(unit_direction, rec.normal), 1.0f);
    const float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
    const bool cannot_refract = fmul(refraction_ratio, sin_theta) > 1.0f;
    // short-circuit: the RNG is only drawn when refraction is possible
    if (cannot_refract || reflectance(cos_theta, refraction_ratio) > rng_float(rng))
      scattered.d = reflect(unit_direction, rec.normal);
    else
      scattered.d = refract(unit_direction, rec.normal, refraction_ratio);
  } else if (kind == PT_MAT_LIGHTSOURCE) {  // material.hpp:104-108
    contribution = textured(sc, m->texture, rec);  // emitted, NOT attenuated (render.hpp:73)
    scattered_ok = false;
  } else {  // isotropic, material.hpp:119-126
    scattered.d = rng_in_unit_ball(rng);
    att = vmul(att, textured(sc, m->texture, rec));
  }
  if (!scattered_ok) return true;  // render.hpp:73 (emitted is zero for everything but lights)
  ray = scattered;
  ++bounce;
  return bounce == depth;  // render.hpp:91: out of depth -> black
}

// Queue position -> pixel of the region (false: the position falls outside the region and is skipped).
//   order_mode 1  tiles sorted by probed cost, heaviest first (longest-processing-time-first: the
//                 deep pixels of the image start at once, the cheapest ones fill the end of the frame)
//   order_mode 0  consecutive positions spread over the image by a multiplicative permutation
//   order_mode 2  the cost probe itself: every kProbeStep-th pixel of every kProbeStep-th row
// A pixel's first sample of this launch: the seed of render.hpp:130-133 and an empty sum, or -- resuming -- the state
// an earlier launch left (sum and RNG after spp_from samples).
PT_DEV void pixel_start(const RenderParams& p, int px, int py, const float* state_px, Rng& rng, V3& acc, int& sample) {
  // std::hash<size_t> is the identity; LocalPseudoRNG takes a uint32_t (rtweekend.hpp:35)
  rng.s = (uint32_t)((unsigned long long)py * (unsigned long long)p.width + (unsigned long long)px);
  acc = v3(0.f, 0.f, 0.f);
  sample = 0;
  if (p.order_mode == 2) {
    rng.s = (rng.s * 2654435761u) | 1u;  // cost probe: a throw-away stream, never the pixel's
  } else if (state_px && p.spp_from > 0) {
    const float4 s = *reinterpret_cast<const float4*>(state_px);
    acc = v3(s.x, s.y, s.z), rng.s = __float_as_uint(s.w), sample = p.spp_from;
  }
}
// The pixel is finished for this launch: render.hpp:102-105, and the state for a later pt_render_resume.
PT_DEV void pixel_finish(const RenderParams& p, float* out_px, float* state_px, V3 acc, Rng rng, float fspp) {
  const V3 fin = vdivs(acc, fspp);
  if (p.out_pixel_floats == 4)  // staged: one aligned 16-byte store; launch_resolve_fb packs the rows afterwards
    *reinterpret_cast<float4*>(out_px) = make_float4(fin.x, fin.y, fin.z, 0.f);
  else
    out_px[0] = fin.x, out_px[1] = fin.y, out_px[2] = fin.z;
  if (state_px) *reinterpret_cast<float4*>(state_px) = make_float4(acc.x, acc.y, acc.z, __uint_as_float(rng.s));
}
PT_DEV bool queue_pixel(const RenderParams& p, unsigned long long pos, int& px, int& py, float*& out_px, float*& state_px) {
  unsigned long long k, xx;
  if (p.order_mode == 1) {
    const int tile = p.tile_order[pos / (unsigned long long)(kTile * kTile)];
    const int i = (int)(pos % (unsigned long long)(kTile * kTile));
    xx = (unsigned long long)((tile % p.tiles_x) * kTile + (i % kTile));
    k = (unsigned long long)((tile / p.tiles_x) * kTile + (i / kTile));
    if (xx >= (unsigned long long)p.region.w || k >= (unsigned long long)p.region.h) return false;
  } else if (p.order_mode == 2) {
    const unsigned long long pw = (unsigned long long)((p.region.w + kProbeStep - 1) / kProbeStep);
    xx = (pos % pw) * kProbeStep, k = (pos / pw) * kProbeStep;
  } else {
    const unsigned long long n_pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
    const unsigned long long i = (pos * p.scramble) % n_pixels;
    k = i / (unsigned long long)p.region.w, xx = i - k * (unsigned long long)p.region.w;
  }
  px = p.region.x0 + (int)xx;
  py = p.region.y0 + (int)k * p.region.y_stride;
  out_px = p.out + (long long)k * p.out_row_pitch + (long long)p.out_pixel_floats * (long long)xx;
  state_px = p.state ? p.state + 4ll * ((long long)k * p.state_row_pitch + (long long)xx) : nullptr;
  return true;
}

}  // namespace
}  // namespace ptb
#endif
