/* TEST INFRASTRUCTURE ONLY -- interface of the plain-C oracle (oracle/pt_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product never does. */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>
#include "pt_abi.h"
#ifdef __cplusplus
extern "C" {
#endif

/* All fields uint64_t (summed field-wise across threads). */
typedef struct pt_oracle_counters {
  uint64_t paths;               /* camera samples (render.hpp:95-101) */
  uint64_t scans;               /* hit_world calls (render.hpp:60) */
  uint64_t tests[5];            /* top-level object tests by PT_HIT_* kind */
  uint64_t moving_sphere_tests; /* subset of tests[PT_HIT_SPHERE] with time0 != time1 */
  uint64_t accepts[5];          /* running-closest replacements by kind (render.hpp:44-47) */
  uint64_t scatters[5];         /* successful scatters by PT_MAT_* kind */
  uint64_t sky;                 /* paths ended on the background (render.hpp:83-87) */
  uint64_t absorbed;            /* scatter() returned false (render.hpp:73) */
  uint64_t exhausted;           /* depth ran out (render.hpp:91) */
  uint64_t draws;               /* xorshift32 draws */
} pt_oracle_counters;

int pt_oracle_render_region(int width, int height, int spp, int depth, const pt_camera* cam,
                            const pt_scene* scene, const pt_region* region, float* out,
                            int64_t out_row_pitch, int dynamic_schedule, int nthreads,
                            pt_oracle_counters* counters);
/* The reference's USE_SINGLE_TASK mode (render.hpp:113-122): one RNG stream, default seed, x-major pixel order. */
int pt_oracle_render_single_task(int width, int height, int spp, int depth, const pt_camera* cam,
                                 const pt_scene* scene, float* fb, pt_oracle_counters* counters);
int pt_oracle_max_threads(void);
void pt_oracle_kat_xorshift(uint32_t seed, int n, uint32_t* out);
void pt_oracle_kat_float(uint32_t seed, int n, float* out);
void pt_oracle_kat_vec(uint32_t seed, int kind, int n, float* out);
void pt_oracle_kat_get_ray(const pt_camera* cam, uint32_t seed, int n, const float* st, float* out);
int pt_oracle_hit_world_batch(const pt_scene* scene, int n, const float* rays7, const uint32_t* seeds, float* out_t,
                              int32_t* out_index, uint32_t* out_rng);
int pt_oracle_kat_hit_scatter(const pt_scene* scene, const float* ray7, uint32_t seed, float* out);
#ifdef __cplusplus
}
#endif
#endif
