// pt/render.hpp -- the drop-in for the reference's entry point
//
//     template <int width, int height, int samples>
//     void render(sycl::queue&, sycl::buffer<color, 2>& frame_buf,
//                 std::vector<hittable_t>& hittables, camera& cam);      (reference render.hpp:141-143)
//
// Same signature, same blocking behaviour, same framebuffer layout (frame_buf[y][x], row 0 at the
// bottom, render.hpp:105), depth 50 like the reference's internal constexpr (render.hpp:144).  The
// work is done by pt_render() of libptb200.so on the GPU(s); there is no CPU path, so a missing GPU
// or a CUDA failure throws std::runtime_error with pt_last_error()'s text.
//
// Environment knobs (host side only): PT_NUM_GPUS=n renders on n GPUs of this box (rows
// interleaved, peer stores over NVLink); PT_DUMP_SCENE=path writes the flattened scene as a
// PTSCENE1 file before rendering.
#ifndef PT_RENDER_HPP
#define PT_RENDER_HPP

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "pt/flatten.hpp"
#include "pt/scene.hpp"
#include "pt_abi.h"

namespace buildparams {
#ifdef OUTPUT_WIDTH
constexpr int output_width = OUTPUT_WIDTH;
#else
constexpr int output_width = 800;
#endif
#ifdef OUTPUT_HEIGHT
constexpr int output_height = OUTPUT_HEIGHT;
#else
constexpr int output_height = 480;
#endif
#ifdef USE_SINGLE_TASK  // build_parameters.hpp:5-9 (CMake option USE_SINGLE_TASK)
constexpr bool use_single_task = true;
#else
constexpr bool use_single_task = false;
#endif
}  // namespace buildparams

namespace pt {
inline void render_runtime(int width, int height, int samples, int depth, sycl::buffer<color, 2>& frame_buf,
                           const std::vector<hittable_t>& hittables, const camera& cam) {
  const ptscene::owned_scene flat = flatten(hittables);
  const pt_scene view = flat.view();
  const pt_camera abi_cam = to_abi(cam);
  if (const char* dump = std::getenv("PT_DUMP_SCENE"))
    ptscene::save(dump, flat, abi_cam, ptscene::file_meta { width, height, samples, depth });
  if (const char* n = std::getenv("PT_NUM_GPUS"))
    if (pt_set_num_gpus(std::atoi(n)) != PT_OK) throw std::runtime_error(std::string("pt_set_num_gpus: ") + pt_last_error());
  static_assert(sizeof(color) == 3 * sizeof(float));
  if constexpr (buildparams::use_single_task) {  // render.hpp:113-122: one generator for the whole image, x-major
    if (pt_render_single_task(width, height, samples, depth, &abi_cam, &view, reinterpret_cast<float*>(frame_buf.data())) != PT_OK)
      throw std::runtime_error(std::string("pt_render_single_task: ") + pt_last_error());
    return;
  }
  if (pt_render(width, height, samples, depth, &abi_cam, &view, reinterpret_cast<float*>(frame_buf.data())) != PT_OK)
    throw std::runtime_error(std::string("pt_render: ") + pt_last_error());
}
}  // namespace pt

template <int width, int height, int samples>
void render(sycl::queue&, sycl::buffer<color, 2>& frame_buf, std::vector<hittable_t>& hittables, camera& cam) {
  constexpr int depth = 50;
  pt::render_runtime(width, height, samples, depth, frame_buf, hittables, cam);
}

#endif
