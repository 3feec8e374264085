// Found as "render.hpp" by scene scripts written for triSYCL/path_tracer: the scene vocabulary and
// the render<W,H,S>() entry point, implemented on top of libptb200.so.
#pragma once
#include "pt/scene.hpp"
#include "pt/render.hpp"
