// Found as "sycl.hpp" by scene scripts written for triSYCL/path_tracer: forwards to the host facade.
#pragma once
#include "pt/sycl_facade.hpp"
