"""path_tracer_b200 -- B200-native (sm_100a CUDA) drop-in for ONE hot path of
triSYCL/path_tracer: the per-pixel path-tracing loop behind render<W,H,S>()
(reference include/render.hpp:25-160).

The product is the C-ABI shared library path_tracer_b200/lib/libptb200.so
(include/pt_abi.h) plus the C++20 host headers in path_tracer_b200/include/
that keep the reference's scene API and `render<>()` signature.  This Python
package is harness-side plumbing only: ctypes bindings, flat-scene I/O and the
one-process-per-GPU launcher used by tests/ and bench.py.  There is no CPU
fallback anywhere in it: without the CUDA library or a GPU every render call
raises.
"""
from . import abi  # noqa: F401
from .scene import Scene, camera_c, make_camera  # noqa: F401

__all__ = ["abi", "Scene", "camera_c", "make_camera"]
