"""raytracer.c_b200 -- B200-native path tracer behind gue-ni/raytracer.c's raytracer.h surface.

The product is native: csrc/ (sm_100a CUDA kernels + the C ABI of include/rtb200.h) and
host/ (C99: render(), init_camera(), load_obj(), scene builders, CLI).  This Python
package is only the ctypes view of those libraries used by tests/ and bench.py.

The directory name contains a dot, so import it through `__graft_entry__.load_package()`
(which registers it as `raytracer_c_b200`).
"""
from . import abi, api  # noqa: F401
from .api import RtbError, Scene, load  # noqa: F401
