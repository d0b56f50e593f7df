"""vanrijn_b200 -- B200-native (sm_100a) implementation of vanrijn's per-pixel / per-sample render
loop behind the reference's library API.

  include/vanrijn_cuda.h    the C ABI (the drop-in boundary)
  include/vanrijn.hpp       C++ host mirror of the reference API (Scene, BVH build, load_obj,
                            partial_render_scene, AccumulationBuffer, Tile...)
  vanrijn_b200/csrc         CUDA kernels + C ABI implementation, host library sources
  vanrijn_b200/{capi,host}  ctypes plumbing used by the tests and bench.py
  vanrijn_b200/scenes       the benchmark / parity scene configurations
"""
from . import capi, scenes  # noqa: F401
from .host import HostScene, build_scene  # noqa: F401

__all__ = ["capi", "scenes", "HostScene", "build_scene"]
