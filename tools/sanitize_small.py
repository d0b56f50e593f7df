"""Smallest run that touches every kernel family (for compute-sanitizer): device BVH build (global levels + subtrees),
scene preparation, all four box-filter walks, both integrators, the deep-recursion tail, tone map, f32-fast."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi, host
rng = np.random.default_rng(0)
tri = (rng.uniform(-1, 1, size=(3000, 1, 3)) + rng.normal(size=(3000, 3, 3)) * 0.05).astype(np.float32).astype(np.float64).reshape(-1, 9)
r = host.bvh_build(tri)
assert sorted(r["order"].tolist()) == list(range(3000))
spec = scenes.scene_main(subdivisions=3, obj=False, variant="mixed")
hs = V.build_scene(spec, device_builder="upload")
W, H = 96, 54
ref = None
for f in (capi.FILTER_F32, capi.FILTER_F64, capi.FILTER_F32X4, capi.FILTER_Q16):
    out = hs.render((0, W, 0, H), H, W, spp=2, max_depth=40, seed=1, want=("colour", "srgb8"), want_photons=True, bvh_filter=f)
    ref = ref if ref is not None else out["photons"]
    assert np.array_equal(ref, out["photons"])
hs.render((0, W, 0, H), H, W, spp=2, max_depth=8, seed=1, precision=capi.PRECISION_F32_FAST)
spec2, lights, amb = scenes.scene_direct(subdivisions=3, obj=False)
hw = V.build_scene(spec2)
hw.render((0, W, 0, H), H, W, spp=1, max_depth=1, seed=1, integrator=capi.INTEGRATOR_WHITTED, lights=lights, ambient=amb)
print("sanitize_small ok")
