"""Driver for ncu / timing of vrj_bvh_build: builds the tree of the C4 mesh (9.9 M triangles) or of the bench mesh."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from vanrijn_b200 import scenes, host
ap = argparse.ArgumentParser()
ap.add_argument("--mesh", default="grid", choices=["grid", "bunny"])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
if a.mesh == "grid":
    v = scenes.scene_grid(copies=11).objects[1][1]
else:
    v = scenes.mesh_arrays(*scenes.bunny_proxy(6))[0]
for i in range(a.reps):
    t0 = time.perf_counter()
    r = host.bvh_build(v)
    dt = time.perf_counter() - t0
    s = r["stats"]
    print("rep %d: %d triangles, device %.2f ms (%.1f Mtris/s), call %.1f ms incl. H2D/D2H; %d global levels, %d radix passes, %d subtrees, depth %d"
          % (i, v.shape[0], s.device_ms, v.shape[0] / s.device_ms / 1e3, 1e3 * dt, s.global_levels, s.radix_passes, s.small_subtrees, r["depth"]), flush=True)
