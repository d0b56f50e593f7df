"""Single-process multi-GPU (vrj_comm_*): C3 at 1080p, `--spp` samples per call over all visible GPUs."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
n = capi.cuda().vrj_device_count()
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H = 1920, 1080
for G in [g for g in (1, 2, 4, 8) if g <= n]:
    for i in range(a.reps):
        t0 = time.perf_counter()
        r = hs.render_sharded(list(range(G)), (0, W, 0, H), H, W, spp=a.spp * G, max_depth=8, seed=1, sample_offset=i * a.spp * G)
        dt = time.perf_counter() - t0
    st = r["stats"]
    print("G=%d: %d spp per call, %.1f ms wall (render %.1f ms device max), %.0f Mrays/s end to end (host arrays), weights ok %s" % (
        G, a.spp * G, dt * 1e3, st.device_ms, st.rays / dt / 1e6, bool(np.all(r["weight"] == a.spp * G))), flush=True)
