"""The bench's e2e_ref_signature sequence on its own (4 and 8 workers, colour-only then five arrays), optionally after a
64-spp call (argument `big`) that leaves a 45 GB scratch block in the pool, as in bench.py.  Diagnostic."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, host
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H = 1920, 1080
if "big" in sys.argv:
    r = hs.render((0, W, 0, H), H, W, spp=64, max_depth=8, seed=1, want=("colour",))
    print("64-spp call: %.1f ms device" % r["stats"].device_ms, flush=True)
if "torch" in sys.argv:
    import torch
    x = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
for rep in range(2):
    for key, kahan, calls in (("colour_only", False, 192),):
        for workers in (4, 8):
            host.render_like_main(hs, W, H, 6 * workers, workers, kahan_state=kahan)
            colour, weight, st = host.render_like_main(hs, W, H, calls, workers, kahan_state=kahan)
            print(key, workers, "%.2f ms/call  %.0f Mrays/s merge %.2f passes %d device %.2f ms/call, %.2f calls per wavefront" % (1e3*st["wall_s"]/calls, st["rays"]/st["wall_s"]/1e6, 1e3*st["merge_s"]/calls, st["merge_passes"], st["device_ms"]/calls, st["wavefront_calls"]/calls), flush=True)
