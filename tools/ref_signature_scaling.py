"""The reference's calling pattern (main.rs:192-217: worker threads x 1-spp partial_render_scene + merge_tile on the calling
thread) at several worker counts: wall time per call and where it goes.
python tools/ref_signature_scaling.py [calls] [workers,workers,...] [kahan: 0|1|both]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, host
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 64
workers_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 12]
kahans = {"0": (False,), "1": (True,), "both": (False, True)}[sys.argv[3] if len(sys.argv) > 3 else "both"]
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H = 1920, 1080
for kahan in kahans:
    for workers in workers_list:
        host.render_like_main(hs, W, H, 8 * workers, workers, kahan_state=kahan)   # warm-up: scratch blocks, pinned pool
        colour, weight, st = host.render_like_main(hs, W, H, calls, workers, kahan_state=kahan)
        assert np.all(weight == calls)
        print("%s workers %2d: %.2f ms/call  %.0f Mrays/s | merge %.2f ms/call, worker wall %.2f ms/call (/%d = %.2f), device events %.2f ms/call, D2H %.0f MB/call, %d merge passes" % (
            "five arrays " if kahan else "colour+weight", workers, 1e3 * st["wall_s"] / calls, st["rays"] / st["wall_s"] / 1e6,
            1e3 * st["merge_s"] / calls, 1e3 * st["call_s"] / calls, workers, 1e3 * st["call_s"] / calls / workers,
            st["device_ms"] / calls, st["bytes_to_host"] / calls / 1e6, st["merge_passes"]), flush=True)
