"""One coalesced wavefront of `n` one-sample 1080p calls (n threads call vrj_render_tile at the same moment behind a busy
gate) -- for `ncu --metrics gpu__time_duration.sum`: where the fixed cost of a small wavefront goes.
python tools/one_wavefront.py [n]"""
import sys, threading
sys.path.insert(0, ".")
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H = 1920, 1080
hs.device_scene(0)
for rep in range(3):
    got = [None] * n
    start = threading.Barrier(n)
    def work(i):
        start.wait()
        got[i] = hs.render((0, W, 0, H), H, W, spp=1, max_depth=128, seed=1, sample_offset=100 * rep + i, want=("colour",))
    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th: t.start()
    for t in th: t.join()
    print("rep %d: calls per wavefront %s, device ms per call %s" % (rep, [int(g["stats"].coalesced_calls) for g in got], ["%.2f" % g["stats"].device_ms for g in got]), flush=True)
