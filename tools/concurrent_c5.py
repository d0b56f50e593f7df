"""Config C5 (mixed materials, recursion limit 128): 16 spp at 1080p as 1 x 16, 2 x 8, 4 x 4 concurrent calls -- does the
latency-bound tail of one wavefront overlap the bulk of another?  python tools/concurrent_c5.py"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True, variant="mixed"))
W, H, TOTAL = 1920, 1080, 16
def run(parts, reps=4):
    spp = TOTAL // parts
    sums = [torch.zeros(W * H * 3, dtype=torch.float64, device="cuda") for _ in range(parts)]
    ws = [torch.zeros(W * H, dtype=torch.float64, device="cuda") for _ in range(parts)]
    best = 1e9
    for rep in range(reps):
        rays = [0] * parts
        def work(i):
            st = hs.render_device((0, W, 0, H), H, W, sums[i].data_ptr(), ws[i].data_ptr(), spp=spp, max_depth=128, seed=1, sample_offset=rep * TOTAL + i * spp)
            rays[i] = st.rays
        th = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep: best = min(best, dt)
    print("%d concurrent calls of %2d spp: %.2f ms per %d spp, %.0f Mrays/s" % (parts, spp, best * 1e3, TOTAL, sum(rays) / best / 1e6), flush=True)
for g in (1, 2, 4, 8):
    run(g)
