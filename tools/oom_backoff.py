import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
L = capi.cuda()
hs = V.build_scene(scenes.scene_main(subdivisions=4, obj=False))
W, H = 1920, 1080
ref = hs.render((0, W, 0, H), H, W, spp=8, max_depth=4, seed=1, want=("colour_sum",))
L.vrj_release_scratch()
import subprocess
free = int(subprocess.run(["nvidia-smi", "--query-gpu=memory.free", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
block_gb = max(0, free // 1024 - 3)      # leave ~3 GB: 8 spp at 1080p wants 4 GB of queues -> must back off to 4 or 2 spp per batch
blk = L.vrj_alloc_device(0, block_gb << 30)
print("free %d MiB, blocker %d GB -> %s" % (free, block_gb, "ok" if blk else "alloc failed"))
out = hs.render((0, W, 0, H), H, W, spp=8, max_depth=4, seed=1, want=("colour_sum",))
print("launches with back-off:", out["stats"].kernel_launches, "vs", ref["stats"].kernel_launches)
assert np.array_equal(out["colour_sum"], ref["colour_sum"])
L.vrj_free_device(blk)
print("oom back-off ok")
