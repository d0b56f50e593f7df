import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H, spp = 480, 270, 64
a = hs.render((0, W, 0, H), H, W, spp=spp, max_depth=8, seed=1, want=("colour",))
b = hs.render((0, W, 0, H), H, W, spp=spp, max_depth=8, seed=1, want=("colour",), precision=capi.PRECISION_F32_FAST)
ca, cb = a["colour"].reshape(-1, 3), b["colour"].reshape(-1, 3)
print("rays f64 %d  f32 %d" % (a["stats"].rays, b["stats"].rays))
print("mean Y f64 %.6f f32 %.6f" % (ca[:, 1].mean(), cb[:, 1].mean()))
rmse = np.sqrt(((ca - cb) ** 2).mean()); print("RMSE %.3e  relative to mean %.3e" % (rmse, rmse / ca.mean()))
d = np.abs(ca - cb).max(1); print("pixels differing by > 1e-3 abs: %.3f%%" % (100 * (d > 1e-3).mean()))
