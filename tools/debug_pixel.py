import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import scenes
spec = scenes.scene_main(subdivisions=3, obj=False, variant="mixed")
spec.objects[0][1].append(("sphere", (1.5, 0.0, 1.0), 0.8, spec.phong_rgb((0.9, 0.2, 0.2), 0.3, 0.5, 20.0)))
hs, orc = V.build_scene(spec), O.OracleScene(spec)
W, H = 80, 45
row, col, s = 35, 74, 2
tile = (col, col + 1, row, row + 1)
for D in list(range(1, 40)) + [60, 100, 127, 128, 200]:
    g = hs.render(tile, H, W, spp=1, max_depth=D, seed=5, sample_offset=s, want_photons=True)
    r = orc.render(tile, H, W, spp=1, max_depth=D, seed=5, sample_offset=s, want_photons=True)
    print(D, "gpu", g["photons"].ravel(), g["stats"].bounce_rays, g["stats"].paths_escaped, g["stats"].paths_depth_limited,
          "| orc", r["photons"].ravel(), r["stats"].bounce_rays, r["stats"].paths_escaped, r["stats"].paths_depth_limited)
