import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import scenes
variant = sys.argv[1] if len(sys.argv) > 1 else "mixed"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 128
spec = scenes.scene_main(subdivisions=3, obj=False, variant=variant)
spec.objects[0][1].append(("sphere", (1.5, 0.0, 1.0), 0.8, spec.phong_rgb((0.9, 0.2, 0.2), 0.3, 0.5, 20.0)))
hs, orc = V.build_scene(spec), O.OracleScene(spec)
W, H, spp = 80, 45, 3
g = hs.render((0, W, 0, H), H, W, spp=spp, max_depth=depth, seed=5, want_photons=True)
r = orc.render((0, W, 0, H), H, W, spp=spp, max_depth=depth, seed=5, want_photons=True)
print("gpu stats", g["stats"].as_dict())
print("orc stats", r["stats"].as_dict())
gp, rp = g["photons"], r["photons"]
bad = np.argwhere(gp[..., 0] != rp[..., 0])
print("n wavelength mismatches", len(bad))
for s, p in bad[:20]:
    print("sample", s, "pixel", p, "row", p // W, "col", p % W, "gpu", gp[s, p], "orc", rp[s, p])
live = (rp[..., 0] != 0) & (gp[..., 0] == rp[..., 0])
rel = np.abs(gp[..., 1][live] - rp[..., 1][live]) / np.maximum(np.abs(rp[..., 1][live]), 1e-300)
print("max rel err on matching", rel.max(), "nan gpu", np.isnan(gp).sum(), "nan orc", np.isnan(rp).sum())
