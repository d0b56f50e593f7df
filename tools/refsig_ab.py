"""`workers` (default 8) x 192 one-sample calls (the bench's e2e_ref_signature case), `reps` times:
python tools/refsig_ab.py [reps] [workers]"""
import sys
sys.path.insert(0, ".")
import vanrijn_b200 as V
from vanrijn_b200 import scenes, host
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True))
W, H, workers, calls = 1920, 1080, (int(sys.argv[2]) if len(sys.argv) > 2 else 8), 192
host.render_like_main(hs, W, H, 6 * workers, workers, kahan_state=False)
out = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    colour, weight, st = host.render_like_main(hs, W, H, calls, workers, kahan_state=False)
    out.append("%.0f (%.2f/wf, %d passes)" % (st["rays"] / st["wall_s"] / 1e6, st["wavefront_calls"] / calls, st["merge_passes"]))
print("Mrays/s:", ", ".join(out), flush=True)
