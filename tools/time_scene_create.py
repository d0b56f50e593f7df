"""Time vrj_scene_create / vrj_scene_destroy of the bench scene: host-built tree vs tree built at upload."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
spec = scenes.scene_main(subdivisions=6, obj=True)
L = capi.cuda()
for mode in (False, "upload"):
    hs = V.build_scene(spec, device_builder=mode)
    desc = hs.desc()
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        h = C.c_void_p()
        capi.check(L.vrj_scene_create(C.byref(desc), 0, C.byref(h)))
        t1 = time.perf_counter()
        L.vrj_scene_destroy(h)
        t2 = time.perf_counter()
        ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    print("builder=%s  create ms: %s" % (mode, " ".join("%.2f" % a for a, b in ts)))
    print("              destroy ms: %s" % " ".join("%.2f" % b for a, b in ts), flush=True)
