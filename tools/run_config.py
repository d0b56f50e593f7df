"""Run one of the BASELINE.json configurations on the GPU and print throughput (device time, CUDA events).
  C2: bunny, Whitted depth 0, one directional light, 1920x1080, 1 spp
  C3: main.rs scene path trace 1920x1080 depth 8 (the bench workload)
  C4: 11x11 bunny grid (9.9 M triangles) + plane, 3840x2160, depth 8
  C5: mirror + diamond spheres + reflective bunny, 1920x1080, depth 128
"""
import argparse, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["C2", "C3", "C4", "C5"])
ap.add_argument("--spp", type=int, default=None)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--filter", default="f32")
ap.add_argument("--precision", default="f64", choices=["f64", "f32"], help="f32 = VRJ_PRECISION_F32_FAST (not a parity mode)")
ap.add_argument("--device-builder", action="store_true", help="BoundingVolumeHierarchy::build on the GPU (vrj_bvh_build)")
a = ap.parse_args()
kw = dict(seed=1, precision=capi.PRECISION_F32_FAST if a.precision == "f32" else capi.PRECISION_F64, bvh_filter={"f64": capi.FILTER_F64, "f32x4": capi.FILTER_F32X4, "q16": capi.FILTER_Q16}.get(a.filter, capi.FILTER_F32))
t0 = time.time()
if a.config == "C2":
    spec, lights, amb = scenes.scene_direct(subdivisions=6, obj=True)
    W, H, spp = 1920, 1080, a.spp or 1
    kw.update(integrator=capi.INTEGRATOR_WHITTED, lights=lights, ambient=amb, max_depth=0)
elif a.config == "C3":
    spec = scenes.scene_main(subdivisions=6, obj=True)
    W, H, spp = 1920, 1080, a.spp or 16
    kw.update(max_depth=8)
elif a.config == "C4":
    spec = scenes.scene_grid(copies=11)
    W, H, spp = 3840, 2160, a.spp or 4
    kw.update(max_depth=8)
else:
    spec = scenes.scene_main(subdivisions=6, obj=True, variant="mixed")
    W, H, spp = 1920, 1080, a.spp or 16
    kw.update(max_depth=128)
hs = V.build_scene(spec, device_builder=a.device_builder)
t1 = time.time()
hs.device_scene(0)
t2 = time.time()
print("%s: host scene build %.1f s, flatten+upload %.1f s, %.2f GB on device" % (a.config, t1 - t0, t2 - t1, hs.device_bytes() / 1e9), flush=True)
for i in range(a.reps):
    r = hs.render((0, W, 0, H), H, W, spp=spp, sample_offset=i * spp, want=("colour_sum", "weight"), **kw)
    st = r["stats"]
    ok = bool(np.all(r["weight"] == spp) and np.all(np.isfinite(r["colour_sum"])))
    print(json.dumps({"config": a.config, "rep": i, "Mrays_per_s": st.rays / st.device_ms / 1e3, "spp_per_s": spp / (st.device_ms / 1e3),
                      "device_ms": st.device_ms, "rays": st.rays, "primary": int(st.primary_rays), "bounce": int(st.bounce_rays),
                      "shadow": int(st.shadow_rays), "staged": int(st.staged_rays), "launches": int(st.kernel_launches),
                      "trace_ms": st.primary_ms + st.bounce_ms, "shade_ms": st.shade_ms, "resolve_ms": st.resolve_ms,
                      "tail_ms": st.tail_ms, "tail_launches": int(st.tail_launches), "depth_limited": int(st.paths_depth_limited),
                      "weights_ok_finite": ok, "mean_Y": float(r["colour_sum"].reshape(-1, 3)[:, 1].mean() / spp)}), flush=True)
