"""Small driver for ncu: the bench workload (C3, 1920x1080, depth 8), `--spp` samples, `--reps` calls."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vanrijn_b200 as V
from vanrijn_b200 import scenes, capi
ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--filter", default="f32")
ap.add_argument("--precision", default="f64", choices=["f64", "f32"], help="f32 = VRJ_PRECISION_F32_FAST (not a parity mode)")
ap.add_argument("--variant", default="lambertian")
ap.add_argument("--count", action="store_true")
a = ap.parse_args()
hs = V.build_scene(scenes.scene_main(subdivisions=6, obj=True, variant=a.variant))
f = {"f64": capi.FILTER_F64, "f32x4": capi.FILTER_F32X4, "q16": capi.FILTER_Q16}.get(a.filter, capi.FILTER_F32)
for i in range(a.reps):
    r = hs.render((0, a.width, 0, a.height), a.height, a.width, spp=a.spp, max_depth=a.depth, seed=1, sample_offset=i * a.spp,
                  bvh_filter=f, want=("colour_sum", "weight"), count_traversal=a.count,
                  precision=capi.PRECISION_F32_FAST if a.precision == "f32" else capi.PRECISION_F64)
    st = r["stats"]
    print("rep %d: %.1f Mrays/s device (%.2f ms; trace primary %.2f bounce %.2f | raygen+shade %.2f | resolve %.2f | tail %.2f), rays %d, staged %d, launches %d" % (
        i, st.rays / st.device_ms / 1e3, st.device_ms, st.primary_ms, st.bounce_ms, st.shade_ms, st.resolve_ms, st.tail_ms, st.rays,
        st.staged_rays, st.kernel_launches), flush=True)
    if a.count:
        print("   node visits/ray %.2f  triangle tests/ray %.3f  staged %.3f of rays" % (st.node_visits / st.rays, st.triangle_tests / st.rays, st.staged_rays / st.rays))
