#!/usr/bin/env python
"""bench.py -- Mrays/s and spp/s of the path-traced "bunny" frame (BASELINE.json metric).

Workload (config C3 of SURVEY.md 8d): src/main.rs scene (plane + 3 spheres + bunny BVH, all
Lambertian), 1920x1080, SimpleRandomIntegrator, recursion limit 8, seed 1.  One "step" = one pass of
the hot path over one batch: --spp samples per pixel of the whole frame on EACH GPU (weak scaling:
GPU g of G renders samples g, g+G, g+2G, ...), followed, for G > 1, by the NCCL reduce of
(sum XYZ, sum weight) into rank 0's AccumulationBuffer.

    python bench.py --gpus 1 --steps K --warmup W                     (one JSON line)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     the reference's algorithm on the host cores (CPU oracle)

The Stanford bunny OBJ is a Git-LFS pointer in the reference snapshot; unless $VANRIJN_BUNNY_OBJ points at
the real file the mesh is the deterministic 81 920-triangle proxy (vanrijn_b200/scenes.py), and
`config.mesh` says so.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oraclelib: used ONLY by the cpu_baseline / reference legs

WIDTH, HEIGHT, MAX_DEPTH, SEED = 1920, 1080, 8, 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=64, help="samples per pixel per step per GPU (one wavefront batch up to 64 at 1080p)")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--filter", default="f32", choices=["f32", "f64", "f32x4", "q16"],
                    help="conservative BVH box filter: 2-wide f32, 2-wide f64, or 4-wide f32 nodes (results identical)")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64 = the reference's type (parity path, the headline); f32 = VRJ_PRECISION_F32_FAST, reported separately")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_scene(spec):
    import oraclelib as O
    return O, O.OracleScene(spec)


def cpu_sample(orc_mod, orc, width, height, spp, traverse, threads=0):
    """One bounded sample of the workload on the host cores: the full frame at reduced resolution."""
    t0 = time.perf_counter()
    r = orc.render((0, width, 0, height), height, width, spp=spp, max_depth=MAX_DEPTH, seed=SEED, traverse=traverse,
                   threads=threads)
    dt = time.perf_counter() - t0
    return r["stats"], dt


def run_reference(args, spec, mesh_name):
    """--impl reference: the reference's own algorithm for the path (reference-order BVH traversal, f64) on
    the host cores.  The Rust crate cannot be built in this image (no cargo/rustc), so this is the C++
    oracle port; it omits the Rust version's per-call heap allocations, i.e. it is a favourable stand-in."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, orc = oracle_scene(spec)
    cores = os.cpu_count() or 1
    sw, sh, sspp = args.width, args.height, 4  # one step = 4 spp of the full frame (~15 M ray queries, ~1.3 s on 16 cores)
    for _ in range(max(1, args.warmup)):
        cpu_sample(O, orc, sw, sh, sspp, O.TRAVERSE_REFERENCE)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        st, dt = cpu_sample(O, orc, sw, sh, sspp, O.TRAVERSE_REFERENCE)
        rays += st.rays
        secs += dt
    value = rays / secs / 1e6
    sample = "%dx%d full frame, %d spp per step, depth %d, reference-order traversal" % (sw, sh, sspp, MAX_DEPTH)
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "spp_per_s": (args.steps * sspp * sw * sh / (args.width * args.height)) / secs,
            "config": {"workload": "C3 main.rs scene path trace %dx%d depth %d" % (args.width, args.height, MAX_DEPTH),
                       "mesh": mesh_name, "integrator": "SimpleRandom", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    from vanrijn_b200 import scenes
    spec = scenes.scene_main(subdivisions=6, obj=True)
    mesh_name = scenes.bunny_obj_path()[1]
    if args.impl == "reference":
        return run_reference(args, spec, mesh_name)

    import numpy as np
    import torch
    import torch.distributed as dist
    import vanrijn_b200 as V
    from vanrijn_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or capi.cuda().vrj_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the render loop has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W, H, spp = args.width, args.height, args.spp
    npix = W * H
    bvh_filter = {"f64": capi.FILTER_F64, "f32x4": capi.FILTER_F32X4, "q16": capi.FILTER_Q16}.get(args.filter, capi.FILTER_F32)
    precision = capi.PRECISION_F32_FAST if args.precision == "f32" else capi.PRECISION_F64
    hs = V.build_scene(spec)
    hs.device_scene(local)
    scene_bytes = hs.device_bytes(local)
    tile = (0, W, 0, H)
    # rank-local accumulation state in HBM: (sum XYZ, weight); these are what NCCL reduces
    acc_sum = torch.zeros(npix * 3, dtype=torch.float64, device=dev)
    acc_w = torch.zeros(npix, dtype=torch.float64, device=dev)
    red_sum = torch.zeros_like(acc_sum)
    red_w = torch.zeros_like(acc_w)

    def step(k, accumulate=True):
        # sharded by sample index: this rank renders samples offset, offset+world, ... (vanrijn_b200/sharding.py)
        offset, stride = sharding.shard_samples(rank, world, k, spp)
        st = hs.render_device(tile, H, W, acc_sum.data_ptr(), acc_w.data_ptr(), device=local, accumulate=accumulate,
                              spp=spp, max_depth=MAX_DEPTH, seed=SEED, sample_offset=offset, sample_stride=stride,
                              bvh_filter=bvh_filter, precision=precision)
        if world > 1:
            # the one exchange step: combine the per-GPU accumulation buffers into rank 0's (NCCL over NVLink)
            red_sum.copy_(acc_sum)
            red_w.copy_(acc_w)
            sharding.reduce_accumulation(red_sum, red_w, dst=0)
        return st

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k, accumulate=(k > 0))
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {"rays": 0, "launches": 0, "device_ms": 0.0, "trace_ms": 0.0, "trace_launches": 0, "staged": 0,
           "shade_ms": 0.0, "resolve_ms": 0.0, "primary": 0, "bounce": 0, "shade_launches": 0}
    fence()
    ev0.record()
    t0 = time.perf_counter()
    for k in range(args.steps):
        st = step(args.warmup + k)
        agg["rays"] += st.rays
        agg["launches"] += int(st.kernel_launches) + (4 if world > 1 else 0)
        agg["device_ms"] += st.device_ms
        agg["trace_ms"] += st.primary_ms + st.bounce_ms
        agg["trace_launches"] += int(st.primary_launches + st.bounce_launches)
        agg["staged"] += int(st.staged_rays)
        agg["shade_ms"] += st.shade_ms
        agg["resolve_ms"] += st.resolve_ms
        agg["primary"] += int(st.primary_rays)
        agg["bounce"] += int(st.bounce_rays)
        agg["shade_launches"] += int(st.shade_launches)
    ev1.record()
    fence()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end: the public call with HOST buffers; scene upload (H2D) and the AccumulationBuffer
    # arrays (D2H) inside the timed region, every step
    import ctypes as C
    # the scene as the host holds it BEFORE any device work: primitives + the mesh's triangles in file order; the BVH is
    # built on the device inside vrj_scene_create (VrjBvh.n_nodes == 0), every step
    hs_e2e = V.build_scene(spec, device_builder="upload")
    desc = hs_e2e.desc()
    e2e_rays, e2e_secs, e2e_create, upload_bytes = 0, 0.0, 0.0, 0
    out_bytes = npix * 11 * 8
    # the caller's AccumulationBuffer arrays, page-locked (vrj_alloc_host) as the e2e contract asks
    pinned = {}
    for name, per in (("colour", 3), ("colour_sum", 3), ("colour_bias", 3), ("weight", 1), ("weight_bias", 1)):
        ptr = capi.cuda().vrj_alloc_host(npix * per * 8)
        if not ptr:
            raise SystemExit("vrj_alloc_host failed")
        pinned[name] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (npix * per,))
    fence()
    for k in range(max(2, min(args.steps, 3)) + 1):
        t1 = time.perf_counter()
        h = C.c_void_p()
        capi.check(capi.cuda().vrj_scene_create(C.byref(desc), local, C.byref(h)))   # H2D: the flattened scene
        t2 = time.perf_counter()
        upload_bytes = int(capi.cuda().vrj_scene_upload_bytes(h))
        hs._dev["e2e"] = h
        r = hs.render(tile, H, W, device="e2e", buffers=pinned, spp=spp, max_depth=MAX_DEPTH, seed=SEED,
                      sample_offset=sharding.shard_samples(rank, world, k, spp)[0], sample_stride=world,
                      bvh_filter=bvh_filter, precision=precision)                                              # D2H: the five arrays
        capi.cuda().vrj_scene_destroy(h)
        del hs._dev["e2e"]
        dt = time.perf_counter() - t1
        if k > 0:  # the first iteration warms allocator and staging paths
            e2e_rays += r["stats"].rays
            e2e_secs += dt
            e2e_create += t2 - t1
        e2e_steps = k
    fence()

    # ---- max over ranks / totals over ranks
    t = torch.tensor([wall, e2e_secs, agg["device_ms"]], dtype=torch.float64, device=dev)
    c = torch.tensor([agg["rays"], e2e_rays, agg["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    wall_max, e2e_max, devms_max = [float(x) for x in t.tolist()]
    rays_all, e2e_rays_all, launches_all = [float(x) for x in c.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0: roofline of the dominant kernel (k_bounce) and the CPU baseline
    peak, peak_src = peaks()
    O, orc = oracle_scene(spec)
    sw, sh = W // 8, H // 8
    st_ord, _ = cpu_sample(O, orc, sw, sh, 1, O.TRAVERSE_ORDERED)   # V, T of the ordered + pruned walk
    V_ = st_ord.node_visits / st_ord.rays
    T_ = st_ord.tri_tests / st_ord.rays
    bytes_per_ray = 32.0 * V_ + 48.0 * T_ + 144.0
    # dominant kernel: k_trace (BVH traversal).  Its algorithmic bytes per launch are the node and triangle
    # records of ALL ray queries of that level (V and T are per-query means, zeros included) plus the ray read
    # and hit write of the rays it was handed; the rest of the 144 B/ray state allowance moves in k_shade.
    trace_s = agg["trace_ms"] / 1e3
    trace_bytes = (32.0 * V_ + 48.0 * T_) * agg["rays"] + 64.0 * agg["staged"]
    achieved = trace_bytes / trace_s / 1e9 if trace_s > 0 else 0.0
    step_achieved = bytes_per_ray * agg["rays"] / (agg["device_ms"] / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_trace_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_ray": bytes_per_ray, "V_nodes_per_ray": V_, "T_tris_per_ray": T_,
                "algorithmic_bytes_per_launch": trace_bytes / max(1, agg["trace_launches"]),
                "avg_launch_ms": agg["trace_ms"] / max(1, agg["trace_launches"]),
                "kernel_share_of_step": agg["trace_ms"] / max(1e-9, agg["device_ms"]),
                "whole_step": {"achieved": step_achieved, "frac": step_achieved / peak,
                               "what": "bytes_per_ray x all ray queries / CUDA-event time of all kernels of the step"},
                "note": "the scene (BVH + triangles) is L2-resident by design, so the traversal kernel can exceed the HBM "
                        "figure; the HBM roofline is SURVEY 8d's conservative yard-stick",
                # the other half of the step: k_raygen + k_shade stream the path queues through HBM.  Algorithmic bytes: a
                # primary path reads its 16-byte hit record and writes a 16-byte result; every bounce ray is written once
                # (96 B state + 16 B stage-1 hit + 4 B list entry) and read once (96 B + 16 B) by the next level
                "k_shade": (lambda b, t: {"achieved": b / t / 1e9 if t > 0 else 0.0, "frac": (b / t / 1e9 if t > 0 else 0.0) / peak,
                                          "unit": "GB/s", "algorithmic_bytes_per_step": b / max(1, args.steps),
                                          "share_of_step": agg["shade_ms"] / max(1e-9, agg["device_ms"]),
                                          "what": "k_raygen + k_shade: 32 B per primary path + 228 B per bounce ray / their CUDA-event time"})(
                    32.0 * agg["primary"] + 228.0 * agg["bounce"], agg["shade_ms"] / 1e3)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cw, ch, cspp = W, H, min(spp, 32)  # the full frame at the step's spp: the same work as one GPU step, ~10 s of CPU
        cpu_sample(O, orc, W // 8, H // 8, 1, O.TRAVERSE_REFERENCE)  # warm-up
        st_ref, dt = cpu_sample(O, orc, cw, ch, cspp, O.TRAVERSE_REFERENCE)
        cpu = {"value": st_ref.rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": "%dx%d full frame x %d spp, depth %d, reference-order traversal, %.1f s" % (cw, ch, cspp, MAX_DEPTH, dt)}

    value = rays_all / wall_max / 1e6
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "f32" else "f64", "data": "synthetic",
            "spp_per_s": args.steps * spp * world / wall_max,
            "device_ms_per_step": devms_max / args.steps,
            "config": {"workload": "C3 main.rs scene path trace %dx%d depth %d" % (W, H, MAX_DEPTH), "mesh": mesh_name,
                       "integrator": "SimpleRandom", "spp_per_step_per_gpu": spp, "bvh_filter": args.filter,
                       "precision": "f32-fast (no parity claim)" if args.precision == "f32" else "f64 (the reference's type)",
                       "sharding": "sample index mod n_gpus; NCCL reduce of (sumXYZ, weight) to rank 0 each step",
                       "l2": "inputs larger than L2: %.1f GB of path state per step; the %.0f MB scene is L2-resident by design"
                             % (min(spp, (1 << 27) // npix) * npix * 240 / 1e9, scene_bytes / 1e6)},
            "e2e": {"value": e2e_rays_all / e2e_max / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(upload_bytes),
                    "d2h_bytes_per_step": int(out_bytes), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_max / e2e_steps, "scene_upload_ms_per_step": 1e3 * e2e_create / e2e_steps,
                    "what": "per step: vrj_scene_create from the host's triangles (BVH built on the device) + vrj_render_tile "
                            "into page-locked host AccumulationBuffer arrays + vrj_scene_destroy"},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernel_ms": {"k_trace": agg["trace_ms"], "k_raygen+k_shade": agg["shade_ms"], "k_resolve": agg["resolve_ms"]}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
