"""ctypes binding of the CPU oracle (oracle/libvanrijn_oracle.so).

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

MAT_LAMBERTIAN, MAT_PHONG, MAT_REFLECTIVE, MAT_DIELECTRIC = 0, 1, 2, 3
SIMPLE_RANDOM, WHITTED = 0, 1
TRAVERSE_REFERENCE, TRAVERSE_ORDERED = 0, 1

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class TraceCounters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("hits", C.c_uint64)]


class Light(C.Structure):
    _fields_ = [("direction", C.c_double * 3), ("spectrum", C.c_int32), ("pad", C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [("spp", C.c_uint32), ("max_depth", C.c_uint32), ("sample_offset", C.c_uint64), ("seed", C.c_uint64),
                ("integrator", C.c_uint32), ("traverse", C.c_uint32), ("bias", C.c_double),
                ("lights", C.POINTER(Light)), ("n_lights", C.c_uint32), ("ambient_spectrum", C.c_int32),
                ("threads", C.c_uint32), ("pad", C.c_uint32)]


class RenderStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primary_rays", "bounce_rays", "shadow_rays", "node_visits", "tri_tests",
                                          "paths_missed", "paths_escaped", "paths_depth_limited")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    @property
    def rays(self):
        return int(self.primary_rays + self.bounce_rays + self.shadow_rays)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "libvanrijn_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    L = C.CDLL(path)
    L.orc_scene_new.restype = C.c_void_p
    L.orc_scene_new.argtypes = [C.c_double] * 3
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_add_spectrum.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, dp]
    L.orc_add_spectrum_rgb.argtypes = [C.c_void_p] + [C.c_double] * 3
    L.orc_add_spectrum_grey.argtypes = [C.c_void_p, C.c_double]
    L.orc_add_spectrum_diamond.argtypes = [C.c_void_p]
    L.orc_add_material.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 3
    L.orc_begin_list.argtypes = [C.c_void_p]
    L.orc_list_add_sphere.argtypes = [C.c_void_p] + [C.c_double] * 4 + [C.c_int]
    L.orc_list_add_plane.argtypes = [C.c_void_p] + [C.c_double] * 4 + [C.c_int]
    L.orc_list_add_triangle.argtypes = [C.c_void_p, dp, dp, C.c_int]
    L.orc_add_bvh.argtypes = [C.c_void_p, C.c_int64, dp, dp, C.c_int]
    L.orc_add_bvh_obj.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.orc_bvh_triangle_count.restype = C.c_int64
    L.orc_bvh_triangle_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_bvh_depth.argtypes = [C.c_void_p, C.c_int]
    L.orc_bvh_leaf_order.restype = C.c_int64
    L.orc_bvh_leaf_order.argtypes = [C.c_void_p, C.c_int, ip]
    L.orc_load_obj.restype = C.c_int64
    L.orc_load_obj.argtypes = [C.c_char_p, C.POINTER(dp), C.POINTER(dp)]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_trace_rays.argtypes = [C.c_void_p, C.c_int64, dp, dp, C.c_int, ip, ip, dp, C.POINTER(TraceCounters)]
    L.orc_trace_rays_edge_distance.argtypes = [C.c_void_p, C.c_int64, dp, dp, dp]
    L.orc_render_tile.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.POINTER(RenderParams),
                                  dp, dp, dp, dp, dp, dp, C.POINTER(RenderStats)]
    for name in ("orc_triangle_intersect",):
        getattr(L, name).argtypes = [dp, dp, dp, dp, dp]
    L.orc_sphere_intersect.argtypes = [dp, C.c_double, dp, dp, dp]
    L.orc_plane_intersect.argtypes = [dp, C.c_double, dp, dp, dp]
    L.orc_aabb_intersect.argtypes = [dp, dp, dp, dp]
    L.orc_triangle_helpers.argtypes = [dp, C.POINTER(C.c_int), dp]
    L.orc_spectrum_intensity.restype = C.c_double
    L.orc_spectrum_intensity.argtypes = [C.c_double, C.c_double, C.c_int, dp, C.c_double]
    L.orc_rgb_to_spectrum.argtypes = [C.c_double] * 3 + [dp]
    L.orc_cmf_xyz.argtypes = [C.c_double, dp]
    L.orc_xyz_to_linear_rgb.argtypes = [dp, dp]
    L.orc_linear_rgb_to_xyz.argtypes = [dp, dp]
    L.orc_srgb_gamma.restype = C.c_double
    L.orc_srgb_gamma.argtypes = [C.c_double]
    L.orc_tone_map.argtypes = [C.c_int, dp, C.c_int64, C.c_void_p]
    L.orc_accum_update.argtypes = [dp, C.c_double, C.c_double, C.c_double]
    L.orc_accum_blend.argtypes = [dp, C.c_double, dp, C.c_double, dp]
    L.orc_camera_ray.argtypes = [C.c_uint64, C.c_uint64, dp, C.c_uint64, C.c_uint64, C.c_double, C.c_double, dp, dp]
    L.orc_mat3_inverse.argtypes = [dp, dp]
    L.orc_mat3_determinant.restype = C.c_double
    L.orc_mat3_determinant.argtypes = [dp]
    L.orc_largest_dimension.argtypes = [dp, dp]
    L.orc_tile_iterator.restype = C.c_int64
    L.orc_tile_iterator.argtypes = [C.c_uint64] * 3 + [C.POINTER(C.c_uint64), C.c_int64]
    L.orc_material_sample.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64,
                                      C.c_uint32, dp, dp, C.POINTER(C.c_uint32)]
    L.orc_material_bsdf.restype = C.c_double
    L.orc_material_bsdf.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_double, C.c_double]
    L.orc_sky.restype = C.c_double
    L.orc_sky.argtypes = [dp, C.c_double]
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    for name, rt in (("orc_rng_f64", C.c_double), ("orc_rng_open01", C.c_double), ("orc_rng_bool", C.c_int)):
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
    _LIB = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def vec(*xs):
    return np.array(xs, dtype=np.float64)


def hit16(fn, *args):
    out = np.zeros(16)
    keep = []
    cargs = []
    for a in args:
        if isinstance(a, (float, int)):
            cargs.append(float(a))
        else:
            arr, p = _d(a)
            keep.append(arr)
            cargs.append(p)
    ok = fn(*cargs, out.ctypes.data_as(dp))
    if not ok:
        return None
    return {"distance": out[0], "location": out[1:4], "normal": out[4:7], "tangent": out[7:10],
            "cotangent": out[10:13], "retro": out[13:16]}


def tone_map(colour, source=0):
    c = np.ascontiguousarray(colour, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((c.shape[0], 3), np.uint8)
    lib().orc_tone_map(source, c.ctypes.data_as(dp), c.shape[0], out.ctypes.data)
    return out


class OracleScene:
    """Builds the oracle's copy of a scene from a scenes.SceneSpec."""

    def __init__(self, spec):
        L = lib()
        self.L = L
        self.h = C.c_void_p(L.orc_scene_new(*[float(x) for x in spec.camera]))
        self.spec = spec
        self.spectrum_ids = []
        for sp in spec.spectra:
            kind = sp[0]
            if kind == "rgb":
                self.spectrum_ids.append(L.orc_add_spectrum_rgb(self.h, *[float(x) for x in sp[1]]))
            elif kind == "grey":
                self.spectrum_ids.append(L.orc_add_spectrum_grey(self.h, float(sp[1])))
            elif kind == "diamond":
                self.spectrum_ids.append(L.orc_add_spectrum_diamond(self.h))
            else:
                arr, p = _d(sp[3])
                self.spectrum_ids.append(L.orc_add_spectrum(self.h, float(sp[1]), float(sp[2]), len(arr), p))
        for m in spec.materials:
            L.orc_add_material(self.h, m.kind, self.spectrum_ids[m.spectrum], m.p0, m.p1, m.p2)
        for obj in spec.objects:
            if obj[0] == "list":
                L.orc_begin_list(self.h)
                for prim in obj[1]:
                    if prim[0] == "sphere":
                        L.orc_list_add_sphere(self.h, *[float(x) for x in prim[1]], float(prim[2]), prim[3])
                    elif prim[0] == "plane":
                        L.orc_list_add_plane(self.h, *[float(x) for x in prim[1]], float(prim[2]), prim[3])
                    else:
                        v, pv = _d(prim[1])
                        n, pn = _d(prim[2])
                        L.orc_list_add_triangle(self.h, pv, pn, prim[3])
            elif obj[0] == "mesh":
                v, pv = _d(obj[1])
                n, pn = _d(obj[2])
                L.orc_add_bvh(self.h, len(v) // 9 if v.ndim == 1 else v.shape[0], pv, pn, obj[3])
            elif obj[0] == "obj":
                r = L.orc_add_bvh_obj(self.h, obj[1].encode(), obj[2])
                assert r >= 0, "cannot load " + obj[1]

    def __del__(self):
        try:
            self.L.orc_scene_free(self.h)
        except Exception:
            pass

    def trace(self, origins, dirs, mode=TRAVERSE_REFERENCE):
        o, po = _d(origins)
        d, pd = _d(dirs)
        n = o.size // 3
        obj = np.empty(n, np.int32)
        prim = np.empty(n, np.int32)
        t = np.empty(n, np.float64)
        cnt = TraceCounters()
        self.L.orc_trace_rays(self.h, n, po, pd, mode, obj.ctypes.data_as(ip), prim.ctypes.data_as(ip),
                              t.ctypes.data_as(dp), C.byref(cnt))
        return obj, prim, t, cnt

    def edge_distance(self, origins, dirs):
        o, po = _d(origins)
        d, pd = _d(dirs)
        n = o.size // 3
        out = np.empty(n)
        self.L.orc_trace_rays_edge_distance(self.h, n, po, pd, out.ctypes.data_as(dp))
        return out

    def render(self, tile, height, width, spp=1, max_depth=128, sample_offset=0, seed=1, integrator=SIMPLE_RANDOM,
               traverse=TRAVERSE_REFERENCE, bias=1e-7, lights=(), ambient=-1, threads=0, want_photons=False):
        sc, ec, sr, er = tile
        npix = (ec - sc) * (er - sr)
        t4 = (C.c_uint64 * 4)(sc, ec, sr, er)
        larr = (Light * max(1, len(lights)))()
        for i, (d, s) in enumerate(lights):
            larr[i].direction[:] = [float(x) for x in d]
            larr[i].spectrum = self.spectrum_ids[s]
        p = RenderParams(spp=spp, max_depth=max_depth, sample_offset=sample_offset, seed=seed, integrator=integrator,
                         traverse=traverse, bias=bias, lights=larr, n_lights=len(lights),
                         ambient_spectrum=(self.spectrum_ids[ambient] if ambient >= 0 else -1), threads=threads)
        out = {k: np.zeros(npix * 3) for k in ("colour_sum", "colour_bias", "colour")}
        out["weight"] = np.zeros(npix)
        out["weight_bias"] = np.zeros(npix)
        photons = np.zeros(spp * npix * 2) if want_photons else None
        stats = RenderStats()
        self.L.orc_render_tile(self.h, t4, height, width, C.byref(p), out["colour_sum"].ctypes.data_as(dp),
                               out["colour_bias"].ctypes.data_as(dp), out["weight"].ctypes.data_as(dp),
                               out["weight_bias"].ctypes.data_as(dp), out["colour"].ctypes.data_as(dp),
                               photons.ctypes.data_as(dp) if want_photons else None, C.byref(stats))
        out["photons"] = photons.reshape(spp, npix, 2) if want_photons else None
        out["stats"] = stats
        return out
