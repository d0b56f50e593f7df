"""The oracle against an INDEPENDENT second source: tests/golden/second_source.py restates the
reference's per-sample algorithm in plain Python straight from the Rust lines and imports nothing from oracle/; its
vectors (tests/golden/second_source.json) pin the oracle rows the reference itself holds no tests for -- integrators,
materials, Fresnel (through the dielectric), BVH build order + traversal with ties, OBJ loader, camera, the rand
conversions.  Bars: bit-exact wherever only + - * / sqrt are involved; 1e-12 relative where libm (acos / exp / pow /
sin / cos) enters.  The product's host builder and loader are checked against the same vectors.  CPU only."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import capi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def vec():
    return json.load(open(os.path.join(HERE, "golden", "second_source.json")))


def f(bits):
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def fv(bits3):
    return np.array([f(b) for b in bits3])


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float64).view(np.uint64), np.asarray(b, np.float64).view(np.uint64))


def close(a, b, rtol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-300)))


def test_rand_conversions(vec):
    L = O.lib()
    for r in vec["rng"]:
        args = (r["seed"], r["pixel"], r["sample"], r["ordinal"])
        assert same_bits(L.orc_rng_f64(*args), f(r["f64"]))
        assert same_bits(L.orc_rng_open01(*args), f(r["open01"]))
        assert int(L.orc_rng_bool(*args)) == r["boolean"]


def test_spectra_and_sky(vec):
    L = O.lib()
    for r in vec["spectrum"]:
        s = np.zeros(32)
        L.orc_rgb_to_spectrum(*[float(x) for x in r["rgb"]], s.ctypes.data_as(O.dp))
        got = L.orc_spectrum_intensity(380.0, 720.0, 32, s.ctypes.data_as(O.dp), float(r["wavelength"]))
        assert same_bits(got, f(r["intensity"])), r
    for r in vec["sky"]:
        w = np.array(r["w"], np.float64)
        assert same_bits(L.orc_sky(w.ctypes.data_as(O.dp), float(r["wavelength"])), f(r["value"])), r


def material_spec():
    spec = scenes.SceneSpec(camera=(0.0, 0.0, 0.0))
    rgb = (0.55, 0.27, 0.04)
    ids = {"lambertian": spec.lambertian_rgb(rgb, 0.1), "phong": spec.phong_rgb(rgb, 0.3, 0.5, 20.0),
           "reflective": spec.reflective_rgb(rgb, 0.05, 0.9), "dielectric": spec.dielectric_diamond()}
    spec.objects.append(("list", [("sphere", (0.0, 0.0, 5.0), 1.0, ids["lambertian"])]))
    return spec, ids


def test_material_sample_and_bsdf(vec):
    """Material::sample / bsdf of the four materials (lambertian_material.rs:27-59, phong_material.rs:16-36 + the trait's
    cosine-hemisphere sampler, reflective_material.rs:15-47, smooth_transparent_dialectric.rs:15-114 incl. fresnel and
    the diamond index table) for fixed draws, in both argument orders the integrators use."""
    spec, ids = material_spec()
    orc = O.OracleScene(spec)
    L = orc.L
    exact = {"lambertian": (True, True), "dielectric": (True, True), "reflective": (True, False), "phong": (False, False)}
    for r in vec["materials"]:
        m = ids[r["material"]]
        w_i = fv(r["w_i"])
        d, pdf, used = np.zeros(3), C.c_double(), C.c_uint32()
        L.orc_material_sample(orc.h, m, w_i.ctypes.data_as(O.dp), float(r["wavelength"]), r["seed"], r["pixel"], r["sample"],
                              r["ordinal"], d.ctypes.data_as(O.dp), C.byref(pdf), C.byref(used))
        sample_exact, bsdf_exact = exact[r["material"]]
        want_d = fv(r["direction"])
        assert (same_bits(d, want_d) if sample_exact else close(d, want_d)), (r["material"], d, want_d)
        assert (same_bits(pdf.value, f(r["pdf"])) if sample_exact else close(pdf.value, f(r["pdf"]))), r["material"]
        assert r["ordinal"] + used.value == r["ordinal_after"], r["material"]
        b1 = L.orc_material_bsdf(orc.h, m, want_d.ctypes.data_as(O.dp), w_i.ctypes.data_as(O.dp), float(r["wavelength"]), 0.75)
        b2 = L.orc_material_bsdf(orc.h, m, w_i.ctypes.data_as(O.dp), want_d.ctypes.data_as(O.dp), float(r["wavelength"]), 0.75)
        for got, want in ((b1, f(r["bsdf_sampled_retro"])), (b2, f(r["bsdf_retro_sampled"]))):
            assert (same_bits(got, want) if bsdf_exact else close(got, want)), (r["material"], got, want)


def test_obj_loader(vec, tmp_path):
    """mesh.rs:13-88 over obj 0.9: all index forms, negative indices, fan triangulation, f32 parse then widening,
    missing normals -> zeros.  Oracle loader and the product's C++ loader against the second source."""
    path = tmp_path / "seven.obj"
    path.write_text(vec["obj"]["text"])
    want_v = np.array([[f(b) for p in t["v"] for b in p] for t in vec["obj"]["triangles"]])
    want_n = np.array([[f(b) for p in t["n"] for b in p] for t in vec["obj"]["triangles"]])
    L = O.lib()
    pv, pn = O.dp(), O.dp()
    n = L.orc_load_obj(str(path).encode(), C.byref(pv), C.byref(pn))
    assert n == len(want_v) == 7
    got_v = np.ctypeslib.as_array(pv, (n, 9)).copy()
    got_n = np.ctypeslib.as_array(pn, (n, 9)).copy()
    L.orc_free(pv), L.orc_free(pn)
    assert same_bits(got_v, want_v) and same_bits(got_n, want_n)
    hv, hn = np.zeros((n, 9)), np.zeros((n, 9))
    assert capi.host().vrjh_load_obj(str(path).encode(), hv.ctypes.data_as(O.dp), hn.ctypes.data_as(O.dp), n) == n
    assert same_bits(hv, want_v) and same_bits(hn, want_n)


def mesh_spec(tris_v, tris_n, camera=(0.0, 0.0, 0.0)):
    spec = scenes.SceneSpec(camera=camera)
    m = spec.lambertian_rgb((1.0, 1.0, 0.0), 0.05)
    spec.objects.append(("mesh", np.asarray(tris_v, np.float64).reshape(-1, 9), np.asarray(tris_n, np.float64).reshape(-1, 9), m))
    return spec


def test_bvh_build_order_and_traversal_with_ties(vec):
    """bounding_volume_hierarchy.rs:38-119 on seven triangles whose centres tie on the split axis: the DFS leaf order of the
    median-split build (stable sort) and the closest hit of the unordered, unpruned traversal ("later leaf wins ties",
    :77-92).  Oracle AND the product's host builder against the second source."""
    v = np.array([[f(b) for p in t["v"] for b in p] for t in vec["obj"]["triangles"]])
    n = np.array([[f(b) for p in t["n"] for b in p] for t in vec["obj"]["triangles"]])
    spec = mesh_spec(v, n)
    orc = O.OracleScene(spec)
    order = np.zeros(7, np.int32)
    assert orc.L.orc_bvh_leaf_order(orc.h, 0, order.ctypes.data_as(O.ip)) == 7
    assert order.tolist() == vec["bvh"]["leaf_order"]
    d = V.build_scene(spec).desc()
    assert np.ctypeslib.as_array(d.tri_prim_id, (7,)).tolist() == vec["bvh"]["leaf_order"]
    o = np.array([r["origin"] for r in vec["bvh"]["rays"]], np.float64)
    dr = np.array([r["direction"] for r in vec["bvh"]["rays"]], np.float64)
    obj, prim, t, _ = orc.trace(o, dr, mode=O.TRAVERSE_REFERENCE)
    hits = 0
    for k, r in enumerate(vec["bvh"]["rays"]):
        assert prim[k] == r["prim"], (k, prim[k], r["prim"])
        if r["prim"] >= 0:
            hits += 1
            assert same_bits(t[k], f(r["distance"])), k
    assert hits >= 5
    obj2, prim2, t2, _ = orc.trace(o, dr, mode=O.TRAVERSE_ORDERED)   # the ordered + pruned walk must agree with the reference's
    assert np.array_equal(prim2, prim) and same_bits(t2, t)


@pytest.mark.parametrize("index", range(6))
def test_whole_samples(vec, index):
    """camera.rs:95-130 end to end for 24 x 16 pixels x 2 samples: camera ray (ImageSampler), Sampler::sample over a
    primitive list + a BVH, SimpleRandomIntegrator (simple_random_integrator.rs:12-65, recursion as written, sky) and
    WhittedIntegrator (whitted_integrator.rs:20-86: shadow rays, ambient term, the limit-0 bounce), all four materials.
    Wavelengths bit-identical; radiance bit-identical in the all-Lambertian scenes (only + - * / sqrt), 1e-12 with the
    libm-dependent materials (where a path may also legitimately fork on a last-ulp direction: at most 2 of 768)."""
    s = vec["samples"][index]
    spec = scenes.SceneSpec(camera=tuple(s["camera"]))
    ground = spec.lambertian_rgb((0.55, 0.27, 0.04), 0.1)
    if s["variant"] == "lambertian":
        m1, m2, m3 = spec.lambertian_rgb(scenes.NAMED["Green"], 0.1), spec.lambertian_rgb(scenes.NAMED["Blue"], 0.1), spec.lambertian_rgb(scenes.NAMED["Red"], 0.05)
        mm = spec.lambertian_rgb(scenes.NAMED["Yellow"], 0.05)
    else:
        m1, m2, m3 = spec.phong_rgb(scenes.NAMED["Green"], 0.3, 0.5, 20.0), spec.reflective_rgb(scenes.NAMED["Blue"], 0.01, 0.99), spec.dielectric_diamond()
        mm = spec.reflective_rgb(scenes.NAMED["Yellow"], 0.05, 0.9)
    spec.objects.append(("list", [("plane", (0.0, 1.0, 0.0), -2.0, ground), ("sphere", (-6.25, -0.5, 1.0), 1.0, m1),
                                  ("sphere", (-4.25, -0.5, 2.0), 1.0, m2), ("sphere", (-5.0, 1.5, 1.0), 1.0, m3)]))
    mv = np.array([[c for p in t["v"] for c in p] for t in s["mesh"]])
    mn = np.array([[c for p in t["n"] for c in p] for t in s["mesh"]])
    spec.objects.append(("mesh", mv, mn, mm))
    orc = O.OracleScene(spec)
    W, H = s["width"], s["height"]
    kw = {}
    if s["integrator"] == "whitted":
        light = spec.spectrum("grey", 1.0)
        amb = spec.spectrum("grey", 0.05)
        orc = O.OracleScene(spec)
        kw = dict(integrator=O.WHITTED, lights=[((1.0, 1.0, -1.0), light)], ambient=amb)
    r = orc.render((0, W, 0, H), H, W, spp=4, max_depth=s["max_depth"], seed=s["seed"], want_photons=True, **kw)
    ph = r["photons"]  # [sample, pixel, 2]
    rows = np.array(s["photons"], dtype=np.uint64)
    pix = (rows[:, 0] * W + rows[:, 1]).astype(np.int64)
    smp = rows[:, 2].astype(np.int64)
    want_wl = rows[:, 3].copy().view(np.float64)
    want_i = rows[:, 4].copy().view(np.float64)
    got_wl, got_i = ph[smp, pix, 0], ph[smp, pix, 1]
    # a depth-limited path keeps wavelength 0 in the reference (simple_random_integrator.rs:20-25) and in both restatements
    assert same_bits(got_wl, want_wl)
    if s["integrator"] == "whitted" or s["max_depth"] > 0:
        assert np.count_nonzero(want_i) > 100
    else:
        assert np.count_nonzero(want_i) == 0      # recursion limit 0: SimpleRandom returns Photon{0, 0} at once
    if s["variant"] == "lambertian":
        assert same_bits(got_i, want_i)
    else:
        bad = np.abs(got_i - want_i) > 1e-12 * np.maximum(np.abs(want_i), 1e-300)
        assert bad.sum() <= 2, (bad.sum(), got_i[bad][:4], want_i[bad][:4])
    assert r["stats"].bounce_rays >= int(rows[:, 5].sum())   # the oracle rendered 4 samples per pixel, the vectors hold 2 of them
