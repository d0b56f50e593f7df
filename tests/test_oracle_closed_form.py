"""Closed-form checks of the oracle where the reference has no tests of its own (integrators, materials,
sky): whole-path results are re-derived in numpy from independently exposed pieces and from the formulas of
simple_random_integrator.rs:12-65 / whitted_integrator.rs:20-86.  CPU only."""
import ctypes as C

import numpy as np

import oraclelib as O
from oraclelib import dp
from vanrijn_b200 import scenes

L = O.lib()


def _plane_scene(material="lambertian"):
    s = scenes.SceneSpec(camera=(0.0, 0.0, 0.0))
    if material == "lambertian":
        m = s.lambertian_rgb((0.55, 0.27, 0.04), 0.1)
    else:
        m = s.reflective_rgb((1.0, 1.0, 0.0), 0.05, 0.9)
    s.objects.append(("list", [("plane", (0.0, 1.0, 0.0), -2.0, m)]))
    return s


def _spectrum_at(rgb, wl):
    smp = np.zeros(32)
    L.orc_rgb_to_spectrum(*rgb, smp.ctypes.data_as(dp))
    return L.orc_spectrum_intensity(380.0, 720.0, 32, smp.ctypes.data_as(dp), wl)


def test_one_bounce_lambertian_plane_under_the_sky():
    """camera -> plane -> sky:  I = colour(l) * ds * ( sky(W, l) * pdf * |W.n| ) * 360, pdf = cos*sin/pi (sic)."""
    spec = _plane_scene()
    orc = O.OracleScene(spec)
    W, H, seed = 16, 16, 11
    r = orc.render((0, W, 0, H), H, W, spp=1, max_depth=128, seed=seed, want_photons=True)
    ph = r["photons"][0]
    checked = 0
    for row in range(H):
        for col in range(W):
            p = row * W + col
            ux, uy = L.orc_rng_f64(seed, p, 0, 0), L.orc_rng_f64(seed, p, 0, 1)
            o, d = np.zeros(3), np.zeros(3)
            L.orc_camera_ray(W, H, O.vec(0, 0, 0).ctypes.data_as(dp), row, col, ux, uy, o.ctypes.data_as(dp), d.ctypes.data_as(dp))
            hit = O.hit16(L.orc_plane_intersect, O.vec(0, 1, 0), -2.0, o, d)
            if hit is None:
                assert ph[p, 0] == 0.0 and ph[p, 1] == 0.0          # camera.rs:110-113: miss is black, not sky
                continue
            wl = 380.0 + 360.0 * L.orc_rng_f64(seed, p, 0, 2)
            assert ph[p, 0] == wl
            M = np.stack([hit["tangent"], hit["cotangent"], hit["normal"]])
            Minv = np.zeros(9)
            assert L.orc_mat3_inverse(M.ravel().copy().ctypes.data_as(dp), Minv.ctypes.data_as(dp)) == 1
            w_i = M @ hit["retro"]
            w_o, pdf, used = np.zeros(3), C.c_double(), C.c_uint32()
            L.orc_material_sample(orc.h, 0, w_i.ctypes.data_as(dp), wl, seed, p, 0, 4, w_o.ctypes.data_as(dp), C.byref(pdf), C.byref(used))
            assert used.value >= 2 and used.value % 2 == 0          # rejection sampling consumes pairs of draws
            assert abs(pdf.value - w_o[2] * np.sqrt(1 - w_o[2] ** 2) / np.pi) < 1e-12   # lambertian_material.rs:57
            Wd = Minv.reshape(3, 3) @ w_o
            assert Wd[1] > 0                                         # the plane's normal is +y: the bounce always escapes
            sky = L.orc_sky(Wd.ctypes.data_as(dp), wl)
            expect = _spectrum_at((0.55, 0.27, 0.04), wl) * 0.1 * (sky * pdf.value * abs(Wd @ hit["normal"])) * 360.0
            assert abs(ph[p, 1] - expect) <= 1e-12 * max(1.0, abs(expect))
            checked += 1
    assert checked > 60
    assert r["stats"].bounce_rays == checked and r["stats"].paths_escaped == checked


def test_sky_is_the_rgb_decomposition_of_y_y_1():
    """simple_random_integrator.rs:57-65: rgb = (W.y, W.y, 1): white*y + blue*(1-y) for y <= 1; 0 above 720 nm."""
    white, blue = np.zeros(32), np.zeros(32)
    L.orc_rgb_to_spectrum(1, 1, 1, white.ctypes.data_as(dp))
    L.orc_rgb_to_spectrum(0, 0, 1, blue.ctypes.data_as(dp))
    for y in (-0.7, 0.0, 0.3, 1.0):
        for wl in (380.0, 455.5, 600.0, 719.9):
            w = O.vec(0.1, y, 0.2)
            mix = y * white + (1.0 - y) * blue
            expect = L.orc_spectrum_intensity(380.0, 720.0, 32, mix.ctypes.data_as(dp), wl)
            assert abs(L.orc_sky(w.ctypes.data_as(dp), wl) - expect) < 1e-14
    assert L.orc_sky(O.vec(0, 0.5, 0).ctypes.data_as(dp), 730.0) == 0.0


def test_depth_limit_returns_wavelength_zero():
    """Two facing planes: no path can escape; with limit D every path does exactly D bounce rays and ends as
    Photon{0,0} (simple_random_integrator.rs:20-25), whose XYZ is ~0."""
    s = scenes.SceneSpec(camera=(0.0, 0.0, 0.0))
    m = s.lambertian_rgb((0.5, 0.5, 0.5), 1.0)
    s.objects.append(("list", [("plane", (0.0, 1.0, 0.0), -2.0, m), ("plane", (0.0, -1.0, 0.0), -2.0, m)]))
    orc = O.OracleScene(s)
    for D in (0, 1, 5):
        r = orc.render((0, 8, 0, 8), 8, 8, spp=1, max_depth=D, seed=2, want_photons=True)
        hit = r["stats"].primary_rays - r["stats"].paths_missed
        assert hit > 0 and r["stats"].paths_depth_limited == hit and r["stats"].bounce_rays == D * hit
        assert np.all(r["photons"][..., 0] == 0.0)
        assert np.all(np.abs(r["colour"]) < 1e-25)


def test_whitted_unoccluded_plane_is_colour_times_cosine():
    """whitted_integrator.rs:33-50 on a lone plane: nothing occludes, the sampled bounce escapes (adds 0):
    I = colour(l) * ds * light(l) * |light.direction . n| with the direction used UN-normalised."""
    spec = _plane_scene()
    light = spec.spectrum("grey", 0.7)
    ambient = spec.spectrum("grey", 0.05)
    orc = O.OracleScene(spec)
    W, H, seed = 12, 12, 3
    r = orc.render((0, W, 0, H), H, W, spp=1, max_depth=0, seed=seed, integrator=O.WHITTED,
                   lights=[((1.0, 1.0, -1.0), light)], ambient=ambient, want_photons=True)
    ph = r["photons"][0]
    lit = ph[:, 0] != 0
    assert lit.sum() > 30
    for p in np.nonzero(lit)[0]:
        wl = ph[p, 0]
        expect = _spectrum_at((0.55, 0.27, 0.04), wl) * 0.1 * (0.7 * 1.0) * 360.0
        assert abs(ph[p, 1] - expect) <= 1e-12 * max(1.0, expect)
    assert r["stats"].shadow_rays == lit.sum() == r["stats"].bounce_rays


def test_whitted_occluded_point_gets_ambient_only():
    """A sphere between the plane and the light: shadowed points return ambient(l) (no bsdf applied, :38)."""
    spec = _plane_scene()
    m2 = spec.lambertian_rgb((1.0, 0.0, 0.0), 0.5)
    spec.objects[0][1].append(("sphere", (0.0, 8.0, 10.0), 9.5, m2))
    light = spec.spectrum("grey", 1.0)
    ambient = spec.spectrum("grey", 0.05)
    orc = O.OracleScene(spec)
    r = orc.render((0, 16, 0, 16), 16, 16, spp=1, max_depth=0, seed=1, integrator=O.WHITTED,
                   lights=[((0.0, 1.0, 0.0), light)], ambient=ambient, want_photons=True)
    ph = r["photons"][0]
    vals = ph[ph[:, 0] != 0, 1] / 360.0
    assert np.any(np.abs(vals - 0.05) < 1e-15)        # shadowed plane points: exactly the ambient term


def test_reflective_material_is_a_deterministic_mirror():
    """reflective_material.rs:42-47 sample = mirror about z with pdf 1, no draws; bsdf at the mirror direction:
    colour*ds*(1-rs) * in + rs."""
    spec = _plane_scene("reflective")
    orc = O.OracleScene(spec)
    w_i = O.vec(0.3, -0.2, 0.9327379053088815)
    w_o, pdf, used = np.zeros(3), C.c_double(), C.c_uint32()
    L.orc_material_sample(orc.h, 0, w_i.ctypes.data_as(dp), 550.0, 1, 0, 0, 3, w_o.ctypes.data_as(dp), C.byref(pdf), C.byref(used))
    assert np.array_equal(w_o, [-0.3, 0.2, w_i[2]]) and pdf.value == 1.0 and used.value == 0
    out = L.orc_material_bsdf(orc.h, 0, w_o.ctypes.data_as(dp), w_i.ctypes.data_as(dp), 550.0, 2.0)
    c = _spectrum_at((1.0, 1.0, 0.0), 550.0)
    assert abs(out - (2.0 * c * 0.05 * (1 - 0.9) + 0.9)) < 1e-9
    below = O.vec(0.3, -0.2, -0.5)
    assert L.orc_material_bsdf(orc.h, 0, w_o.ctypes.data_as(dp), below.ctypes.data_as(dp), 550.0, 2.0) == 0.0


def test_dielectric_fresnel_conserves_and_halves():
    """smooth_transparent_dialectric.rs: R + T = 1, normal incidence R = ((n-1)/(n+1))^2, pdf is always 0.5."""
    s = scenes.SceneSpec(camera=(0, 0, 0))
    m = s.dielectric_diamond()
    s.objects.append(("list", [("sphere", (0, 0, 5), 1.0, m)]))
    orc = O.OracleScene(s)
    wl = 550.0
    w_i = O.vec(0.0, 0.0, 1.0)
    refl, trans = O.vec(0, 0, 1), O.vec(0, 0, -1)
    R = L.orc_material_bsdf(orc.h, 0, refl.ctypes.data_as(dp), w_i.ctypes.data_as(dp), wl, 1.0)
    T = L.orc_material_bsdf(orc.h, 0, trans.ctypes.data_as(dp), w_i.ctypes.data_as(dp), wl, 1.0)
    assert abs(R + T - 1.0) < 1e-15
    diamond = np.array([2.505813241, 2.487866556, 2.473323675, 2.464986815, 2.455051934, 2.441251728, 2.431478974,
                        2.427076431, 2.420857286, 2.411429037, 2.406543164, 2.406202402])
    n = L.orc_spectrum_intensity(326.27, 774.9, 12, diamond.ctypes.data_as(dp), wl)
    assert abs(R - ((n - 1) / (n + 1)) ** 2) < 1e-15
    w_o, pdf, used = np.zeros(3), C.c_double(), C.c_uint32()
    L.orc_material_sample(orc.h, 0, w_i.ctypes.data_as(dp), wl, 1, 0, 0, 3, w_o.ctypes.data_as(dp), C.byref(pdf), C.byref(used))
    assert pdf.value == 0.5 and used.value == 1
    # total internal reflection from inside (w_i.z < 0, grazing): "reflect", no draw.  As written
    # (smooth_transparent_dialectric.rs:54-57) the back-side fix-up flips z of the already mirrored direction,
    # so the result is -w_i: a quirk the build reproduces, not fixes.
    w_in = O.vec(0.9, 0.0, -np.sqrt(1 - 0.81))
    L.orc_material_sample(orc.h, 0, w_in.ctypes.data_as(dp), wl, 1, 0, 0, 3, w_o.ctypes.data_as(dp), C.byref(pdf), C.byref(used))
    assert used.value == 0 and np.allclose(w_o, -w_in)


def test_reference_and_ordered_traversal_agree_and_bvh_matches_brute_force():
    """bounding_volume_hierarchy.rs:94-120 visits both children; the ordered + pruned walk must return the same
    triangle and distance; both must equal testing every triangle in a flat list (no BVH at all)."""
    pos, nrm, faces = scenes.bunny_proxy(2)
    v, n = scenes.mesh_arrays(pos, nrm, faces)
    s1 = scenes.SceneSpec(camera=(0, 0, -5))
    m = s1.lambertian_rgb((1, 1, 0), 0.05)
    s1.objects.append(("mesh", v, n, m))
    s2 = scenes.SceneSpec(camera=(0, 0, -5))
    m = s2.lambertian_rgb((1, 1, 0), 0.05)
    s2.objects.append(("list", [("triangle", v[i], n[i], m) for i in range(len(v))]))
    a, b = O.OracleScene(s1), O.OracleScene(s2)
    rng = np.random.default_rng(5)
    o = rng.normal(size=(4000, 3)) * 4.0
    d = (rng.normal(size=(4000, 3)) * 0.7 + np.array([0, -0.5, 0])) - o
    ra = a.trace(o, d, mode=O.TRAVERSE_REFERENCE)
    rb = a.trace(o, d, mode=O.TRAVERSE_ORDERED)
    rc = b.trace(o, d)
    assert (ra[0] >= 0).sum() > 1000
    assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])
    edge = a.edge_distance(o, d) > 1e-9            # exact ties on shared edges resolve by order, which differs for a flat list
    assert np.array_equal(ra[1][edge], rc[1][edge]) and np.array_equal(ra[2][edge], rc[2][edge])
    assert rb[3].node_visits < ra[3].node_visits
