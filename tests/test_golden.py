"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle --
the Rust reference cannot run in the build image).  CPU: the oracle still reproduces them.  GPU: the CUDA path
reproduces them through the C ABI without needing the oracle at all."""
import os

import numpy as np
import pytest

import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import capi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def load(variant):
    g = np.load(os.path.join(HERE, "golden", "tiny_%s.npz" % variant))
    spec = scenes.scene_main(subdivisions=2, obj=False, variant=variant)
    return g, spec


def whitted_args(spec):
    return [((1.0, 1.0, -1.0), spec.spectrum("grey", 1.0))], spec.spectrum("grey", 0.05)


@pytest.mark.parametrize("variant", ["lambertian", "mixed"])
def test_oracle_reproduces_golden(variant):
    g, spec = load(variant)
    orc = O.OracleScene(spec)
    obj, prim, t, _ = orc.trace(g["origins"], g["dirs"])
    assert np.array_equal(obj, g["object_id"]) and np.array_equal(prim, g["prim_id"])
    assert np.array_equal(t.view(np.uint64), g["t"].view(np.uint64))
    W, H = int(g["width"]), int(g["height"])
    r = orc.render((0, W, 0, H), H, W, spp=int(g["spp"]), max_depth=int(g["max_depth"]), seed=int(g["seed"]), want_photons=True)
    assert np.array_equal(r["photons"][..., 0], g["photons"][..., 0])
    np.testing.assert_allclose(r["photons"][..., 1], g["photons"][..., 1], rtol=1e-12, atol=0)   # libm ulps across machines
    np.testing.assert_allclose(r["colour_sum"], g["colour_sum"], rtol=1e-12, atol=1e-30)
    lights, amb = whitted_args(spec)
    w = O.OracleScene(spec).render((0, W, 0, H), H, W, spp=1, max_depth=1, seed=int(g["seed"]), integrator=O.WHITTED,
                                   lights=lights, ambient=amb)
    np.testing.assert_allclose(w["colour"], g["whitted_colour"], rtol=1e-12, atol=1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["lambertian", "mixed"])
@pytest.mark.parametrize("bvh_filter", [capi.FILTER_F32, capi.FILTER_F64])
def test_gpu_reproduces_golden(variant, bvh_filter):
    g, spec = load(variant)
    lights, amb = whitted_args(spec)          # light spectra are registered with the spec before the scene is built
    hs = V.build_scene(spec)
    obj, prim, t, _ = hs.trace(g["origins"], g["dirs"], bvh_filter=bvh_filter)
    ok = (g["min_bary"] > 1e-6) & (g["dirs"].max(axis=1) > 0)
    assert np.array_equal(obj[ok], g["object_id"][ok]) and np.array_equal(prim[ok], g["prim_id"][ok])
    assert np.array_equal(t[ok].view(np.uint64), g["t"][ok].view(np.uint64))
    W, H = int(g["width"]), int(g["height"])
    r = hs.render((0, W, 0, H), H, W, spp=int(g["spp"]), max_depth=int(g["max_depth"]), seed=int(g["seed"]),
                  bvh_filter=bvh_filter, want_photons=True)
    assert np.array_equal(r["photons"][..., 0], g["photons"][..., 0])
    live = g["photons"][..., 0] != 0
    np.testing.assert_allclose(r["photons"][..., 1][live], g["photons"][..., 1][live], rtol=1e-9, atol=0)
    np.testing.assert_allclose(r["colour_sum"], g["colour_sum"], rtol=1e-8, atol=1e-20)
    w = hs.render((0, W, 0, H), H, W, spp=1, max_depth=1, seed=int(g["seed"]), integrator=capi.INTEGRATOR_WHITTED,
                  lights=lights, ambient=amb, bvh_filter=bvh_filter)
    gc = g["whitted_colour"]
    lit = np.abs(gc) > 1e-12
    assert np.max(np.abs(w["colour"][lit] - gc[lit]) / np.abs(gc[lit])) < 1e-4     # north_star's deterministic-image bar
