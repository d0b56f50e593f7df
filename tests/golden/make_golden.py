#!/usr/bin/env python
"""Generates tests/golden/*.npz.

The reference is a Rust crate that cannot be built or imported in the build image, so these vectors are
outputs of the CPU ORACLE (oracle/vanrijn_oracle.cpp), not of the reference itself: they pin the oracle
against regressions and give the GPU tests a fixture that does not depend on rebuilding the oracle.  The
reference's own known-answer tests are restated directly in tests/test_oracle_kat.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
import oraclelib as O  # noqa: E402
from vanrijn_b200 import scenes  # noqa: E402


def golden_scene(variant):
    spec = scenes.scene_main(subdivisions=2, obj=False, variant=variant)
    return spec


def main():
    for variant in ("lambertian", "mixed"):
        spec = golden_scene(variant)
        orc = O.OracleScene(spec)
        W, H, spp = 32, 18, 3
        o, d = helpers.camera_rays(W, H, spec.camera)
        o2, d2 = helpers.sphere_rays(424, (-3.0, -0.5, 1.0), 6.0, seed=1)
        o, d = np.concatenate([o, o2]), np.concatenate([d, d2])
        obj, prim, t, _ = orc.trace(o, d)
        edge = orc.edge_distance(o, d)
        r = orc.render((0, W, 0, H), H, W, spp=spp, max_depth=16, seed=42, want_photons=True)
        lights = [((1.0, 1.0, -1.0), spec.spectrum("grey", 1.0))]
        amb = spec.spectrum("grey", 0.05)
        orc_w = O.OracleScene(spec)
        w = orc_w.render((0, W, 0, H), H, W, spp=1, max_depth=1, seed=42, integrator=O.WHITTED, lights=lights, ambient=amb)
        np.savez_compressed(os.path.join(HERE, "tiny_%s.npz" % variant), origins=o, dirs=d, object_id=obj, prim_id=prim, t=t,
                            min_bary=edge, photons=r["photons"], colour_sum=r["colour_sum"], weight=r["weight"],
                            whitted_colour=w["colour"], width=W, height=H, spp=spp, max_depth=16, seed=42)
        print(variant, "hits", int((obj >= 0).sum()), "of", len(obj))


if __name__ == "__main__":
    main()
