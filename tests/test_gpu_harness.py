"""GPU test of the harness (examples/vanrijn_main.cpp = src/main.rs:104-247 without the SDL window): the PNG it writes
holds exactly the bytes ClampingToneMapper gives for the frame the library renders."""
import os
import subprocess

import numpy as np
import pytest

import helpers
import vanrijn_b200 as V
from vanrijn_b200 import host, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def harness_binary():
    """build/vanrijn is built by `make` / __graft_entry__.build(); if only the libraries travelled, link it here (g++ only)."""
    exe = os.path.join(ROOT, "build", "vanrijn")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", ROOT, "build/vanrijn"], check=True)
    return exe


@pytest.mark.parametrize("size,builder", [((160, 90), "upload"), ((2100, 6), "host")])
def test_harness_png_equals_library_frame(tmp_path, size, builder):
    exe = os.path.join(ROOT, "build", "vanrijn")
    assert os.path.exists(exe), "build/vanrijn is missing: run make (or __graft_entry__.build())"
    W, H = size
    obj, _ = scenes.bunny_obj_path(subdivisions=3)
    out = str(tmp_path / "frame.png")
    r = subprocess.run([exe, "--size", str(W), str(H), "--out", out, "--obj", obj, "--spp", "4", "--depth", "8",
                        "--builder", builder, "--seed", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Mrays/s" in r.stdout and "wrote" in r.stdout
    png = helpers.read_png_rgb8(out)
    assert png.shape == (H, W, 3)
    # the same frame through the library: one call, 4 spp (2100 px wide = two 2048-px tiles in the harness; samples are
    # pure functions of (seed, pixel, sample), so tiling cannot change them); merging a 4-sample tile into an empty
    # frame is exact (x*4/4)
    hs = V.build_scene(scenes.scene_main(subdivisions=3, obj=True))
    ref = hs.render((0, W, 0, H), H, W, spp=4, max_depth=8, seed=5, want=("colour",))
    want = host.tone_map(ref["colour"]).reshape(H, W, 3)
    assert np.array_equal(png, want)
    assert png.max() > 0


def test_harness_resident_frame_equals_one_long_accumulation(tmp_path):
    """--resident: the frame stays on the GPU and every pass continues the Kahan accumulators, so 3 passes x 2 spp
    give exactly the frame of one 6-spp call (update_pixel applied 6 times per pixel, accumulation_buffer.rs:44-60)."""
    exe = os.path.join(ROOT, "build", "vanrijn")
    W, H = 192, 108
    obj, _ = scenes.bunny_obj_path(subdivisions=3)
    out = str(tmp_path / "resident.png")
    r = subprocess.run([exe, "--size", str(W), str(H), "--out", out, "--obj", obj, "--spp", "2", "--passes", "3", "--depth", "8",
                        "--resident", "--preview-every", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    hs = V.build_scene(scenes.scene_main(subdivisions=3, obj=True))
    ref = hs.render((0, W, 0, H), H, W, spp=6, max_depth=8, seed=1, want=("colour",))
    assert np.array_equal(helpers.read_png_rgb8(out), host.tone_map(ref["colour"]).reshape(H, W, 3))


def test_harness_time_limit_and_usage_errors(tmp_path):
    exe = os.path.join(ROOT, "build", "vanrijn")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "--size" in r.stderr
    out = str(tmp_path / "t.png")
    r = subprocess.run([exe, "--size", "64", "36", "--time", "0.2", "--spp", "2", "--depth", "4", "--out", out,
                        "--preview-every", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    passes = int(r.stdout.split(" passes")[0].split()[-1])
    assert passes >= 1
    assert helpers.read_png_rgb8(out).shape == (36, 64, 3)


def test_scene_cache_renders_the_same_frame(tmp_path):
    """--cache: the first run parses the OBJ and saves the flattened scene, the second run loads it (no OBJ, no primitive
    objects) and must write the same PNG; the library renders the cached scene bit-identically too."""
    exe = os.path.join(ROOT, "build", "vanrijn")
    obj, _ = scenes.bunny_obj_path(subdivisions=3)
    cache = str(tmp_path / "scene.vrjscene")
    outs = []
    for run in range(2):
        out = str(tmp_path / ("c%d.png" % run))
        args = [exe, "--size", "160", "90", "--out", out, "--spp", "3", "--depth", "6", "--cache", cache]
        r = subprocess.run(args + (["--obj", obj] if run == 0 else []), capture_output=True, text=True, timeout=300,
                           env={k: v for k, v in os.environ.items() if k != "VANRIJN_BUNNY_OBJ"})
        assert r.returncode == 0, r.stderr
        assert ("Saved the flattened scene" in r.stdout) if run == 0 else ("Loaded the flattened scene" in r.stdout)
        outs.append(helpers.read_png_rgb8(out))
    assert np.array_equal(outs[0], outs[1]) and outs[0].max() > 0
    spec = scenes.scene_main(subdivisions=3, obj=True)
    a = V.build_scene(spec)
    a.save_cache(tmp_path / "lib.vrjscene")
    b = V.HostScene.from_cache(tmp_path / "lib.vrjscene")
    ra = a.render((0, 96, 0, 54), 54, 96, spp=2, max_depth=6, seed=2, want=("colour_sum",), want_photons=True)
    rb = b.render((0, 96, 0, 54), 54, 96, spp=2, max_depth=6, seed=2, want=("colour_sum",), want_photons=True)
    assert np.array_equal(ra["photons"], rb["photons"])
