"""GPU tests of vrj_bvh_build: the device builder must return EXACTLY the tree the host mirror of
BoundingVolumeHierarchy::build (bounding_volume_hierarchy.rs:38-75, stable sort) returns -- same leaf order, same node
numbering, same boxes -- so hit ids and tie-breaking cannot depend on where the tree was built."""
import numpy as np
import pytest

import helpers
import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import capi, host, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    """BASELINE config C4 (11 x 11 bunny copies = 9.9 M triangles + ground plane): the GPU scene with its tree built during
    the upload, and the ORACLE's scene -- its tree built by the oracle's own restatement of
    BoundingVolumeHierarchy::build (bounding_volume_hierarchy.rs:38-75), about 20-30 s on the host cores."""
    spec = scenes.scene_grid(copies=11)
    return spec, V.build_scene(spec, device_builder="upload"), O.OracleScene(spec)


def host_tree(verts):
    """The host builder's arrays for one mesh, read back from the flattened scene description."""
    v = np.ascontiguousarray(verts, np.float64).reshape(-1, 9)
    spec = scenes.SceneSpec(camera=(0.0, 0.0, -5.0))
    m = spec.lambertian_rgb((1.0, 1.0, 0.0), 0.05)
    spec.objects.append(("mesh", v, np.zeros_like(v), m))
    hs = V.build_scene(spec)
    d = hs.desc()
    n_nodes, n = int(d.n_nodes), int(d.n_triangles)
    take = lambda p, count, dt: np.ctypeslib.as_array(p, shape=(count,)).astype(dt).copy() if count else np.zeros(0, dt)
    out = dict(order=take(d.tri_prim_id, n, np.uint32),
               node_min=take(d.node_min, 4 * n_nodes, np.float64).reshape(-1, 4),
               node_max=take(d.node_max, 4 * n_nodes, np.float64).reshape(-1, 4),
               node_child=take(d.node_child, 2 * n_nodes, np.int32).reshape(-1, 2),
               depth=int(d.bvhs[0].depth))
    return out, hs


def check_same(verts):
    ref, _ = host_tree(verts)
    got = host.bvh_build(verts)
    assert got["depth"] == ref["depth"]
    assert np.array_equal(got["order"], ref["order"]), "leaf order differs"
    assert np.array_equal(got["node_child"], ref["node_child"]), "node numbering differs"
    assert np.array_equal(got["node_min"], ref["node_min"]) and np.array_equal(got["node_max"], ref["node_max"]), "boxes differ"
    return got


def random_triangles(n, seed, extent=10.0, size=0.2):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n, 1, 3))
    v = c + rng.normal(size=(n, 3, 3)) * size
    return v.astype(np.float32).astype(np.float64).reshape(n, 9)  # f32-parsed like an OBJ (mesh.rs:21-26)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 100, 1023, 1024, 1025, 2047, 2048, 2049, 4097, 5000, 70001])
def test_device_tree_equals_host_tree_random(n):
    got = check_same(random_triangles(n, seed=n))
    if n > 1024:
        assert got["stats"].global_levels >= 1 and got["stats"].radix_passes >= 1


def test_device_tree_equals_host_tree_full_f64_keys():
    """Vertices that are NOT f32-representable: all 64 key bits vary, every radix digit pass runs."""
    rng = np.random.default_rng(7)
    v = rng.normal(size=(30000, 9)) * np.exp(rng.uniform(-20, 20, size=(30000, 1)))
    got = check_same(v)
    assert got["stats"].radix_passes >= 8


def test_device_tree_equals_host_tree_with_ties():
    """Translated copies share centre coordinates exactly (C4's grid does): the sort must be stable, level after level."""
    base = random_triangles(700, seed=3, extent=1.0)
    copies = []
    for ix in range(4):
        for iz in range(5):
            off = np.array([4.0 * ix, 0.0, 4.0 * iz] * 3)
            copies.append(base + off)
    check_same(np.concatenate(copies, 0))
    # every triangle identical: every key equal at every level, the order must stay the input order
    same = np.tile(random_triangles(1, seed=5), (3000, 1))
    got = check_same(same)
    assert np.array_equal(got["order"], np.arange(3000, dtype=np.uint32))
    # signed zeros compare equal (operator< on the host, x + 0.0 on the device)
    z = random_triangles(2500, seed=9, extent=0.0, size=1.0)
    z[::2, 0::3] = 0.0
    z[1::2, 0::3] = -0.0
    check_same(z)


def test_empty_and_error_paths():
    got = host.bvh_build(np.zeros((0, 9)))
    assert got["depth"] == 1 and got["node_child"].tolist() == [[-1, 0]]
    assert np.all(got["node_min"][0, :3] == np.inf) and np.all(got["node_max"][0, :3] == -np.inf)
    import ctypes as C
    L = capi.cuda()
    v = np.zeros(9)
    assert L.vrj_bvh_build(0, 1, v.ctypes.data_as(capi.dp), None, None, None, None, None, None) != capi.OK
    assert b"NULL" in L.vrj_last_error()


def test_scene_built_on_the_device_is_the_same_scene():
    """The bench mesh through the scene path: identical flattened arrays, identical hits and image."""
    spec = scenes.scene_main(subdivisions=6, obj=False)
    a, b = V.build_scene(spec), V.build_scene(spec, device_builder=True)
    da, db = a.desc(), b.desc()
    assert int(da.n_nodes) == int(db.n_nodes) and int(da.n_triangles) == int(db.n_triangles) == 81920
    n_nodes, n = int(da.n_nodes), int(da.n_triangles)
    arr = lambda p, count: np.ctypeslib.as_array(p, shape=(count,))
    assert np.array_equal(arr(da.tri_prim_id, n), arr(db.tri_prim_id, n))
    assert np.array_equal(arr(da.node_child, 2 * n_nodes), arr(db.node_child, 2 * n_nodes))
    assert np.array_equal(arr(da.node_min, 4 * n_nodes), arr(db.node_min, 4 * n_nodes))
    assert np.array_equal(arr(da.node_max, 4 * n_nodes), arr(db.node_max, 4 * n_nodes))
    assert np.array_equal(arr(da.tri_v0, 4 * n), arr(db.tri_v0, 4 * n))
    W, H = 160, 90
    ra = a.render((0, W, 0, H), H, W, spp=2, max_depth=4, seed=1, want=("colour_sum",), want_photons=True)
    rb = b.render((0, W, 0, H), H, W, spp=2, max_depth=4, seed=1, want=("colour_sum",), want_photons=True)
    assert np.array_equal(ra["photons"], rb["photons"])


def test_large_build_properties():
    """C4-sized input (9.9 M triangles): the tree is too big to rebuild on the host inside a test, so check the
    properties a correct build has: order is a permutation, every parent box is the union of its children's boxes,
    leaves hold their triangle's box, and the left subtree's centres do not exceed the right subtree's on the split axis."""
    spec = scenes.scene_grid(copies=11)
    v = spec.objects[1][1]
    n = v.shape[0]
    got = host.bvh_build(v)
    order, child, mn, mx = got["order"], got["node_child"], got["node_min"][:, :3], got["node_max"][:, :3]
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    internal = child[:, 0] >= 0
    l, r = child[internal, 0], child[internal, 1]
    assert np.array_equal(mn[internal], np.minimum(mn[l], mn[r])) and np.array_equal(mx[internal], np.maximum(mx[l], mx[r]))
    leaf = ~internal
    pos = ~child[leaf, 0]
    tri = v[order[pos]].reshape(-1, 3, 3)
    assert np.array_equal(mn[leaf], tri.min(1)) and np.array_equal(mx[leaf], tri.max(1))
    # root split: sorted by box centre on the largest axis
    ext = mx[0] - mn[0]
    axis = int(np.argmax(ext))
    c = (v.reshape(-1, 3, 3).min(1)[:, axis] + v.reshape(-1, 3, 3).max(1)[:, axis]) / 2.0
    half = n // 2
    assert c[order[:half]].max() <= c[order[half:]].min()
    print("C4 build: %.1f ms on the device, %d global levels, %d radix passes, %d small subtrees" % (
        got["stats"].device_ms, got["stats"].global_levels, got["stats"].radix_passes, got["stats"].small_subtrees))


def test_tree_built_at_upload_renders_the_same_image():
    """VrjBvh.n_nodes == 0: vrj_scene_create builds the tree and permutes the triangles on the device.  Hit ids (original
    primitive indices), distances and the per-sample radiance must equal those of the host-built scene bit for bit --
    including a scene with two meshes and a flat-list triangle, where triangle offsets matter."""
    import helpers
    spec = scenes.scene_main(subdivisions=4, obj=False)
    v2 = random_triangles(3000, seed=11, extent=2.0, size=0.3) + np.array([3.0, 0.5, 2.0] * 3)
    spec.objects.append(("mesh", v2, np.tile(np.array([0.0, 0.0, -1.0] * 3), (3000, 1)), 1))
    spec.objects[0][1].append(("triangle", np.array([-8.0, -2.0, 6.0, 8.0, -2.0, 6.0, 0.0, 6.0, 6.0]), np.array([0.0, 0.0, -1.0] * 3), 2))
    a, b = V.build_scene(spec), V.build_scene(spec, device_builder="upload")
    assert int(b.desc().n_nodes) == 0 and int(a.desc().n_nodes) > 0
    W, H = 200, 120
    o, d = helpers.camera_rays(W, H, spec.camera)
    ia, ib = a.trace(o, d), b.trace(o, d)
    for x, y in zip(ia[:3], ib[:3]):
        assert np.array_equal(x, y)
    for f in (capi.FILTER_F32, capi.FILTER_F64, capi.FILTER_F32X4, capi.FILTER_Q16):  # two meshes share Q16's scene-wide grid
        ra = a.render((0, W, 0, H), H, W, spp=2, max_depth=5, seed=3, want=("colour_sum",), want_photons=True, bvh_filter=f)
        rb = b.render((0, W, 0, H), H, W, spp=2, max_depth=5, seed=3, want=("colour_sum",), want_photons=True, bvh_filter=f)
        assert np.array_equal(ra["photons"], rb["photons"])
        assert ra["stats"].rays == rb["stats"].rays


def test_c4_sized_scene_filters_agree_at_4k():
    """BASELINE config C4 at full size (9.9 M triangles, 3840x2160): the tree is built during the upload, and one sample per
    pixel must come out bit-identical whether the boxes are culled in binary32, on the 16-bit grid or through 4-wide nodes --
    a size-independent property that exercises 25 tree levels and the scene-wide quantisation at C4's 44-unit extent."""
    spec = scenes.scene_grid(copies=11)
    hs = V.build_scene(spec, device_builder="upload")
    W, H = 3840, 2160
    ref = None
    for f in (capi.FILTER_F32, capi.FILTER_Q16, capi.FILTER_F32X4):
        r = hs.render((0, W, 0, H), H, W, spp=1, max_depth=4, seed=9, want=("weight",), want_photons=True, bvh_filter=f)
        assert np.all(r["weight"] == 1.0)
        if ref is None:
            ref = r
            assert r["stats"].primary_rays == W * H and np.all(np.isfinite(r["photons"]))
            assert np.count_nonzero(r["photons"][..., 1]) > 0.2 * W * H
        else:
            assert np.array_equal(r["photons"], ref["photons"]), f
            assert r["stats"].rays == ref["stats"].rays


def test_c4_full_size_against_the_oracle(c4):
    """C4 at FULL size against the oracle (parity hole 1): bounding_volume_hierarchy.rs:95-119 on the 25-level
    tree of 9.9 M triangles.  (i) 120 000 pixel-centre rays of the 3840x2160 frame: object id, primitive id and distance
    BIT-exact against the oracle's reference-order, unpruned traversal (edge rays excluded as north_star says), in every
    filter mode; (ii) a non-square tile in the middle of the meshes, at its 4K offsets, 2 spp, recursion limit 8: every
    sample's wavelength bit-identical and its radiance within 1e-12 of the oracle's, ray counts equal."""
    spec, hs, orc = c4
    W, H = 3840, 2160
    assert orc.L.orc_bvh_triangle_count(orc.h, 1) == 11 * 11 * 81920
    assert orc.L.orc_bvh_depth(orc.h, 1) == 25
    o, d = helpers.camera_rays(W, H, spec.camera)
    rng = np.random.default_rng(5)
    lower = np.flatnonzero(np.arange(W * H) // W >= 1284)          # the rows the meshes cover
    idx = np.concatenate([rng.choice(lower, 100000, replace=False), rng.choice(W * H, 20000, replace=False)])
    oo, dd = o[idx], d[idx]
    r_obj, r_prim, r_t, _ = orc.trace(oo, dd, mode=O.TRAVERSE_REFERENCE)
    ok = orc.edge_distance(oo, dd) > 1e-6
    assert (r_obj[ok] == 1).sum() > 30000                          # tens of thousands of triangle hits, not plane hits
    assert len(np.unique(r_prim[ok & (r_obj == 1)])) > 20000       # spread over the whole mesh
    for f in (capi.FILTER_F32, capi.FILTER_F64, capi.FILTER_F32X4, capi.FILTER_Q16):
        g_obj, g_prim, g_t, _ = hs.trace(oo, dd, bvh_filter=f)
        assert np.array_equal(g_obj[ok], r_obj[ok]), f
        assert np.array_equal(g_prim[ok], r_prim[ok]), f
        assert np.array_equal(g_t[ok].view(np.uint64), r_t[ok].view(np.uint64)), f
        assert ((g_obj != r_obj) | (g_prim != r_prim))[~ok].sum() <= 12
    tile = (1900, 2003, 1450, 1497)                                 # 103 x 47, inside the grid of meshes
    g = hs.render(tile, H, W, spp=2, max_depth=8, seed=9, want_photons=True)
    r = orc.render(tile, H, W, spp=2, max_depth=8, seed=9, want_photons=True)
    assert r["stats"].bounce_rays > 103 * 47                       # the crop is on the meshes: paths bounce
    assert np.array_equal(g["photons"][..., 0], r["photons"][..., 0])
    live = r["photons"][..., 0] != 0.0
    gi, ri = g["photons"][..., 1][live], r["photons"][..., 1][live]
    assert np.max(np.abs(gi - ri) / np.maximum(np.abs(ri), 1e-300)) < 1e-12
    for k in ("primary_rays", "bounce_rays", "paths_missed", "paths_escaped", "paths_depth_limited"):
        assert getattr(g["stats"], k) == getattr(r["stats"], k), k
    np.testing.assert_allclose(g["colour_sum"], r["colour_sum"], rtol=1e-9, atol=1e-25)
    assert np.all(g["weight"] == 2.0)
