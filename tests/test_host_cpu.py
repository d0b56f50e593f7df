"""CPU-only tests of the host side and of the C-ABI library's surface (no compute calls: there is no
GPU here and the library has no CPU fallback).  The oracle is used as the checker for the host logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
import oraclelib as O
import vanrijn_b200 as V
from vanrijn_b200 import capi, scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vanrijn_cuda.h")).read()
    declared = sorted(set(re.findall(r"VRJ_API[^;(]*?\b(vrj_\w+)\s*\(", header)))
    assert declared == sorted(capi.CUDA_SYMBOLS)
    lib = capi.cuda()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vrj_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors of the POD structs have the sizes and field offsets the C compiler gives the header's."""
    import subprocess
    pairs = [("VrjSpectrum", capi.Spectrum), ("VrjMaterial", capi.Material), ("VrjSphere", capi.Sphere),
             ("VrjPlane", capi.Plane), ("VrjBvh", capi.Bvh), ("VrjItem", capi.Item), ("VrjSceneDesc", capi.SceneDesc),
             ("VrjTile", capi.Tile), ("VrjSpectrumData", capi.SpectrumData), ("VrjLight", capi.Light),
             ("VrjRenderParams", capi.RenderParams), ("VrjStats", capi.Stats), ("VrjAccumOut", capi.AccumOut),
             ("VrjBvhBuildStats", capi.BvhBuildStats)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vanrijn_cuda.h"', 'int main(void){']
    for cname, ct in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)


@pytest.mark.skipif(capi.cuda().vrj_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    hs = V.build_scene(scenes.scene_main(subdivisions=1, obj=False))
    with pytest.raises(capi.VrjError):
        hs.render((0, 4, 0, 4), 4, 4, spp=1)
    with pytest.raises(capi.VrjError):
        hs.trace(np.zeros((1, 3)), np.array([[0.0, 0.0, 1.0]]))
    with pytest.raises(capi.VrjError):
        hs.partial_render_scene((0, 4, 0, 4), 4, 4)


def test_scene_validation_rejects_bad_descriptions():
    hs = V.build_scene(scenes.scene_main(subdivisions=1, obj=False))
    d = hs.desc()
    h = C.c_void_p()
    bad = capi.SceneDesc.from_buffer_copy(d)
    bad.abi_version = 99
    assert capi.cuda().vrj_scene_create(C.byref(bad), 0, C.byref(h)) == 1
    assert b"abi_version" in capi.cuda().vrj_last_error()
    bad = capi.SceneDesc.from_buffer_copy(d)
    bad.n_materials = 1          # triangle/sphere materials now out of range
    assert capi.cuda().vrj_scene_create(C.byref(bad), 0, C.byref(h)) == 1
    assert capi.cuda().vrj_scene_create(None, 0, C.byref(h)) == 1


def test_scene_validation_rejects_malformed_trees():
    """The C ABI is the trust boundary for caller-built trees: a child outside the BVH's own node range, a cycle, a node
    with two parents, a leaf triangle outside the BVH's range, an understated depth or a tree that is not full must be
    refused before anything reaches the device (the scene-prep kernels index per-BVH temporaries by child - first_node
    and the traversal stacks hold 32 entries)."""
    hs = V.build_scene(scenes.scene_main(subdivisions=1, obj=False))  # 80-triangle mesh, host-built tree
    d = hs.desc()
    nn = int(d.n_nodes)
    assert nn == 2 * 80 - 1 and d.n_bvhs == 1
    child0 = np.ctypeslib.as_array(d.node_child, (nn, 2)).copy()
    bvh0 = capi.Bvh.from_buffer_copy(d.bvhs[0])
    h = C.c_void_p()
    L = capi.cuda()

    def create(child, bvh=None):
        bad = capi.SceneDesc.from_buffer_copy(d)
        child = np.ascontiguousarray(child, np.int32)
        bad.node_child = child.ctypes.data_as(C.POINTER(C.c_int32))
        b = (capi.Bvh * 1)(bvh if bvh is not None else capi.Bvh.from_buffer_copy(bvh0))
        bad.bvhs = b
        st = L.vrj_scene_create(C.byref(bad), 0, C.byref(h))
        return st, L.vrj_last_error().decode()

    internal = np.flatnonzero(child0[:, 0] >= 0)
    leaves = np.flatnonzero(child0[:, 0] < 0)
    # the untouched description passes validation (status 2 = no CUDA device here; 0 on a GPU box)
    st, msg = create(child0)
    assert st in (0, 2), msg
    if st == 0:
        L.vrj_scene_destroy(h)
    for name, mutate in [
        ("child before its parent (cycle)", lambda c: c.__setitem__((int(internal[3]), 0), 0)),
        ("child equal to the node itself", lambda c: c.__setitem__((int(internal[2]), 1), int(internal[2]))),
        ("child past the bvh's node range", lambda c: c.__setitem__((int(internal[1]), 1), nn)),
        ("two parents", lambda c: c.__setitem__((int(internal[0]), 1), int(c[internal[1], 0]))),
        ("leaf triangle outside the bvh", lambda c: c.__setitem__((int(leaves[0]), 0), ~80)),
        ("leaf with two triangles", lambda c: c.__setitem__((int(leaves[1]), 1), 2)),
        ("triangles out of leaf order", lambda c: (c.__setitem__((int(leaves[0]), 0), int(child0[leaves[1], 0])),
                                                   c.__setitem__((int(leaves[1]), 0), int(child0[leaves[0], 0])))),
        ("internal node turned into a leaf (tree not full)", lambda c: (c.__setitem__((int(internal[-1]), 0), int(child0[leaves[0], 0])),
                                                                        c.__setitem__((int(internal[-1]), 1), 0))),
    ]:
        c = child0.copy()
        mutate(c)
        st, msg = create(c)
        assert st in (1, 3), (name, st, msg)
    # a bvh whose ranges lie outside the arrays, and one that claims nodes of another tree
    b = capi.Bvh.from_buffer_copy(bvh0)
    b.first_node = 1
    assert create(child0, b)[0] == 1
    b = capi.Bvh.from_buffer_copy(bvh0)
    b.n_nodes = nn - 2
    assert create(child0, b)[0] == 1
    # `depth` is not trusted: a degenerate 40-level chain is refused whatever the field says
    n_leaves = 41
    chain = np.zeros((2 * n_leaves - 1, 2), np.int32)
    k = 0
    for lvl in range(n_leaves - 1):  # internal node k: left = leaf k+1, right = next internal k+2
        chain[k] = (k + 1, k + 2)
        chain[k + 1] = (~lvl, 1)
        k += 2
    chain[k] = (~(n_leaves - 1), 1)
    bad = capi.SceneDesc.from_buffer_copy(d)
    assert n_leaves <= d.n_triangles and len(chain) <= nn
    b = capi.Bvh.from_buffer_copy(bvh0)
    b.n_nodes, b.n_triangles, b.depth = len(chain), n_leaves, 3
    full = child0.copy()
    full[:len(chain)] = chain
    st, msg = create(full, b)
    assert st == 3 and "deeper" in msg, (st, msg)


def test_same_signature_call_draws_fresh_sample_indices():
    """partial_render_scene(scene, tile, h, w) must not repeat samples (main.rs:199-217 merges call after call): its
    sample indices come from one process-wide counter."""
    H = capi.host()
    a = H.vrjh_next_sample_index(1)
    b = H.vrjh_next_sample_index(5)
    c = H.vrjh_next_sample_index(1)
    assert b == a + 1 and c == b + 5


def _flat_arrays(d):
    nt, nn = d.n_triangles, d.n_nodes
    tri = [np.ctypeslib.as_array(getattr(d, k), (nt, 4)) for k in ("tri_v0", "tri_v1", "tri_v2", "tri_n0", "tri_n1", "tri_n2")]
    return (tri, np.ctypeslib.as_array(d.tri_prim_id, (nt,)), np.ctypeslib.as_array(d.node_min, (nn, 4)),
            np.ctypeslib.as_array(d.node_max, (nn, 4)), np.ctypeslib.as_array(d.node_child, (nn, 2)))


@pytest.mark.parametrize("sub", [0, 2, 4])
def test_flattened_bvh_matches_reference_topology(sub):
    """bounding_volume_hierarchy.rs:49-75: 2N-1 nodes, <=1 triangle per leaf, leaves in DFS order, every node
    box = union of its children's, depth = oracle's depth; 16-byte alignment of the SoA arrays."""
    spec = scenes.scene_bench(subdivisions=sub, obj=False)
    hs, orc = V.build_scene(spec), O.OracleScene(spec)
    d = hs.desc()
    tri, prim_id, nmin, nmax, child = _flat_arrays(d)
    n = d.n_triangles
    assert n == 20 * 4 ** sub and d.n_nodes == 2 * n - 1 and d.n_bvhs == 1 and d.n_items == 1
    assert d.bvhs[0].depth == orc.L.orc_bvh_depth(orc.h, 0)
    assert sorted(prim_id.tolist()) == list(range(n))
    for k in ("tri_v0", "tri_n0", "node_min", "node_max"):
        assert C.cast(getattr(d, k), C.c_void_p).value % 16 == 0
    next_leaf = [0]

    def walk(i):
        l, r = child[i]
        if l < 0:
            assert r == 1 and ~l == next_leaf[0]          # leaves appear in DFS order
            next_leaf[0] += 1
            t = ~l
            pts = np.stack([tri[0][t, :3], tri[1][t, :3], tri[2][t, :3]])
            assert np.array_equal(nmin[i, :3], pts.min(0)) and np.array_equal(nmax[i, :3], pts.max(0))
            return nmin[i, :3], nmax[i, :3]
        lo_l, hi_l = walk(l)
        lo_r, hi_r = walk(r)
        assert np.array_equal(nmin[i, :3], np.minimum(lo_l, lo_r)) and np.array_equal(nmax[i, :3], np.maximum(hi_l, hi_r))
        return nmin[i, :3], nmax[i, :3]

    walk(0)
    assert next_leaf[0] == n


def test_median_split_on_largest_axis():
    """heuristic_split (bounding_volume_hierarchy.rs:38-46): the root's children hold len/2 and len-len/2
    triangles, separated by box-centre on the root's largest dimension."""
    spec = scenes.scene_bench(subdivisions=2, obj=False)
    hs = V.build_scene(spec)  # owns the description's arrays
    d = hs.desc()
    tri, prim_id, nmin, nmax, child = _flat_arrays(d)
    n = d.n_triangles
    axis = int(np.argmax(nmax[0, :3] - nmin[0, :3]))
    pts = np.stack([tri[0][:, :3], tri[1][:, :3], tri[2][:, :3]], 1)
    centre = (pts.min(1)[:, axis] + pts.max(1)[:, axis]) / 2.0
    assert centre[: n // 2].max() <= centre[n // 2:].min()


def test_obj_loader_matches_oracle_loader(tmp_path):
    """mesh.rs:13-88 / obj 0.9: index forms a, a/b, a//c, a/b/c, negative indices, polygons (fan), f32 parse."""
    p = tmp_path / "t.obj"
    p.write_text("# test\nv 0 0 0\nv 1 0 0.1\nv 1 1 0\nv 0 1 0.333333343\nv 0.5 0.5 1e-3\nvn 0 0 1\nvn 0 1 0\nvt 0 0\n"
                 "f 1//1 2//1 3//2\nf 1 2 3 4\nf 1/1/1 3/1/2 5/1/1\nf -1//-1 -2//-2 -3//-1 -4//-2 -5//1\nf 1/1 2/1 5/1\n")
    verts = np.zeros((16, 9))
    norms = np.zeros((16, 9))
    n = capi.host().vrjh_load_obj(str(p).encode(), verts.ctypes.data_as(capi.dp), norms.ctypes.data_as(capi.dp), 16)
    pv, pn = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    m = O.lib().orc_load_obj(str(p).encode(), C.byref(pv), C.byref(pn))
    assert n == m == 1 + 2 + 1 + 3 + 1
    ov = np.ctypeslib.as_array(pv, (m, 9)).copy()
    on = np.ctypeslib.as_array(pn, (m, 9)).copy()
    assert np.array_equal(verts[:n], ov) and np.array_equal(norms[:n], on)
    assert verts[1, 8] == 0.0 and verts[2, 8] == float(np.float32(0.333333343))    # fan: (v0,v1,v2),(v0,v2,v3); f32 widened
    assert np.all(norms[1] == 0.0)                                                    # no normal index -> zero normal
    assert capi.host().vrjh_load_obj(b"/nonexistent.obj", verts.ctypes.data_as(capi.dp), norms.ctypes.data_as(capi.dp), 16) == -1


def test_proxy_obj_round_trips_through_the_loader():
    path, name = scenes.bunny_obj_path(subdivisions=3)
    pos, nrm, faces = scenes.bunny_proxy(3)
    v, n = scenes.mesh_arrays(pos, nrm, faces)
    verts = np.zeros((len(v), 9))
    norms = np.zeros((len(v), 9))
    cnt = capi.host().vrjh_load_obj(path.encode(), verts.ctypes.data_as(capi.dp), norms.ctypes.data_as(capi.dp), len(v))
    assert cnt == len(v) == 1280
    assert np.array_equal(verts, v) and np.array_equal(norms, n)


def test_plane_and_items_flatten_like_the_reference_constructs_them():
    spec = scenes.scene_main(subdivisions=1, obj=False)
    hs = V.build_scene(spec)  # owns the description's arrays
    d = hs.desc()
    assert [d.items[i].kind for i in range(d.n_items)] == [1, 0, 0, 0, 3]      # plane, 3 spheres, bvh: Scene.objects order
    assert [d.items[i].object_id for i in range(d.n_items)] == [0, 0, 0, 0, 1]
    assert [d.items[i].prim_id for i in range(4)] == [0, 1, 2, 3]
    p = d.planes[0]
    # Plane::new (plane.rs:17-32) for normal (0,1,0): smallest coord -> z axis (x == z -> index 2)
    out = O.hit16(O.lib().orc_plane_intersect, O.vec(0, 1, 0), -2.0, O.vec(0, 0, 0), O.vec(0, -1, 0))
    assert np.array_equal(np.array(p.normal[:]), out["normal"])
    assert np.array_equal(np.array(p.tangent[:]), out["tangent"]) and np.array_equal(np.array(p.cotangent[:]), out["cotangent"])
    # spectra: reflection_from_linear_rgb on the host == oracle's
    s = np.zeros(32)
    O.lib().orc_rgb_to_spectrum(0.55, 0.27, 0.04, s.ctypes.data_as(O.dp))
    sp = d.spectra[d.materials[0].spectrum]
    got = np.ctypeslib.as_array(d.spectrum_samples, (d.n_spectrum_samples,))[sp.first_sample: sp.first_sample + sp.n_samples]
    assert np.array_equal(got, s) and sp.shortest_wavelength == 380.0 and sp.longest_wavelength == 720.0


def test_tile_iterator_and_merge_tile_match_the_oracle():
    cap = 512
    a, b = (C.c_uint64 * (4 * cap))(), (C.c_uint64 * (4 * cap))()
    for w, h, ts in [(20, 15, 5), (21, 16, 5), (640, 480, 32), (7, 3, 2048)]:
        n = capi.host().vrjh_tile_iterator(w, h, ts, a, cap)
        assert n == O.lib().orc_tile_iterator(w, h, ts, b, cap)
        assert list(a[:4 * n]) == list(b[:4 * n])
    rng = np.random.default_rng(1)
    # widths 4 and 11: whole 4-pixel vectors, and vectors + a scalar remainder (the AVX2 row of merge_tile)
    for W, H, tile in [(16, 12, (3, 7, 4, 9)), (16, 12, (1, 12, 2, 9))]:
        tw, th = tile[1] - tile[0], tile[3] - tile[2]
        dc, dw = rng.random((H, W, 3)), rng.random((H, W)) + 0.1
        sc_, sw = rng.random((th, tw, 3)), rng.random((th, tw)) + 0.1
        got_c, got_w = dc.copy(), dw.copy()
        t4 = (C.c_uint64 * 4)(*tile)
        assert capi.host().vrjh_merge_tile(got_c.ctypes.data_as(capi.dp), got_w.ctypes.data_as(capi.dp), W, H, t4,
                                           np.ascontiguousarray(sc_).ctypes.data_as(capi.dp), np.ascontiguousarray(sw).ctypes.data_as(capi.dp)) == 0
        for i in range(th):
            for j in range(tw):
                out = np.zeros(3)
                O.lib().orc_accum_blend(dc[tile[2] + i, tile[0] + j].copy().ctypes.data_as(O.dp), dw[tile[2] + i, tile[0] + j],
                                        sc_[i, j].copy().ctypes.data_as(O.dp), sw[i, j], out.ctypes.data_as(O.dp))
                assert np.array_equal(got_c[tile[2] + i, tile[0] + j], out)
                assert got_w[tile[2] + i, tile[0] + j] == dw[tile[2] + i, tile[0] + j] + sw[i, j]
        mask = np.ones((H, W), bool)
        mask[tile[2]:tile[3], tile[0]:tile[1]] = False
        assert np.array_equal(got_c[mask], dc[mask]) and np.array_equal(got_w[mask], dw[mask])
    # a tile that leaves the destination is an error (the reference panics on the Array2D index, array2d.rs:58-66)
    t_bad = (C.c_uint64 * 4)(14, 18, 4, 9)
    assert capi.host().vrjh_merge_tile(got_c.ctypes.data_as(capi.dp), got_w.ctypes.data_as(capi.dp), W, H, t_bad,
                                       np.zeros(60).ctypes.data_as(capi.dp), np.zeros(20).ctypes.data_as(capi.dp)) == 1


def test_merge_tile_large_tile_threaded_path():
    """Tiles of >= 2^18 pixels are merged by several threads: every pixel must still be the blend of accumulation_buffer.rs:81-85
    ((c1*w1 + c2*w2) * (1/(w1+w2)), weight summed), here checked against the same expression in numpy (binary64, same order)."""
    rng = np.random.default_rng(4)
    W, H = 700, 520
    tile = (20, 660, 10, 510)  # 640 x 500 = 320 000 pixels
    th, tw = tile[3] - tile[2], tile[1] - tile[0]
    dc, dw = rng.random((H, W, 3)), rng.random((H, W)) + 0.1
    sc_, sw = rng.random((th, tw, 3)), rng.random((th, tw)) + 0.1
    got_c, got_w = dc.copy(), dw.copy()
    t4 = (C.c_uint64 * 4)(*tile)
    assert capi.host().vrjh_merge_tile(got_c.ctypes.data_as(capi.dp), got_w.ctypes.data_as(capi.dp), W, H, t4,
                                       np.ascontiguousarray(sc_).ctypes.data_as(capi.dp), np.ascontiguousarray(sw).ctypes.data_as(capi.dp)) == 0
    w1 = dw[tile[2]:tile[3], tile[0]:tile[1]]
    inv = 1.0 / (w1 + sw)
    want = (dc[tile[2]:tile[3], tile[0]:tile[1]] * w1[..., None] + sc_ * sw[..., None]) * inv[..., None]
    assert np.array_equal(got_c[tile[2]:tile[3], tile[0]:tile[1]], want)
    assert np.array_equal(got_w[tile[2]:tile[3], tile[0]:tile[1]], w1 + sw)
    mask = np.ones((H, W), bool)
    mask[tile[2]:tile[3], tile[0]:tile[1]] = False
    assert np.array_equal(got_c[mask], dc[mask]) and np.array_equal(got_w[mask], dw[mask])


def test_shard_samples_partition_the_sample_indices():
    for world in (1, 2, 4, 8):
        for spp in (1, 3, 16):
            seen = []
            for step in range(3):
                for rank in range(world):
                    seen += sharding.shard_sample_indices(rank, world, step, spp)
            assert sorted(seen) == list(range(3 * spp * world))
    with pytest.raises(ValueError):
        sharding.shard_samples(2, 2, 0, 1)


def test_array_path_and_per_triangle_object_path_flatten_identically():
    """BoundingVolumeHierarchy::build(TriangleMesh) (arrays) and build(Vec<Arc<dyn Primitive>>) (one object per triangle,
    as mesh.rs:74-88 + bounding_volume_hierarchy.rs:49-51 do it) must flatten to the same scene, for a mesh given as
    arrays and for one read from OBJ text."""
    for spec in (scenes.tiny_mesh_scene(subdivisions=3), scenes.scene_main(subdivisions=2, obj=True)):
        sa, sb = V.build_scene(spec), V.build_scene(spec, per_triangle_objects=True)  # own the arrays desc() points into
        a, b = sa.desc(), sb.desc()
        assert int(a.n_triangles) == int(b.n_triangles) > 0 and int(a.n_nodes) == int(b.n_nodes) > 0
        n, nn = int(a.n_triangles), int(a.n_nodes)
        arr = lambda p, count: np.ctypeslib.as_array(p, shape=(count,))
        for name, count in (("tri_v0", 4 * n), ("tri_v1", 4 * n), ("tri_v2", 4 * n), ("tri_n0", 4 * n), ("tri_n1", 4 * n),
                            ("tri_n2", 4 * n), ("tri_material", n), ("tri_prim_id", n), ("node_min", 4 * nn),
                            ("node_max", 4 * nn), ("node_child", 2 * nn)):
            assert np.array_equal(arr(getattr(a, name), count), arr(getattr(b, name), count)), name


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (300, 240)])
def test_write_png_round_trips(tmp_path, w, h):
    """image.rs:52-66: the PNG decodes (zlib's inflate + CRC/Adler checks) to exactly the pixels written; 300x240x3 + filter
    bytes > 65535 exercises several stored deflate blocks."""
    rng = np.random.default_rng(w * 1000 + h)
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    path = str(tmp_path / "t.png")
    assert capi.host().vrjh_write_png(path.encode(), w, h, rgb.ctypes.data) == 0
    assert np.array_equal(helpers.read_png_rgb8(path), rgb)


def test_scene_cache_round_trip_and_corruption(tmp_path):
    """the flattened scene saved to disk reloads to the same VrjSceneDesc, for a host-built tree and for a
    tree left to the device (n_nodes == 0); a flipped byte or a truncated file is refused."""
    for builder in (False, "upload"):
        hs = V.build_scene(scenes.scene_main(subdivisions=3, obj=False, variant="mixed"), device_builder=builder)
        path = tmp_path / ("scene_%s.vrjscene" % builder)
        hs.save_cache(path)
        back = V.HostScene.from_cache(path)
        a, b = hs.desc(), back.desc()
        for f in ("n_spectra", "n_spectrum_samples", "n_materials", "n_spheres", "n_planes", "n_bvhs", "n_triangles", "n_nodes", "n_items"):
            assert int(getattr(a, f)) == int(getattr(b, f)), f
        assert list(a.camera_location) == list(b.camera_location)
        n, nn = int(a.n_triangles), int(a.n_nodes)
        arr = lambda p, count, dt=None: np.ctypeslib.as_array(p, shape=(count,)) if count else np.zeros(0)
        for name, count in (("tri_v0", 4 * n), ("tri_v2", 4 * n), ("tri_n1", 4 * n), ("tri_material", n), ("tri_prim_id", n),
                            ("node_min", 4 * nn), ("node_max", 4 * nn), ("node_child", 2 * nn), ("spectrum_samples", int(a.n_spectrum_samples))):
            assert np.array_equal(arr(getattr(a, name), count), arr(getattr(b, name), count)), name
        raw = lambda p, count, T: bytes(C.string_at(C.cast(p, C.c_void_p), count * C.sizeof(T)))
        for name, count, T in (("spectra", a.n_spectra, capi.Spectrum), ("materials", a.n_materials, capi.Material),
                               ("spheres", a.n_spheres, capi.Sphere), ("planes", a.n_planes, capi.Plane),
                               ("bvhs", a.n_bvhs, capi.Bvh), ("items", a.n_items, capi.Item)):
            assert raw(getattr(a, name), int(count), T) == raw(getattr(b, name), int(count), T), name
        data = bytearray(open(path, "rb").read())
        data[len(data) // 2] ^= 0x40
        bad = tmp_path / "bad.vrjscene"
        bad.write_bytes(bytes(data))
        with pytest.raises(capi.VrjError, match="checksum"):
            V.HostScene.from_cache(bad)
        bad.write_bytes(bytes(data[:100]))
        with pytest.raises(capi.VrjError):
            V.HostScene.from_cache(bad)
    with pytest.raises(capi.VrjError):
        V.HostScene.from_cache(tmp_path / "missing.vrjscene")


def test_rust_sys_crate_declares_every_header_symbol():
    """rust/vanrijn-cuda-sys cannot be compiled here (no cargo), so at least keep its extern block in step with the header."""
    header = open(os.path.join(ROOT, "include", "vanrijn_cuda.h")).read()
    crate = open(os.path.join(ROOT, "rust", "vanrijn-cuda-sys", "src", "lib.rs")).read()
    for name in sorted(set(re.findall(r"VRJ_API[^;(]*?\b(vrj_\w+)\s*\(", header))):
        assert ("fn %s(" % name) in crate, name
    for const in re.findall(r"\b(VRJ_(?:FILTER|PRECISION|MEM|ITEM|MAT|INTEGRATOR|TONEMAP)_\w+)\s*=\s*(\d+)", header):
        m = re.search(r"pub const %s: u32 = (\d+);" % const[0], crate)
        assert m and m.group(1) == const[1], const


def test_merge_tiles_equals_consecutive_merge_tile_calls():
    """merge_tiles(tile, [a, b, c]) -- what the main.rs loop does with several waiting messages -- must equal
    merge_tile(a); merge_tile(b); merge_tile(c) bit for bit (accumulation_buffer.rs:62-85 per pixel, in arrival order), for
    buffers with per-pixel weights and for colour-only buffers that carry one weight; small (single-thread) and large tiles."""
    rng = np.random.default_rng(9)
    H = capi.host()
    for W, Hh, tile in [(40, 30, (3, 34, 2, 29)), (700, 520, (20, 659, 10, 510))]:
        tw, th = tile[1] - tile[0], tile[3] - tile[2]
        t4 = (C.c_uint64 * 4)(*tile)
        for n, uniform in [(1, False), (3, False), (5, True)]:
            dc, dw = rng.random((Hh, W, 3)), rng.random((Hh, W)) + 0.1
            sc_ = rng.random((n, th, tw, 3))
            sw = np.full((n, th, tw), 2.0) if uniform else rng.random((n, th, tw)) + 0.1
            seq_c, seq_w = dc.copy(), dw.copy()
            for k in range(n):
                assert H.vrjh_merge_tile(seq_c.ctypes.data_as(capi.dp), seq_w.ctypes.data_as(capi.dp), W, Hh, t4,
                                         np.ascontiguousarray(sc_[k]).ctypes.data_as(capi.dp), np.ascontiguousarray(sw[k]).ctypes.data_as(capi.dp)) == 0
            got_c, got_w = dc.copy(), dw.copy()
            assert H.vrjh_merge_tiles(got_c.ctypes.data_as(capi.dp), got_w.ctypes.data_as(capi.dp), W, Hh, t4, n, sc_.ctypes.data_as(capi.dp),
                                      None if uniform else sw.ctypes.data_as(capi.dp), 2.0) == 0
            assert np.array_equal(got_c, seq_c) and np.array_equal(got_w, seq_w)
            # and against the expression itself, in numpy (binary64, same order)
            want_c, want_w = dc[tile[2]:tile[3], tile[0]:tile[1]].copy(), dw[tile[2]:tile[3], tile[0]:tile[1]].copy()
            for k in range(n):
                inv = 1.0 / (want_w + sw[k])
                want_c = (want_c * want_w[..., None] + sc_[k] * sw[k][..., None]) * inv[..., None]
                want_w = want_w + sw[k]
            assert np.array_equal(got_c[tile[2]:tile[3], tile[0]:tile[1]], want_c) and np.array_equal(got_w[tile[2]:tile[3], tile[0]:tile[1]], want_w)
