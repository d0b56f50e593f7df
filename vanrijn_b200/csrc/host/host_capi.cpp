// host_capi.cpp -- a flat C view of the C++ host mirror (include/vanrijn.hpp) so that scripts
// (the pytest suite, bench.py) can build scenes through the SAME host code a C++ caller uses:
// load_obj, BoundingVolumeHierarchy::build, flattening, and partial_render_scene.
#include "../../../include/vanrijn.hpp"

#include <cstring>
#include <map>
#include <string>

using namespace vanrijn;

namespace {
thread_local std::string g_host_error;
struct HostScene {
    Scene scene;
    std::vector<Spectrum> spectra;
    std::vector<std::shared_ptr<Material>> materials;
    PrimitiveList *open_list = nullptr;
    FlatSceneBuilder flat;
    bool flattened = false;
};
template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return 1;
    }
}
} // namespace

extern "C" {

const char *vrjh_last_error(void) { return g_host_error.c_str(); }

void *vrjh_scene_new(double cx, double cy, double cz) {
    HostScene *h = new HostScene();
    h->scene.camera_location = Vec3(cx, cy, cz);
    return h;
}
void vrjh_scene_free(void *p) { delete static_cast<HostScene *>(p); }

int vrjh_add_spectrum(void *p, double lo, double hi, int n, const double *samples) {
    HostScene *h = static_cast<HostScene *>(p);
    Spectrum s;
    s.shortest_wavelength = lo, s.longest_wavelength = hi;
    s.samples.assign(samples, samples + n);
    h->spectra.push_back(s);
    return (int)h->spectra.size() - 1;
}
int vrjh_add_spectrum_rgb(void *p, double r, double g, double b) {
    HostScene *h = static_cast<HostScene *>(p);
    h->spectra.push_back(Spectrum::reflection_from_linear_rgb(ColourRgbF(r, g, b)));
    return (int)h->spectra.size() - 1;
}
int vrjh_add_spectrum_grey(void *p, double v) {
    HostScene *h = static_cast<HostScene *>(p);
    h->spectra.push_back(Spectrum::grey(v));
    return (int)h->spectra.size() - 1;
}
int vrjh_add_spectrum_diamond(void *p) {
    HostScene *h = static_cast<HostScene *>(p);
    h->spectra.push_back(Spectrum::diamond_index_of_refraction());
    return (int)h->spectra.size() - 1;
}
/* copies spectrum `id` out: returns n_samples, fills lo/hi and up to cap samples */
int vrjh_get_spectrum(void *p, int id, double *lo, double *hi, double *samples, int cap) {
    HostScene *h = static_cast<HostScene *>(p);
    if (id < 0 || id >= (int)h->spectra.size()) {
        g_host_error = "spectrum id out of range";
        return -1;
    }
    const Spectrum &s = h->spectra[id];
    *lo = s.shortest_wavelength, *hi = s.longest_wavelength;
    for (int i = 0; i < (int)s.samples.size() && i < cap; i++) samples[i] = s.samples[i];
    return (int)s.samples.size();
}
int vrjh_add_material(void *p, int kind, int spectrum, double p0, double p1, double p2) {
    HostScene *h = static_cast<HostScene *>(p);
    if (spectrum < 0 || spectrum >= (int)h->spectra.size()) {
        g_host_error = "spectrum id out of range";
        return -1;
    }
    const Spectrum &s = h->spectra[spectrum];
    std::shared_ptr<Material> m;
    switch (kind) {
    case VRJ_MAT_LAMBERTIAN: m = std::make_shared<LambertianMaterial>(s, p0); break;
    case VRJ_MAT_PHONG: m = std::make_shared<PhongMaterial>(s, p0, p1, p2); break;
    case VRJ_MAT_REFLECTIVE: m = std::make_shared<ReflectiveMaterial>(s, p0, p1); break;
    default: m = std::make_shared<SmoothTransparentDialectric>(s); break;
    }
    h->materials.push_back(m);
    return (int)h->materials.size() - 1;
}
int vrjh_begin_list(void *p) {
    HostScene *h = static_cast<HostScene *>(p);
    std::unique_ptr<PrimitiveList> l(new PrimitiveList());
    h->open_list = l.get();
    h->scene.objects.push_back(std::move(l));
    return (int)h->scene.objects.size() - 1;
}
void vrjh_list_add_sphere(void *p, double cx, double cy, double cz, double r, int material) {
    HostScene *h = static_cast<HostScene *>(p);
    h->open_list->primitives.push_back(std::make_shared<Sphere>(Vec3(cx, cy, cz), r, h->materials.at(material)));
}
void vrjh_list_add_plane(void *p, double nx, double ny, double nz, double d, int material) {
    HostScene *h = static_cast<HostScene *>(p);
    h->open_list->primitives.push_back(std::make_shared<Plane>(Vec3(nx, ny, nz), d, h->materials.at(material)));
}
void vrjh_list_add_triangle(void *p, const double *v, const double *n, int material) {
    HostScene *h = static_cast<HostScene *>(p);
    std::array<Vec3, 3> vs{Vec3(v[0], v[1], v[2]), Vec3(v[3], v[4], v[5]), Vec3(v[6], v[7], v[8])};
    std::array<Vec3, 3> ns{Vec3(n[0], n[1], n[2]), Vec3(n[3], n[4], n[5]), Vec3(n[6], n[7], n[8])};
    h->open_list->primitives.push_back(std::make_shared<Triangle>(vs, ns, h->materials.at(material)));
}
/* BoundingVolumeHierarchy::build over ntri triangles; returns object id or -1.
 * builder: 0 host, 1 device (vrj_bvh_build), 2 at upload (inside vrj_scene_create);
 * +4 = go through one Triangle object per triangle as the reference's Vec<Arc<dyn Primitive>> does (same result, slower) */
int vrjh_add_bvh(void *p, int64_t ntri, const double *verts, const double *normals, int material, int builder) {
    HostScene *h = static_cast<HostScene *>(p);
    int id = -1;
    guarded([&] {
        std::shared_ptr<Material> m = h->materials.at(material);
        const auto b = static_cast<BoundingVolumeHierarchy::Builder>(builder & 3);
        if (builder & 4) {
            std::vector<std::shared_ptr<Primitive>> prims((size_t)ntri);
            for (int64_t i = 0; i < ntri; i++) {
                const double *v = verts + 9 * i, *n = normals + 9 * i;
                std::array<Vec3, 3> vs{Vec3(v[0], v[1], v[2]), Vec3(v[3], v[4], v[5]), Vec3(v[6], v[7], v[8])};
                std::array<Vec3, 3> ns{Vec3(n[0], n[1], n[2]), Vec3(n[3], n[4], n[5]), Vec3(n[6], n[7], n[8])};
                prims[i] = std::make_shared<Triangle>(vs, ns, m);
            }
            h->scene.objects.push_back(BoundingVolumeHierarchy::build(prims, b));
        } else {
            TriangleMesh mesh;
            mesh.vertices.assign(verts, verts + 9 * ntri), mesh.normals.assign(normals, normals + 9 * ntri);
            h->scene.objects.push_back(BoundingVolumeHierarchy::build(mesh, m, b));
        }
        id = (int)h->scene.objects.size() - 1;
    });
    return id;
}
/* load_obj + BoundingVolumeHierarchy::build; returns object id or -1 (builder as above) */
int vrjh_add_bvh_obj(void *p, const char *path, int material, int builder) {
    HostScene *h = static_cast<HostScene *>(p);
    int id = -1;
    guarded([&] {
        const auto b = static_cast<BoundingVolumeHierarchy::Builder>(builder & 3);
        if (builder & 4) {
            auto prims = load_obj(path, h->materials.at(material));
            h->scene.objects.push_back(BoundingVolumeHierarchy::build(prims, b));
        } else {
            h->scene.objects.push_back(BoundingVolumeHierarchy::build(load_obj_mesh(path), h->materials.at(material), b));
        }
        id = (int)h->scene.objects.size() - 1;
    });
    return id;
}
/* load_obj alone: triangle count, or -1; fills verts/normals (9 doubles per triangle) up to cap triangles */
int64_t vrjh_load_obj(const char *path, double *verts, double *normals, int64_t cap) {
    int64_t n = -1;
    guarded([&] {
        auto prims = load_obj(path, std::make_shared<LambertianMaterial>(Spectrum::black(), 1.0));
        n = (int64_t)prims.size();
        for (int64_t i = 0; i < n && i < cap; i++) {
            const Triangle *t = static_cast<const Triangle *>(prims[i].get());
            for (int k = 0; k < 3; k++) {
                verts[9 * i + 3 * k] = t->vertices[k].x, verts[9 * i + 3 * k + 1] = t->vertices[k].y, verts[9 * i + 3 * k + 2] = t->vertices[k].z;
                normals[9 * i + 3 * k] = t->normals[k].x, normals[9 * i + 3 * k + 1] = t->normals[k].y, normals[9 * i + 3 * k + 2] = t->normals[k].z;
            }
        }
    });
    return n;
}
/* the flattened SoA description (owned by the scene handle; valid until it is freed) */
const VrjSceneDesc *vrjh_flatten(void *p) {
    HostScene *h = static_cast<HostScene *>(p);
    const VrjSceneDesc *d = nullptr;
    guarded([&] {
        if (h->scene.flattened) {
            d = &h->scene.flattened->desc();
            return;
        }
        if (!h->flattened) {
            for (size_t i = 0; i < h->scene.objects.size(); i++) h->scene.objects[i]->flatten(h->flat, (uint32_t)i);
            h->flattened = true;
        }
        d = &h->flat.desc(h->scene.camera_location);
    });
    return d;
}
/* flatten + upload (cached); NULL on failure */
const VrjScene *vrjh_device_scene(void *p, int device) {
    HostScene *h = static_cast<HostScene *>(p);
    const VrjScene *s = nullptr;
    guarded([&] { s = device_scene(h->scene, device); });
    return s;
}
/* partial_render_scene(scene, tile, height, width) with the reference's signature and defaults
 * (1 spp, RECURSION_LIMIT 128, SimpleRandomIntegrator).  Arrays are tile.width*tile.height. */
int vrjh_partial_render_scene(void *p, const uint64_t tile[4], uint64_t height, uint64_t width, uint64_t seed,
                              uint64_t sample_offset, double *colour, double *colour_sum, double *colour_bias,
                              double *weight, double *weight_bias) {
    HostScene *h = static_cast<HostScene *>(p);
    return guarded([&] {
        RenderOptions o;
        // UINT64_MAX: a fresh sample index, as the reference's same-signature call draws fresh random numbers
        o.seed = seed, o.sample_offset = sample_offset == UINT64_MAX ? next_sample_index(o.spp) : sample_offset;
        Tile t{(size_t)tile[0], (size_t)tile[1], (size_t)tile[2], (size_t)tile[3]};
        AccumulationBuffer b = partial_render_scene(h->scene, t, (size_t)height, (size_t)width, o);
        size_t n = b.width() * b.height();
        if (colour) std::memcpy(colour, b.colour.data(), 3 * n * sizeof(double));
        if (colour_sum) std::memcpy(colour_sum, b.colour_sum.data(), 3 * n * sizeof(double));
        if (colour_bias) std::memcpy(colour_bias, b.colour_bias.data(), 3 * n * sizeof(double));
        if (weight) std::memcpy(weight, b.weight.data(), n * sizeof(double));
        if (weight_bias) std::memcpy(weight_bias, b.weight_bias.data(), n * sizeof(double));
    });
}
uint64_t vrjh_next_sample_index(uint64_t count) { return next_sample_index(count); }
/* render_like_main (main.rs:192-217): `calls` partial_render_scene calls from `workers` threads, merged into the caller's
 * width x height colour / weight arrays (which hold the image so far; zero weight = empty).  spp == 0: the reference's own
 * call (1 spp, limit 128, fresh samples); otherwise spp / max_depth / seed as given, samples counted from sample_offset.
 * kahan_state: whether each call also brings its Kahan arrays back (the reference's AccumulationBuffer has them; merge_tile
 * never reads them).  stats8 (9 doubles): wall_s, call_s, merge_s, device_ms, rays, calls, bytes_to_host, merge_passes, wavefront_calls. */
int vrjh_render_like_main(void *p, uint64_t width, uint64_t height, uint64_t tile_size, uint64_t calls, uint32_t workers,
                          uint32_t spp, uint32_t max_depth, uint64_t seed, uint64_t sample_offset, int kahan_state, int device,
                          double *colour, double *weight, double *stats8) { /* stats8: 9 doubles */
    HostScene *h = static_cast<HostScene *>(p);
    return guarded([&] {
        AccumulationBuffer image(width, height);
        const size_t n = width * height;
        std::memcpy(image.colour.data(), colour, 3 * n * sizeof(double));
        std::memcpy(image.weight.data(), weight, n * sizeof(double));
        RenderOptions o; // defaults = what the reference hard-codes
        if (spp) o.spp = spp, o.max_depth = max_depth, o.seed = seed, o.sample_offset = sample_offset;
        o.kahan_state = kahan_state != 0, o.device = device;
        MainLoopStats st = render_like_main(h->scene, width, height, tile_size, calls, workers, o, spp == 0, image);
        std::memcpy(colour, image.colour.data(), 3 * n * sizeof(double));
        std::memcpy(weight, image.weight.data(), n * sizeof(double));
        if (stats8) {
            stats8[0] = st.wall_s, stats8[1] = st.call_s, stats8[2] = st.merge_s, stats8[3] = st.device_ms;
            stats8[4] = (double)st.rays, stats8[5] = (double)st.calls, stats8[6] = (double)st.bytes_to_host, stats8[7] = (double)st.merge_passes, stats8[8] = (double)st.wavefront_calls;
        }
    });
}
/* AccumulationBuffer::merge_tile on raw arrays (dst is dst_w x dst_h, src is the tile's size) */
int vrjh_merge_tile(double *dst_colour, double *dst_weight, uint64_t dst_w, uint64_t dst_h, const uint64_t tile[4],
                    const double *src_colour, const double *src_weight) {
    return guarded([&] {
        Tile t{(size_t)tile[0], (size_t)tile[1], (size_t)tile[2], (size_t)tile[3]};
        AccumulationBuffer dst(dst_w, dst_h), src(t.width(), t.height());
        std::memcpy(dst.colour.data(), dst_colour, 3 * dst_w * dst_h * sizeof(double));
        std::memcpy(dst.weight.data(), dst_weight, dst_w * dst_h * sizeof(double));
        std::memcpy(src.colour.data(), src_colour, 3 * src.width() * src.height() * sizeof(double));
        std::memcpy(src.weight.data(), src_weight, src.width() * src.height() * sizeof(double));
        dst.merge_tile(t, src);
        std::memcpy(dst_colour, dst.colour.data(), 3 * dst_w * dst_h * sizeof(double));
        std::memcpy(dst_weight, dst.weight.data(), dst_w * dst_h * sizeof(double));
    });
}
/* AccumulationBuffer::merge_tiles: `n` source buffers (concatenated in src_colour / src_weight) merged in order in one pass;
 * src_weight == NULL: colour-only buffers that all carry `uniform_weight` */
int vrjh_merge_tiles(double *dst_colour, double *dst_weight, uint64_t dst_w, uint64_t dst_h, const uint64_t tile[4], uint32_t n,
                     const double *src_colour, const double *src_weight, double uniform_weight) {
    return guarded([&] {
        Tile t{(size_t)tile[0], (size_t)tile[1], (size_t)tile[2], (size_t)tile[3]};
        const size_t npix = t.width() * t.height();
        AccumulationBuffer dst(dst_w, dst_h);
        std::memcpy(dst.colour.data(), dst_colour, 3 * dst_w * dst_h * sizeof(double));
        std::memcpy(dst.weight.data(), dst_weight, dst_w * dst_h * sizeof(double));
        std::vector<AccumulationBuffer> srcs;
        for (uint32_t k = 0; k < n; k++) {
            srcs.emplace_back(t.width(), t.height());
            std::memcpy(srcs[k].colour.data(), src_colour + 3 * npix * k, 3 * npix * sizeof(double));
            if (src_weight) std::memcpy(srcs[k].weight.data(), src_weight + npix * k, npix * sizeof(double));
            else srcs[k].weight.clear(), srcs[k].uniform_weight = uniform_weight;
        }
        std::vector<const AccumulationBuffer *> ptrs;
        for (const AccumulationBuffer &b : srcs) ptrs.push_back(&b);
        dst.merge_tiles(t, ptrs);
        std::memcpy(dst_colour, dst.colour.data(), 3 * dst_w * dst_h * sizeof(double));
        std::memcpy(dst_weight, dst.weight.data(), dst_w * dst_h * sizeof(double));
    });
}
/* save_scene_cache / load_scene_cache: the flattened scene on disk.  vrjh_scene_load_cache returns a new scene handle or NULL. */
int vrjh_scene_save_cache(void *p, const char *path) {
    HostScene *h = static_cast<HostScene *>(p);
    return guarded([&] { save_scene_cache(h->scene, path); });
}
void *vrjh_scene_load_cache(const char *path) {
    HostScene *h = new HostScene();
    if (guarded([&] { h->scene = load_scene_cache(path); })) {
        delete h;
        return nullptr;
    }
    return h;
}
/* ImageRgbU8::write_png on a raw RGB array */
int vrjh_write_png(const char *path, uint64_t width, uint64_t height, const uint8_t *rgb) {
    return guarded([&] {
        ImageRgbU8 image(width, height);
        std::memcpy(image.pixel_data().data(), rgb, 3 * width * height);
        image.write_png(path);
    });
}
int64_t vrjh_tile_iterator(uint64_t width, uint64_t height, uint64_t tile_size, uint64_t *tiles, int64_t cap) {
    int64_t n = 0;
    guarded([&] {
        TileIterator it(width, height, tile_size);
        Tile t;
        while (it.next(t)) {
            if (n < cap) tiles[4 * n] = t.start_column, tiles[4 * n + 1] = t.end_column, tiles[4 * n + 2] = t.start_row, tiles[4 * n + 3] = t.end_row;
            n++;
        }
    });
    return n;
}

} // extern "C"
