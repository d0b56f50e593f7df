/*
 * vanrijn_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain C++17/f64 restatement of the
 * reference's per-pixel / per-sample render loop (partial_render_scene and
 * everything below it, /root/reference/src/camera.rs:95-130).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker or the timed CPU baseline -- never as
 * part of the product path.
 *
 * Pinning: every known-answer test the reference holds for this path
 *  is re-run against these entry points by
 * tests/test_oracle_kat.py.  The reference itself (Rust, nightly, crates.io
 * dependencies) cannot be compiled in the build container, and it has no
 * tests for the integrators, the materials, the BVH traversal or the RNG, so at
 * the integrator / material / BVH level parity is "unpinned by the reference's
 * tests": there the oracle is validated by closed-form checks
 * (tests/test_oracle_closed_form.py) and by review against the cited lines.
 *
 * The one deliberate departure from the reference: all randomness comes from a
 * counter-based generator (Philox-4x32-10) instead of rand 0.7's thread RNG
 * (SURVEY.md section 8a row 27), so a sample is a pure function of
 * (seed, pixel, sample index).
 */
#ifndef VANRIJN_ORACLE_H
#define VANRIJN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcScene OrcScene;

enum { ORC_MAT_LAMBERTIAN = 0, ORC_MAT_PHONG = 1, ORC_MAT_REFLECTIVE = 2, ORC_MAT_DIELECTRIC = 3 };
enum { ORC_INTEGRATOR_SIMPLE_RANDOM = 0, ORC_INTEGRATOR_WHITTED = 1 };
enum { ORC_TRAVERSE_REFERENCE = 0, ORC_TRAVERSE_ORDERED_PRUNED = 1 };

/* ---- scene construction (mirrors Scene / Sphere / Plane / Triangle / BVH / materials) ---- */
OrcScene *orc_scene_new(double cam_x, double cam_y, double cam_z);
void orc_scene_free(OrcScene *);
int orc_add_spectrum(OrcScene *, double shortest, double longest, int n, const double *samples);
int orc_add_spectrum_rgb(OrcScene *, double r, double g, double b); /* Spectrum::reflection_from_linear_rgb */
int orc_add_spectrum_grey(OrcScene *, double brightness);
int orc_add_spectrum_diamond(OrcScene *);
int orc_add_material(OrcScene *, int kind, int spectrum, double p0, double p1, double p2);
int orc_begin_list(OrcScene *); /* new Vec<Box<dyn Primitive>> object; returns object id */
void orc_list_add_sphere(OrcScene *, double cx, double cy, double cz, double r, int material);
void orc_list_add_plane(OrcScene *, double nx, double ny, double nz, double d, int material);
void orc_list_add_triangle(OrcScene *, const double *v9, const double *n9, int material);
/* BoundingVolumeHierarchy::build over ntri triangles (verts/normals: ntri*9 doubles) */
int orc_add_bvh(OrcScene *, int64_t ntri, const double *verts, const double *normals, int material);
/* mesh.rs load_obj + BVH build; returns object id or -1 */
int orc_add_bvh_obj(OrcScene *, const char *path, int material);
int64_t orc_bvh_triangle_count(const OrcScene *, int object_id);
int orc_bvh_depth(const OrcScene *, int object_id);
int64_t orc_bvh_leaf_order(const OrcScene *, int object_id, int32_t *prim_ids); /* DFS leaf order */
/* load_obj only: returns triangle count, fills verts/normals (caller frees with orc_free) */
int64_t orc_load_obj(const char *path, double **verts, double **normals);
void orc_free(void *);

/* ---- closest-hit queries on a supplied ray list (Sampler::sample) ---- */
typedef struct OrcTraceCounters {
    uint64_t rays, node_visits, tri_tests, hits;
} OrcTraceCounters;
/* dirs are passed through Ray::new (normalised).  object_id / prim_id = -1 on miss;
 * prim_id is the primitive's index in the order it was handed to the object. */
void orc_trace_rays(const OrcScene *, int64_t n, const double *origins, const double *dirs, int mode,
                    int32_t *object_id, int32_t *prim_id, double *t, OrcTraceCounters *counters);
/* barycentric distance to the nearest edge of the hit triangle (min b_i); 2.0 for non-triangles / misses */
void orc_trace_rays_edge_distance(const OrcScene *, int64_t n, const double *origins, const double *dirs,
                                  double *min_bary);

/* ---- render ---- */
typedef struct OrcLight {
    double direction[3];
    int32_t spectrum;
    int32_t pad;
} OrcLight;

typedef struct OrcRenderParams {
    uint32_t spp;
    uint32_t max_depth;  /* RECURSION_LIMIT; reference value 128 */
    uint64_t sample_offset;
    uint64_t seed;
    uint32_t integrator; /* ORC_INTEGRATOR_* */
    uint32_t traverse;   /* ORC_TRAVERSE_* */
    double bias;         /* reference value 1e-7 */
    const OrcLight *lights;
    uint32_t n_lights;
    int32_t ambient_spectrum; /* Whitted only; -1 = black */
    uint32_t threads;         /* 0 = all cores */
    uint32_t pad;
} OrcRenderParams;

typedef struct OrcRenderStats {
    uint64_t primary_rays, bounce_rays, shadow_rays;
    uint64_t node_visits, tri_tests;
    uint64_t paths_missed, paths_escaped, paths_depth_limited;
} OrcRenderStats;

/* tile = {start_column, end_column, start_row, end_row}.  Outputs are tile-local, row-major
 * (tile.height rows of tile.width): colour_sum 3 doubles per pixel, weight 1 per pixel;
 * photons (optional, may be NULL): spp * npix * 2 doubles (wavelength, intensity*360) indexed
 * [(s * npix + pixel) * 2]. */
void orc_render_tile(const OrcScene *, const uint64_t tile[4], uint64_t height, uint64_t width,
                     const OrcRenderParams *, double *colour_sum, double *colour_bias, double *weight,
                     double *weight_bias, double *colour, double *photons, OrcRenderStats *stats);

/* ---- unit-level entry points for the reference's known-answer tests ---- */
/* Triangle::intersect (triangle.rs:35-97). out: distance, location[3], normal[3], tangent[3], cotangent[3], retro[3] */
int orc_triangle_intersect(const double *v9, const double *n9, const double *origin, const double *dir, double *out16);
int orc_sphere_intersect(const double *centre, double radius, const double *origin, const double *dir, double *out16);
int orc_plane_intersect(const double *normal, double dist, const double *origin, const double *dir, double *out16);
int orc_aabb_intersect(const double *lo, const double *hi, const double *origin, const double *dir);
void orc_triangle_helpers(const double *dir, int *perm3, double *shear2);
double orc_spectrum_intensity(double lo, double hi, int n, const double *samples, double wavelength);
void orc_rgb_to_spectrum(double r, double g, double b, double *samples32);
void orc_cmf_xyz(double wavelength, double *xyz);
void orc_xyz_to_linear_rgb(const double *xyz, double *rgb);
void orc_linear_rgb_to_xyz(const double *rgb, double *xyz);
double orc_srgb_gamma(double u);
/* ClampingToneMapper (image.rs:130-187): source 0 = XYZ -> sRGB -> clamp -> u8, 1 = linear RGB -> clamp -> u8 */
void orc_tone_map(int source, const double *colour, int64_t n, uint8_t *rgb8);
/* AccumulationBuffer::update_pixel on one pixel: state = {colour[3], sum[3], bias[3], weight, weight_bias} */
void orc_accum_update(double *state11, double wavelength, double intensity, double weight);
void orc_accum_blend(const double *c1, double w1, const double *c2, double w2, double *out3);
void orc_camera_ray(uint64_t width, uint64_t height, const double *cam, uint64_t row, uint64_t col, double ux,
                    double uy, double *origin3, double *dir3);
int orc_mat3_inverse(const double *m9, double *out9); /* try_inverse: cofactor^T * det (sic) */
double orc_mat3_determinant(const double *m9);
int orc_largest_dimension(const double *lo, const double *hi);
int64_t orc_tile_iterator(uint64_t width, uint64_t height, uint64_t tile_size, uint64_t *tiles, int64_t cap);
/* materials: sample() with draws taken from the counter RNG at (seed,pixel,sample,first_ordinal) */
void orc_material_sample(const OrcScene *, int material, const double *w_i, double wavelength, uint64_t seed,
                         uint32_t pixel, uint64_t sample, uint32_t first_ordinal, double *dir3, double *pdf,
                         uint32_t *draws_used);
double orc_material_bsdf(const OrcScene *, int material, const double *w_o, const double *w_i, double wavelength,
                         double intensity_in);
double orc_sky(const double *w, double wavelength);
/* RNG */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_rng_f64(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal);
double orc_rng_open01(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal);
int orc_rng_bool(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal);

#ifdef __cplusplus
}
#endif
#endif
