/*
 * vanrijn_oracle.cpp -- CPU oracle: a plain C++17 / IEEE-binary64 restatement of the
 * reference's hot path.  TEST INFRASTRUCTURE ONLY (see vanrijn_oracle.h).
 *
 * Build with -ffp-contract=off: rustc never fuses a*b+c, GCC does by default.
 * Every function names the reference lines it follows (paths relative to
 * /root/reference/).  Operation order follows the reference so that results
 * are bit-comparable where only + - * / sqrt are involved.
 */
#include "vanrijn_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kInf = std::numeric_limits<double>::infinity();
const double kPi = 3.14159265358979323846264338327950288; /* std::f64::consts::PI */

/* ------------------------------------------------------------------ math/vec3.rs */
struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
/* vec3.rs:76-82 -- Iterator::sum over the products, folding from 0.0 in x,y,z order */
inline double dot(V3 a, V3 b) { return ((0.0 + a.x * b.x) + a.y * b.y) + a.z * b.z; }
/* vec3.rs:84-89 */
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline double norm_squared(V3 a) { return dot(a, a); }
inline double norm(V3 a) { return std::sqrt(norm_squared(a)); }
/* vec3.rs:103-110 -- multiply by the reciprocal of the norm */
inline V3 normalize(V3 a) {
    double inv = 1.0 / norm(a);
    return v3(a.x * inv, a.y * inv, a.z * inv);
}
inline V3 component_mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
/* vec3.rs:112-127 */
inline int smallest_coord(V3 a) {
    double x = std::fabs(a.x), y = std::fabs(a.y), z = std::fabs(a.z);
    if (x < y) return x < z ? 0 : 2;
    return y < z ? 1 : 2;
}

/* ------------------------------------------------------------------ math/mat3.rs, mat2.rs */
struct M3 {
    double e[3][3];
};
/* mat3.rs:34-42 */
inline M3 from_rows(V3 r0, V3 r1, V3 r2) {
    M3 m = {{{r0.x, r0.y, r0.z}, {r1.x, r1.y, r1.z}, {r2.x, r2.y, r2.z}}};
    return m;
}
/* mat3.rs:72-90 + mat2.rs:13-15 */
inline double first_minor(const M3 &m, int row, int col) {
    double s[2][2];
    int id = 0;
    for (int i = 0; i < 3; i++) {
        if (i == row) continue;
        int jd = 0;
        for (int j = 0; j < 3; j++) {
            if (j == col) continue;
            s[id][jd++] = m.e[i][j];
        }
        id++;
    }
    return s[0][0] * s[1][1] - s[0][1] * s[1][0];
}
/* mat3.rs:92-94 */
inline double cofactor(const M3 &m, int row, int col) {
    double sign = ((row + col) & 1) ? -1.0 : 1.0;
    return sign * first_minor(m, row, col);
}
/* mat3.rs:106-109 */
inline double determinant(const M3 &m) {
    return m.e[0][0] * first_minor(m, 0, 0) - m.e[0][1] * first_minor(m, 0, 1) + m.e[0][2] * first_minor(m, 0, 2);
}
/* mat3.rs:111-118 -- transpose(cofactors) TIMES the determinant (sic) */
inline bool try_inverse(const M3 &m, M3 *out) {
    double det = determinant(m);
    if (det == 0.0) return false;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out->e[i][j] = cofactor(m, j, i) * det;
    return true;
}
/* mat3.rs:147-157 */
inline V3 mul(const M3 &m, V3 v) {
    return v3(dot(v3(m.e[0][0], m.e[0][1], m.e[0][2]), v), dot(v3(m.e[1][0], m.e[1][1], m.e[1][2]), v),
              dot(v3(m.e[2][0], m.e[2][1], m.e[2][2]), v));
}

/* ------------------------------------------------------------------ RNG (departure from rand 0.7) */
inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
}
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        if (r) k[0] += 0x9E3779B9u, k[1] += 0xBB67AE85u;
        philox_round(c, k);
    }
    for (int i = 0; i < 4; i++) out[i] = c[i];
}
/* One stream per (seed, pixel, sample); `ordinal` counts draws along the path:
 * 0 = camera x, 1 = camera y, 2 = wavelength, (3 unused,) 4.. = material sampling in call order. */
struct Rng {
    uint64_t seed;
    uint32_t pixel;
    uint64_t sample;
    uint32_t ordinal;
    uint64_t bits() {
        uint32_t ctr[4] = {ordinal >> 1, pixel, (uint32_t)sample, (uint32_t)(sample >> 32)};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t w[4];
        philox4x32_10(ctr, key, w);
        int h = ordinal & 1;
        ordinal++;
        return ((uint64_t)w[2 * h + 1] << 32) | w[2 * h];
    }
    /* rand 0.7 Standard f64: 53 random bits -> [0,1) */
    double f64() { return (double)(bits() >> 11) * (1.0 / 9007199254740992.0); }
    /* rand 0.7 Open01: 52 random bits -> (0,1) */
    double open01() { return ((double)(bits() >> 12) + 0.5) * (1.0 / 4503599627370496.0); }
    /* rand 0.7 Standard bool: sign bit */
    bool boolean() { return (bits() >> 63) != 0; }
};

/* ------------------------------------------------------------------ colour/spectrum.rs */
const double kRgbBasis[7][32] = {
#include "rgb_basis_tables.inc"
};
enum { B_WHITE = 0, B_CYAN, B_MAGENTA, B_YELLOW, B_RED, B_GREEN, B_BLUE };

struct Spectrum {
    double shortest, longest;
    std::vector<double> samples;
    /* spectrum.rs:50-79 */
    double intensity_at(double wavelength) const {
        if (wavelength < shortest || wavelength > longest) return 0.0;
        size_t n = samples.size();
        double range = longest - shortest;
        double fidx = (double)(n - 1) * ((wavelength - shortest) / range);
        size_t before = (fidx != fidx || fidx < 0.0) ? 0 : (size_t)fidx; /* `as usize` saturates; NaN -> 0 */
        double wl_before = (double)before / (double)(n - 1) * range + shortest;
        if (before == n - 1) return samples[before];
        double wl_after = (double)(before + 1) / (double)(n - 1) * range + shortest;
        double delta = wl_after - wl_before;
        double ratio = (wavelength - wl_before) / delta;
        return samples[before] * (1.0 - ratio) + samples[before + 1] * ratio;
    }
};
/* spectrum.rs:81-165 -- six-way decomposition on the channel ordering */
Spectrum reflection_from_linear_rgb(double r, double g, double b) {
    Spectrum s;
    s.shortest = 380.0, s.longest = 720.0; /* spectrum.rs:179-180 */
    s.samples.resize(32);
    int second, third;
    double c0, c1, c2;
    if (r <= g && r <= b) {
        if (g <= b) { second = B_CYAN, third = B_BLUE, c0 = r, c1 = g - r, c2 = b - g; }
        else        { second = B_CYAN, third = B_GREEN, c0 = r, c1 = b - r, c2 = g - b; }
    } else if (g <= r && g < b) {
        if (r <= b) { second = B_MAGENTA, third = B_BLUE, c0 = g, c1 = r - g, c2 = b - r; }
        else        { second = B_MAGENTA, third = B_RED, c0 = g, c1 = b - g, c2 = r - b; }
    } else {
        if (r <= g) { second = B_YELLOW, third = B_GREEN, c0 = b, c1 = r - b, c2 = g - r; }
        else        { second = B_YELLOW, third = B_RED, c0 = b, c1 = g - b, c2 = r - g; }
    }
    for (int i = 0; i < 32; i++)
        s.samples[i] = c0 * kRgbBasis[B_WHITE][i] + c1 * kRgbBasis[second][i] + c2 * kRgbBasis[third][i];
    return s;
}
/* spectrum.rs:29-48 */
Spectrum diamond_index_of_refraction() {
    Spectrum s;
    s.shortest = 326.27, s.longest = 774.9;
    s.samples = {2.505813241, 2.487866556, 2.473323675, 2.464986815, 2.455051934, 2.441251728,
                 2.431478974, 2.427076431, 2.420857286, 2.411429037, 2.406543164, 2.406202402};
    return s;
}
/* spectrum.rs:20-26 */
Spectrum grey(double brightness) {
    Spectrum s;
    s.shortest = 380.0, s.longest = 740.0;
    s.samples = {brightness, brightness};
    return s;
}

/* ------------------------------------------------------------------ colour/photon.rs, colour_xyz.rs */
struct Photon {
    double wavelength, intensity;
};
/* colour_xyz.rs:86-89 */
inline double gaussian(double w, double alpha, double mu, double s1, double s2) {
    double sigma = w < mu ? s1 : s2;
    double denominator = 2.0 * (sigma * sigma);
    return alpha * std::exp(-((w - mu) * (w - mu)) / denominator);
}
/* colour_xyz.rs:91-103 */
inline V3 cmf(double w) {
    double x = gaussian(w, 1.056, 599.8, 37.9, 31.0) + gaussian(w, 0.362, 442.0, 16.0, 26.7) +
               gaussian(w, -0.065, 501.1, 20.4, 26.2);
    double y = gaussian(w, 0.821, 568.8, 46.9, 40.5) + gaussian(w, 0.286, 530.9, 16.3, 31.1);
    double z = gaussian(w, 1.217, 437.0, 11.8, 36.0) + gaussian(w, 0.681, 459.0, 26.0, 13.8);
    return v3(x, y, z);
}
/* colour_xyz.rs:31-35 */
inline V3 xyz_from_photon(Photon p) { return cmf(p.wavelength) * p.intensity; }

/* ------------------------------------------------------------------ materials */
struct Material {
    int kind;
    int spectrum;
    double p0, p1, p2; /* Lambertian: diffuse. Phong: diffuse, specular, smoothness. Reflective: diffuse, reflection */
};
struct SampleResult {
    V3 direction;
    double pdf;
};

/* smooth_transparent_dialectric.rs:15-60 */
struct Fresnel {
    V3 reflection_direction;
    double reflection_strength;
    V3 transmission_direction;
    double transmission_strength;
};
Fresnel fresnel(V3 w_i, double eta1, double eta2) {
    V3 normal = w_i.z > 0.0 ? v3(0, 0, 1) : -v3(0, 0, 1);
    Fresnel f;
    f.reflection_direction = v3(-w_i.x, -w_i.y, w_i.z);
    double r = eta1 / eta2;
    double cos1 = dot(normal, w_i);
    double cos2sq = 1.0 - r * r * (1.0 - cos1 * cos1);
    if (cos2sq >= 0.0) {
        double cos2 = std::sqrt(cos2sq);
        double rpar = (eta1 * cos2 - eta2 * cos1) / (eta1 * cos2 + eta2 * cos1);
        double rperp = (eta1 * cos1 - eta2 * cos2) / (eta1 * cos1 + eta2 * cos2);
        f.reflection_strength = 0.5 * (rpar * rpar + rperp * rperp);
        f.transmission_direction = normalize((w_i * (-r)) + (normal * (r * cos1 - cos2)));
        f.transmission_strength = 1.0 - f.reflection_strength;
    } else {
        f.reflection_strength = 1.0;
        f.transmission_strength = 0.0;
        f.transmission_direction = v3(0, 0, 0);
    }
    if (w_i.z < 0.0) {
        f.reflection_direction.z *= -1.0;
        f.transmission_direction.z *= -1.0;
    }
    return f;
}

/* ------------------------------------------------------------------ raycasting */
struct Ray {
    V3 origin, direction;
};
/* raycasting/mod.rs:41-46 -- always normalises */
inline Ray ray_new(V3 o, V3 d) { return Ray{o, normalize(d)}; }
/* raycasting/mod.rs:49-51 */
inline V3 point_at(const Ray &r, double t) { return r.origin + r.direction * t; }
/* raycasting/mod.rs:58-60 */
inline Ray ray_bias(const Ray &r, double amount) { return ray_new(r.origin + r.direction * amount, r.direction); }

/* raycasting/mod.rs:67-97 (+ ids, which the reference does not carry) */
struct Hit {
    double distance;
    V3 location, normal, tangent, cotangent, retro;
    int material;
    int object_id, prim_id;
    double min_bary;
};

/* util/interval.rs:7-62 */
struct Interval {
    double lo, hi;
};
inline Interval interval_new(double a, double b) { return a > b ? Interval{b, a} : Interval{a, b}; }
inline double rust_max(double a, double b) { return std::fmax(a, b); } /* f64::max ignores NaN, like fmax */
inline double rust_min(double a, double b) { return std::fmin(a, b); }
inline Interval interval_intersection(Interval a, Interval b) { return Interval{rust_max(a.lo, b.lo), rust_min(a.hi, b.hi)}; }
inline bool interval_is_empty(Interval a) { return a.lo > a.hi; }
inline Interval interval_union(Interval a, Interval b) {
    if (interval_is_empty(a)) return b;
    if (interval_is_empty(b)) return a;
    return Interval{rust_min(a.lo, b.lo), rust_max(a.hi, b.hi)};
}
inline Interval interval_expand(Interval a, double v) {
    if (interval_is_empty(a)) return Interval{v, v};
    return Interval{rust_min(a.lo, v), rust_max(a.hi, v)};
}

/* util/axis_aligned_bounding_box.rs:6-100 */
struct Box {
    Interval b[3];
};
inline Box box_empty() { return Box{{{kInf, -kInf}, {kInf, -kInf}, {kInf, -kInf}}}; }
inline Box box_expand(Box a, V3 p) {
    return Box{{interval_expand(a.b[0], p.x), interval_expand(a.b[1], p.y), interval_expand(a.b[2], p.z)}};
}
inline Box box_union(Box a, Box o) {
    return Box{{interval_union(a.b[0], o.b[0]), interval_union(a.b[1], o.b[1]), interval_union(a.b[2], o.b[2])}};
}
/* util/axis_aligned_bounding_box.rs:76-99 -- first strictly largest; degenerate dims count as -1 */
inline int largest_dimension(const Box &bx) {
    int dim = 0;
    double size = 0.0;
    for (int i = 0; i < 3; i++) {
        double s = (bx.b[i].lo == bx.b[i].hi) ? -1.0 : bx.b[i].hi - bx.b[i].lo;
        if (s > size) dim = i, size = s;
    }
    return dim;
}
/* raycasting/axis_aligned_bounding_box.rs:9-27 -- a LINE test: the running interval starts at (-inf,+inf) */
inline bool box_intersect(const Box &bx, const Ray &r) {
    Interval t = {-kInf, kInf};
    for (int i = 0; i < 3; i++) {
        double o = r.origin[i], d = r.direction[i];
        t = interval_intersection(t, interval_new((bx.b[i].lo - o) / d, (bx.b[i].hi - o) / d));
        if (interval_is_empty(t)) return false;
    }
    return true;
}

struct Triangle {
    V3 v[3], n[3];
    int material;
    int prim_id;
};
inline Box triangle_box(const Triangle &t) { return box_expand(box_expand(box_expand(box_empty(), t.v[0]), t.v[1]), t.v[2]); }

/* triangle.rs:108-122 -- SIGNED largest component goes last; cyclic permutations */
inline void permutation_largest_last(V3 d, int idx[3]) {
    if (d.x > d.y) {
        if (d.z > d.x) idx[0] = 0, idx[1] = 1, idx[2] = 2;
        else idx[0] = 1, idx[1] = 2, idx[2] = 0;
    } else {
        if (d.z > d.y) idx[0] = 0, idx[1] = 1, idx[2] = 2;
        else idx[0] = 2, idx[1] = 0, idx[2] = 1;
    }
}
inline V3 permute(V3 a, const int idx[3]) { return v3(a[idx[0]], a[idx[1]], a[idx[2]]); }
/* triangle.rs:141-143 */
inline double edge_fn(V3 a, V3 b) { return a.x * b.y - b.x * a.y; }

/* triangle.rs:35-97 */
bool triangle_intersect(const Triangle &tri, const Ray &ray, Hit *out) {
    V3 translation = -ray.origin;
    int idx[3];
    permutation_largest_last(ray.direction, idx);
    V3 pd = permute(ray.direction, idx);
    double sx = -pd.x / pd.z, sy = -pd.y / pd.z; /* :133-135 */
    V3 tv[3];
    for (int i = 0; i < 3; i++) {
        V3 p = permute(tri.v[i] + translation, idx);
        tv[i] = v3(p.x + sx * p.z, p.y + sy * p.z, p.z); /* :137-139 */
    }
    double e[3] = {edge_fn(tv[1], tv[2]), edge_fn(tv[2], tv[0]), edge_fn(tv[0], tv[1])}; /* :145-158 */
    bool all_pos = !std::signbit(e[0]) && !std::signbit(e[1]) && !std::signbit(e[2]);
    bool all_neg = std::signbit(e[0]) && std::signbit(e[1]) && std::signbit(e[2]);
    if (!(all_pos || all_neg)) return false;
    double ae[3] = {std::fabs(e[0]), std::fabs(e[1]), std::fabs(e[2])};
    double inv_sum = 1.0 / (((0.0 + ae[0]) + ae[1]) + ae[2]); /* :160-162 */
    double b[3] = {ae[0] * inv_sum, ae[1] * inv_sum, ae[2] * inv_sum};
    double tz = ((0.0 + tv[0].z * b[0]) + tv[1].z * b[1]) + tv[2].z * b[2]; /* :57-62 */
    if ((!std::signbit(tz)) != (!std::signbit(pd.z))) return false;            /* :63-65 */
    V3 location = ((v3(0, 0, 0) + tri.v[0] * b[0]) + tri.v[1] * b[1]) + tri.v[2] * b[2];
    out->distance = norm(ray.origin - location);
    out->location = location;
    out->normal = normalize(((v3(0, 0, 0) + tri.n[0] * b[0]) + tri.n[1] * b[1]) + tri.n[2] * b[2]);
    out->cotangent = normalize(cross(tri.v[0] - tri.v[1], out->normal));
    out->tangent = normalize(cross(out->cotangent, out->normal));
    out->retro = normalize(ray.origin - location);
    out->material = tri.material;
    out->prim_id = tri.prim_id;
    out->min_bary = std::min(b[0], std::min(b[1], b[2]));
    return true;
}

struct Sphere {
    V3 centre;
    double radius;
    int material;
};
/* sphere.rs:39-93 */
bool sphere_intersect(const Sphere &s, const Ray &ray, Hit *out) {
    V3 o = ray.origin, c = s.centre, d = ray.direction;
    V3 dd = component_mul(d, d);
    double a = ((0.0 + dd.x) + dd.y) + dd.z;
    V3 bv = (component_mul(o, d) - component_mul(c, d)) * 2.0;
    double b = ((0.0 + bv.x) + bv.y) + bv.z;
    V3 cv = (component_mul(o, o) + component_mul(c, c)) - component_mul(c, o) * 2.0;
    double cc = (((0.0 + cv.x) + cv.y) + cv.z) - s.radius * s.radius;
    double delta_squared = b * b - 4.0 * a * cc;
    if (delta_squared < 0.0) return false;
    double delta = std::sqrt(delta_squared);
    double one_over_2a = 1.0 / (2.0 * a);
    double t1 = (-b - delta) * one_over_2a;
    double t2 = (-b + delta) * one_over_2a;
    double distance = (t1 < 0.0 || (t2 >= 0.0 && t1 >= t2)) ? t2 : t1;
    if (distance <= 0.0) return false;
    out->distance = distance;
    out->location = point_at(ray, distance);
    out->normal = normalize(out->location - s.centre);
    out->tangent = normalize(cross(out->normal, v3(0, 0, 1)));
    out->cotangent = cross(out->normal, out->tangent);
    out->retro = -ray.direction;
    out->material = s.material;
    out->min_bary = 2.0;
    return true;
}

struct Plane {
    V3 normal, tangent, cotangent;
    double distance_from_origin;
    int material;
};
/* plane.rs:17-32 */
Plane plane_new(V3 normal, double d, int material) {
    Plane p;
    p.normal = normalize(normal);
    double axis[3] = {0, 0, 0};
    axis[smallest_coord(p.normal)] = 1.0;
    p.cotangent = normalize(cross(p.normal, v3(axis[0], axis[1], axis[2])));
    p.tangent = cross(p.normal, p.cotangent);
    p.distance_from_origin = d;
    p.material = material;
    return p;
}
/* plane.rs:48-75 */
bool plane_intersect(const Plane &p, const Ray &ray, Hit *out) {
    double d_dot_n = dot(ray.direction, p.normal);
    V3 point_on_plane = p.normal * p.distance_from_origin;
    double num = dot(point_on_plane - ray.origin, p.normal);
    if (d_dot_n == 0.0) {
        if (num != 0.0) return false;
    }
    double t = num / d_dot_n;
    if (t < 0.0) return false;
    out->distance = t;
    out->location = point_at(ray, t);
    out->normal = p.normal;
    out->tangent = p.tangent;
    out->cotangent = p.cotangent;
    out->retro = -ray.direction;
    out->material = p.material;
    out->min_bary = 2.0;
    return true;
}

/* ------------------------------------------------------------------ bounding_volume_hierarchy.rs */
struct Counters {
    uint64_t node_visits = 0, tri_tests = 0;
};

struct Bvh {
    struct Node {
        Box bounds;
        int32_t left = -1, right = -1; /* internal */
        int64_t first = 0, count = 0;  /* leaf */
        bool leaf = false;
    };
    std::vector<Node> nodes;
    std::vector<Triangle> tris; /* reordered in place by the build, as the reference's slice is */
    int depth = 0;

    static double centre_on(const Triangle &t, int axis) {
        Box b = triangle_box(t);
        return (b.b[axis].lo + b.b[axis].hi) / 2.0; /* :30-36 */
    }
    /* :49-75.  sort_unstable_by in the reference: order among equal keys is unspecified there;
     * a stable sort is used here (and in the product's host builder) so both trees agree.
     *
     * The recursion is the reference's (bounds -> largest_dimension -> sort the slice by box centre on that axis ->
     * split at len/2 -> recurse; a slice of <= 1 primitive is a leaf).  Two things are arranged for speed only, so the
     * 10 M-triangle tree of config C4 can be built inside a test: (i) the slice is ordered by sorting (key, position)
     * pairs -- the key computed once per element, with the comparator's own arithmetic -- and then moving the
     * triangles, which is what a stable sort of the triangles themselves with that comparator produces; (ii) node
     * indices are assigned up front (pre-order: a node, then its whole left subtree of 2 * len_left - 1 nodes, then
     * the right one -- the order the sequential recursion creates them in), so the two halves can be built by
     * different threads. */
    void build_all() {
        const int64_t n = (int64_t)tris.size();
        nodes.assign((size_t)std::max<int64_t>(1, 2 * n - 1), Node());
        int d = 0;
#pragma omp parallel
#pragma omp single nowait
        d = build(0, n, 0, 0);
        depth = d;
    }
    int build(int64_t begin, int64_t end, int level, int32_t me) {
        Box bounds = box_empty();
        for (int64_t i = begin; i < end; i++) bounds = box_union(bounds, triangle_box(tris[i]));
        nodes[me].bounds = bounds;
        if (end - begin <= 1) {
            nodes[me].leaf = true, nodes[me].first = begin, nodes[me].count = end - begin;
            return level + 1;
        }
        int axis = largest_dimension(bounds); /* :38-46 */
        const int64_t len = end - begin;
        {
            std::vector<std::pair<double, int64_t>> keyed((size_t)len);
            for (int64_t i = 0; i < len; i++) keyed[(size_t)i] = {centre_on(tris[begin + i], axis), i};
            std::stable_sort(keyed.begin(), keyed.end(),
                             [](const std::pair<double, int64_t> &a, const std::pair<double, int64_t> &b) { return a.first < b.first; });
            std::vector<Triangle> moved((size_t)len);
            for (int64_t i = 0; i < len; i++) moved[(size_t)i] = tris[begin + keyed[(size_t)i].second];
            std::copy(moved.begin(), moved.end(), tris.begin() + begin);
        }
        int64_t pivot = begin + len / 2;
        const int32_t l = me + 1, r = me + (int32_t)(2 * (pivot - begin)); /* left subtree: 2 * (pivot - begin) - 1 nodes */
        int dl = 0, dr = 0;
#pragma omp task shared(dl) if (len > 8192)
        dl = build(begin, pivot, level + 1, l);
        dr = build(pivot, end, level + 1, r);
#pragma omp taskwait
        nodes[me].left = l, nodes[me].right = r;
        return std::max(dl, dr);
    }

    /* :77-92 -- a.distance < b.distance ? a : b (ties and NaN pick b) */
    static bool closest(bool has_a, const Hit &a, bool has_b, const Hit &b, Hit *out) {
        if (!has_a) {
            if (has_b) *out = b;
            return has_b;
        }
        if (!has_b) {
            *out = a;
            return true;
        }
        *out = (a.distance < b.distance) ? a : b;
        return true;
    }
    /* :94-120 -- visit BOTH children whenever the node's box is hit; no ordering, no t_max */
    bool intersect_reference(int32_t ni, const Ray &ray, Hit *out, Counters *c) const {
        const Node &n = nodes[ni];
        c->node_visits++;
        if (!box_intersect(n.bounds, ray)) return false;
        if (n.leaf) {
            bool has = false;
            Hit acc{};
            for (int64_t i = n.first; i < n.first + n.count; i++) {
                Hit h{};
                c->tri_tests++;
                bool hb = triangle_intersect(tris[i], ray, &h);
                Hit merged{};
                has = closest(has, acc, hb, h, &merged);
                if (has) acc = merged;
            }
            if (has) *out = acc;
            return has;
        }
        Hit a{}, b{};
        bool ha = intersect_reference(n.left, ray, &a, c);
        bool hb = intersect_reference(n.right, ray, &b, c);
        return closest(ha, a, hb, b, out);
    }

    /* Ordered + t_max-pruned walk of the SAME tree (counts V and T for the roofline's
     * algorithmic bytes; cross-checked against intersect_reference in tests).
     * Ties keep the reference's rule: the later leaf in DFS order wins. */
    static bool slab(const Box &bx, const Ray &r, double *t_enter, double *t_exit) {
        double lo = -kInf, hi = kInf;
        for (int i = 0; i < 3; i++) {
            Interval s = interval_new((bx.b[i].lo - r.origin[i]) / r.direction[i], (bx.b[i].hi - r.origin[i]) / r.direction[i]);
            lo = rust_max(lo, s.lo), hi = rust_min(hi, s.hi);
        }
        *t_enter = lo, *t_exit = hi;
        return !(lo > hi);
    }
    bool intersect_ordered(const Ray &ray, Hit *out, Counters *c) const {
        if (nodes.empty()) return false;
        struct Entry {
            int32_t node;
            double t_enter;
        };
        Entry stack[128];
        int sp = 0;
        bool has = false;
        Hit best{};
        int64_t best_order = -1;
        double te, tx;
        c->node_visits++;
        if (!slab(nodes[0].bounds, ray, &te, &tx) || tx < 0.0) return false;
        stack[sp++] = {0, te};
        while (sp) {
            Entry e = stack[--sp];
            if (has && e.t_enter > best.distance * (1.0 + 1e-9)) continue;
            const Node &n = nodes[e.node];
            if (n.leaf) {
                for (int64_t i = n.first; i < n.first + n.count; i++) {
                    Hit h{};
                    c->tri_tests++;
                    if (triangle_intersect(tris[i], ray, &h)) {
                        if (!has || h.distance < best.distance || (h.distance == best.distance && i > best_order))
                            best = h, best_order = i, has = true;
                    }
                }
                continue;
            }
            double tl, txl, tr, txr;
            c->node_visits += 2;
            bool hl = slab(nodes[n.left].bounds, ray, &tl, &txl) && !(txl < 0.0);
            bool hr = slab(nodes[n.right].bounds, ray, &tr, &txr) && !(txr < 0.0);
            if (hl && hr) {
                if (tl <= tr) stack[sp++] = {n.right, tr}, stack[sp++] = {n.left, tl};
                else stack[sp++] = {n.left, tl}, stack[sp++] = {n.right, tr};
            } else if (hl) stack[sp++] = {n.left, tl};
            else if (hr) stack[sp++] = {n.right, tr};
        }
        if (has) *out = best;
        return has;
    }
};

/* ------------------------------------------------------------------ scene.rs, sampler.rs, vec_aggregate.rs */
struct ListPrim {
    int kind; /* 0 sphere, 1 plane, 2 triangle */
    Sphere s;
    Plane p;
    Triangle t;
};
struct Object {
    bool is_bvh = false;
    std::vector<ListPrim> prims;
    std::unique_ptr<Bvh> bvh;
};

} // namespace

struct OrcScene {
    V3 camera;
    std::vector<Spectrum> spectra;
    std::vector<Material> materials;
    std::vector<Object> objects;
};

namespace {

/* Iterator::min_by as used in sampler.rs:14-19 and vec_aggregate.rs:13-21:
 * the new element replaces the kept one only when kept > new; NaN compares as "Less" -> keep. */
inline bool min_by_replaces(double kept, double candidate) { return kept > candidate; }

bool list_intersect(const Object &obj, const Ray &ray, Hit *out, Counters *) {
    bool has = false;
    Hit best{};
    for (size_t i = 0; i < obj.prims.size(); i++) {
        const ListPrim &lp = obj.prims[i];
        Hit h{};
        bool hit = lp.kind == 0 ? sphere_intersect(lp.s, ray, &h)
                 : lp.kind == 1 ? plane_intersect(lp.p, ray, &h)
                                : triangle_intersect(lp.t, ray, &h);
        if (!hit) continue;
        h.prim_id = (int)i;
        if (!has || min_by_replaces(best.distance, h.distance)) best = h, has = true;
    }
    if (has) *out = best;
    return has;
}

/* sampler.rs:9-20 */
bool scene_sample(const OrcScene &sc, const Ray &ray, int mode, Hit *out, Counters *c) {
    bool has = false;
    Hit best{};
    for (size_t oi = 0; oi < sc.objects.size(); oi++) {
        const Object &obj = sc.objects[oi];
        Hit h{};
        bool hit;
        if (obj.is_bvh) {
            hit = mode == ORC_TRAVERSE_REFERENCE ? (!obj.bvh->nodes.empty() && obj.bvh->intersect_reference(0, ray, &h, c))
                                                 : obj.bvh->intersect_ordered(ray, &h, c);
        } else {
            hit = list_intersect(obj, ray, &h, c);
        }
        if (!hit) continue;
        h.object_id = (int)oi;
        if (!has || min_by_replaces(best.distance, h.distance)) best = h, has = true;
    }
    if (has) *out = best;
    return has;
}

/* ------------------------------------------------------------------ material sample / bsdf */
/* materials/mod.rs:28-33 + cosine_weighted_hemisphere.rs:19-33 + unit_disc.rs:27-44 + uniform_square.rs:20-25 */
SampleResult sample_default(Rng &rng) {
    double sx = -1.0 + rng.open01() * 2.0; /* corner + Vec2(u0,u1)*size */
    double sy = -1.0 + rng.open01() * 2.0;
    double dx, dy;
    if (sx == 0.0 && sy == 0.0) {
        dx = sx, dy = sy;
    } else {
        double radius, angle;
        if (std::fabs(sx) > std::fabs(sy)) radius = sx, angle = (kPi / 4.0) * sy / sx;
        else radius = sy, angle = kPi / 2.0 - (kPi / 4.0) * sx / sy;
        dx = std::cos(angle) * radius, dy = std::sin(angle) * radius;
    }
    double z = std::sqrt(rust_max(0.0, 1.0 - dx * dx - dy * dy));
    SampleResult r;
    r.direction = v3(dx, dy, z);
    r.pdf = std::sqrt(dx * dx + dy * dy) / kPi;
    return r;
}
/* lambertian_material.rs:36-59 */
SampleResult sample_lambertian(Rng &rng) {
    double x = 2.0 * rng.open01() - 1.0;
    double y = 2.0 * rng.open01() - 1.0;
    while (norm_squared(v3(x, y, 0.0)) > 1.0) {
        x = 2.0 * rng.open01() - 1.0;
        y = 2.0 * rng.open01() - 1.0;
    }
    double z = rust_max(std::sqrt(1.0 - x * x - y * y), 0.0);
    V3 w = v3(x, y, z);
    double cos_theta = dot(w, v3(0, 0, 1));
    double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    SampleResult r;
    r.direction = normalize(w);
    r.pdf = (cos_theta * sin_theta) / kPi;
    return r;
}
SampleResult material_sample(const OrcScene &sc, const Material &m, V3 w_i, double wavelength, Rng &rng) {
    switch (m.kind) {
    case ORC_MAT_LAMBERTIAN: return sample_lambertian(rng);
    case ORC_MAT_REFLECTIVE: return SampleResult{v3(-w_i.x, -w_i.y, w_i.z), 1.0}; /* reflective_material.rs:42-47 */
    case ORC_MAT_DIELECTRIC: { /* smooth_transparent_dialectric.rs:91-114 */
        double eta = sc.spectra[m.spectrum].intensity_at(wavelength);
        double eta1 = w_i.z >= 0.0 ? 1.0 : eta, eta2 = w_i.z >= 0.0 ? eta : 1.0;
        Fresnel f = fresnel(w_i, eta1, eta2);
        if (f.transmission_strength <= 0.0000000001) return SampleResult{f.reflection_direction, 0.5};
        if (f.reflection_strength <= 0.0000000001 || rng.boolean()) return SampleResult{f.transmission_direction, 0.5};
        return SampleResult{f.reflection_direction, 0.5};
    }
    default: return sample_default(rng); /* Phong uses the trait default */
    }
}
/* bsdf(w_o, w_i, photon_in).intensity ; the wavelength passes through unchanged */
double material_bsdf(const OrcScene &sc, const Material &m, V3 w_o, V3 w_i, double wavelength, double in) {
    switch (m.kind) {
    case ORC_MAT_LAMBERTIAN: { /* lambertian_material.rs:27-34 */
        double r = in * sc.spectra[m.spectrum].intensity_at(wavelength);
        return r * m.p0;
    }
    case ORC_MAT_PHONG: { /* phong_material.rs:16-36 */
        if (w_i.z < 0.0 || w_o.z < 0.0) return 0.0;
        V3 refl = v3(-w_i.x, -w_i.y, w_i.z);
        return in * sc.spectra[m.spectrum].intensity_at(wavelength) * m.p0 +
               std::pow(std::fabs(dot(w_o, refl)), m.p2) * (m.p1 / dot(w_i, v3(0, 0, 1)));
    }
    case ORC_MAT_REFLECTIVE: { /* reflective_material.rs:15-40 */
        if (w_i.z <= 0.0 || w_o.z <= 0.0) return 0.0;
        V3 refl = v3(-w_o.x, -w_o.y, w_o.z);
        double out = in * sc.spectra[m.spectrum].intensity_at(wavelength);
        out *= m.p0;
        double sigma = 0.05, two = 2.0;
        double c = dot(w_i, refl);
        c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c); /* f64::clamp */
        double theta = std::acos(std::fabs(c));
        double rf = m.p1 * std::exp(-(std::pow(theta, two)) / (two * sigma * sigma));
        return out * (1.0 - rf) + rf;
    }
    default: { /* smooth_transparent_dialectric.rs:74-89 */
        double eta = sc.spectra[m.spectrum].intensity_at(wavelength);
        double eta1 = w_i.z >= 0.0 ? 1.0 : eta, eta2 = w_i.z >= 0.0 ? eta : 1.0;
        Fresnel f = fresnel(w_i, eta1, eta2);
        if (norm_squared(w_o - f.reflection_direction) < 0.0000000001) return in * f.reflection_strength;
        if (norm_squared(w_o - f.transmission_direction) < 0.0000000001) return in * f.transmission_strength;
        return 0.0;
    }
    }
}

/* simple_random_integrator.rs:57-65 */
double sky(V3 w, double wavelength) { return reflection_from_linear_rgb(w.y, w.y, 1.0).intensity_at(wavelength); }

struct RenderCtx {
    const OrcScene *sc;
    const OrcRenderParams *p;
    Counters counters;
    uint64_t bounce_rays = 0, shadow_rays = 0, escaped = 0, depth_limited = 0;
};

/* simple_random_integrator.rs:12-55 (recursive, as written) */
Photon integrate_simple(RenderCtx &cx, const Hit &info, Photon photon, Rng &rng, uint32_t limit) {
    if (limit == 0) {
        cx.depth_limited++;
        return Photon{0.0, 0.0};
    }
    const OrcScene &sc = *cx.sc;
    M3 world_to_bsdf = from_rows(info.tangent, info.cotangent, info.normal); /* algebra_utils.rs:3-5 */
    M3 bsdf_to_world;
    if (!try_inverse(world_to_bsdf, &bsdf_to_world)) {
        /* the reference panics here; the oracle reports a NaN sample instead of aborting */
        return Photon{photon.wavelength, std::numeric_limits<double>::quiet_NaN()};
    }
    V3 w_i = mul(world_to_bsdf, info.retro);
    const Material &m = sc.materials[info.material];
    SampleResult s = material_sample(sc, m, w_i, photon.wavelength, rng);
    V3 w_o = s.direction;
    V3 world_w_o = mul(bsdf_to_world, w_o);
    Ray next = ray_bias(ray_new(info.location, world_w_o), cx.p->bias);
    Hit h{};
    cx.bounce_rays++;
    Photon incoming;
    if (!scene_sample(sc, next, cx.p->traverse, &h, &cx.counters)) {
        cx.escaped++;
        incoming = Photon{photon.wavelength, sky(world_w_o, photon.wavelength)};
    } else {
        incoming = integrate_simple(cx, h, photon, rng, limit - 1);
    }
    incoming.intensity = incoming.intensity * s.pdf;                         /* :51 multiplies by the pdf (sic) */
    incoming.intensity = incoming.intensity * std::fabs(dot(world_w_o, info.normal)); /* :52 */
    return Photon{incoming.wavelength, material_bsdf(sc, m, w_o, w_i, incoming.wavelength, incoming.intensity)};
}

/* whitted_integrator.rs:20-86 */
Photon integrate_whitted(RenderCtx &cx, const Hit &info, Photon photon, Rng &rng, uint32_t limit) {
    const OrcScene &sc = *cx.sc;
    M3 world_to_bsdf = from_rows(info.tangent, info.cotangent, info.normal);
    M3 bsdf_to_world;
    if (!try_inverse(world_to_bsdf, &bsdf_to_world)) return Photon{photon.wavelength, std::numeric_limits<double>::quiet_NaN()};
    const Material &m = sc.materials[info.material];
    Photon result = photon; /* fold(photon.clone(), ...) */
    for (uint32_t li = 0; li < cx.p->n_lights; li++) {
        const OrcLight &L = cx.p->lights[li];
        V3 ldir = v3(L.direction[0], L.direction[1], L.direction[2]);
        Hit sh{};
        cx.shadow_rays++;
        double term;
        if (scene_sample(sc, ray_bias(ray_new(info.location, ldir), cx.p->bias), cx.p->traverse, &sh, &cx.counters)) {
            term = cx.p->ambient_spectrum >= 0 ? sc.spectra[cx.p->ambient_spectrum].intensity_at(photon.wavelength) : 0.0;
        } else {
            double emitted = sc.spectra[L.spectrum].intensity_at(photon.wavelength);
            emitted = emitted * std::fabs(dot(ldir, info.normal));
            term = material_bsdf(sc, m, mul(world_to_bsdf, info.retro), mul(world_to_bsdf, ldir), photon.wavelength, emitted);
        }
        result.intensity += term;
    }
    {
        V3 w_retro = mul(world_to_bsdf, info.retro);
        SampleResult s = material_sample(sc, m, w_retro, photon.wavelength, rng);
        V3 world_dir = mul(bsdf_to_world, s.direction);
        Hit h{};
        cx.bounce_rays++;
        double term;
        if (scene_sample(sc, ray_bias(ray_new(info.location, world_dir), cx.p->bias), cx.p->traverse, &h, &cx.counters)) {
            if (limit > 0) {
                Photon rec = integrate_whitted(cx, h, photon, rng, limit - 1);
                double v = material_bsdf(sc, m, w_retro, s.direction, rec.wavelength, rec.intensity);
                term = v * std::fabs(dot(world_dir, info.normal));
            } else {
                term = photon.intensity * 0.0;
            }
        } else {
            term = photon.intensity * 0.0;
        }
        result.intensity += term;
    }
    return result;
}

/* camera.rs:24-66 */
struct ImageSampler {
    uint64_t w, h;
    double film_w, film_h;
    V3 cam;
    ImageSampler(uint64_t width, uint64_t height, V3 camera) : w(width), h(height), cam(camera) {
        double fw = (double)width, fh = (double)height;
        if (fw > fh) film_w = fw / fh, film_h = 1.0;
        else film_w = 1.0, film_h = fw / fh;
    }
    static double scale(uint64_t i, uint64_t n, double l, double u) {
        double pixel_size = l * (1.0 / (double)n);
        return ((double)i + u) * pixel_size;
    }
    Ray ray_for_pixel(uint64_t row, uint64_t col, double ux, double uy) const {
        return ray_new(cam, v3(scale(col, w, film_w, ux) - film_w * 0.5, scale(h - (row + 1), h, film_h, uy) - film_h * 0.5, 1.0));
    }
};

/* accumulation_buffer.rs:44-60 on one pixel */
struct PixelAccum {
    V3 colour{0, 0, 0}, sum{0, 0, 0}, bias{0, 0, 0};
    double weight = 0, weight_bias = 0;
    void update(Photon photon, double w) {
        V3 pc = xyz_from_photon(photon);
        double wy = w - weight_bias;
        double wt = weight + wy;
        weight_bias = (wt - weight) - wy;
        weight = wt;
        V3 cy = pc * w - bias;
        V3 ct = sum + cy;
        bias = (ct - sum) - cy;
        sum = ct;
        colour = sum * (1.0 / weight);
    }
};

/* mesh.rs:13-88 with the behaviour of obj 0.9's Obj::<SimplePolygon>::load restated:
 * v / vn / f lines; index forms a, a/b, a//c, a/b/c; 1-based, negative = relative to the
 * end; positions and normals are parsed as f32 and widened; polygons fan-triangulated
 * around their first vertex; a vertex without a normal index gets a zero normal. */
bool load_obj(const char *path, std::vector<Triangle> *out) {
    FILE *f = std::fopen(path, "r");
    if (!f) return false;
    std::vector<float> pos, nrm;
    char line[4096];
    while (std::fgets(line, sizeof line, f)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char *q = p + 1;
            for (int k = 0; k < 3; k++) pos.push_back(std::strtof(q, &q));
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            char *q = p + 2;
            for (int k = 0; k < 3; k++) nrm.push_back(std::strtof(q, &q));
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            std::vector<std::pair<long, long>> poly; /* (position index, normal index or -1), 0-based */
            char *q = p + 1;
            for (;;) {
                while (*q == ' ' || *q == '\t') q++;
                if (*q == 0 || *q == '\n' || *q == '\r' || *q == '#') break;
                long vi = std::strtol(q, &q, 10), ni = 0;
                bool has_n = false;
                if (*q == '/') {
                    q++;
                    if (*q != '/') std::strtol(q, &q, 10); /* texture index, unused */
                    if (*q == '/') {
                        q++;
                        ni = std::strtol(q, &q, 10);
                        has_n = true;
                    }
                }
                long npos = (long)(pos.size() / 3), nn = (long)(nrm.size() / 3);
                long v0 = vi < 0 ? npos + vi : vi - 1;
                long n0 = has_n ? (ni < 0 ? nn + ni : ni - 1) : -1;
                poly.push_back({v0, n0});
            }
            auto fetch = [&](const std::pair<long, long> &ix, V3 *v, V3 *n) {
                *v = v3((double)pos[ix.first * 3], (double)pos[ix.first * 3 + 1], (double)pos[ix.first * 3 + 2]);
                *n = ix.second >= 0 ? v3((double)nrm[ix.second * 3], (double)nrm[ix.second * 3 + 1], (double)nrm[ix.second * 3 + 2])
                                    : v3(0, 0, 0);
            };
            for (size_t k = 1; k + 1 < poly.size(); k++) { /* mesh.rs:42-72 */
                Triangle t{};
                fetch(poly[0], &t.v[0], &t.n[0]);
                fetch(poly[k], &t.v[1], &t.n[1]);
                fetch(poly[k + 1], &t.v[2], &t.n[2]);
                t.prim_id = (int)out->size();
                out->push_back(t);
            }
        }
    }
    std::fclose(f);
    return true;
}

void fill_hit16(const Hit &h, double *o) {
    o[0] = h.distance;
    const V3 *vs[5] = {&h.location, &h.normal, &h.tangent, &h.cotangent, &h.retro};
    for (int i = 0; i < 5; i++) o[1 + 3 * i] = vs[i]->x, o[2 + 3 * i] = vs[i]->y, o[3 + 3 * i] = vs[i]->z;
}
inline V3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }
inline int all_cores() {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

} // namespace

/* ====================================================================== C interface */
extern "C" {

OrcScene *orc_scene_new(double x, double y, double z) {
    OrcScene *s = new OrcScene();
    s->camera = v3(x, y, z);
    return s;
}
void orc_scene_free(OrcScene *s) { delete s; }
void orc_free(void *p) { std::free(p); }

int orc_add_spectrum(OrcScene *s, double lo, double hi, int n, const double *samples) {
    Spectrum sp;
    sp.shortest = lo, sp.longest = hi;
    sp.samples.assign(samples, samples + n);
    s->spectra.push_back(sp);
    return (int)s->spectra.size() - 1;
}
int orc_add_spectrum_rgb(OrcScene *s, double r, double g, double b) {
    s->spectra.push_back(reflection_from_linear_rgb(r, g, b));
    return (int)s->spectra.size() - 1;
}
int orc_add_spectrum_grey(OrcScene *s, double v) {
    s->spectra.push_back(grey(v));
    return (int)s->spectra.size() - 1;
}
int orc_add_spectrum_diamond(OrcScene *s) {
    s->spectra.push_back(diamond_index_of_refraction());
    return (int)s->spectra.size() - 1;
}
int orc_add_material(OrcScene *s, int kind, int spectrum, double p0, double p1, double p2) {
    s->materials.push_back(Material{kind, spectrum, p0, p1, p2});
    return (int)s->materials.size() - 1;
}
int orc_begin_list(OrcScene *s) {
    s->objects.emplace_back();
    return (int)s->objects.size() - 1;
}
void orc_list_add_sphere(OrcScene *s, double cx, double cy, double cz, double r, int material) {
    ListPrim lp{};
    lp.kind = 0;
    lp.s = Sphere{v3(cx, cy, cz), r, material};
    s->objects.back().prims.push_back(lp);
}
void orc_list_add_plane(OrcScene *s, double nx, double ny, double nz, double d, int material) {
    ListPrim lp{};
    lp.kind = 1;
    lp.p = plane_new(v3(nx, ny, nz), d, material);
    s->objects.back().prims.push_back(lp);
}
void orc_list_add_triangle(OrcScene *s, const double *v9, const double *n9, int material) {
    ListPrim lp{};
    lp.kind = 2;
    for (int i = 0; i < 3; i++) lp.t.v[i] = ld3(v9 + 3 * i), lp.t.n[i] = ld3(n9 + 3 * i);
    lp.t.material = material;
    s->objects.back().prims.push_back(lp);
}
static int add_bvh_from(OrcScene *s, std::vector<Triangle> &&tris, int material) {
    s->objects.emplace_back();
    Object &o = s->objects.back();
    o.is_bvh = true;
    o.bvh.reset(new Bvh());
    o.bvh->tris = std::move(tris);
    for (auto &t : o.bvh->tris) t.material = material;
    o.bvh->build_all();
    return (int)s->objects.size() - 1;
}
int orc_add_bvh(OrcScene *s, int64_t ntri, const double *verts, const double *normals, int material) {
    std::vector<Triangle> tris((size_t)ntri);
    for (int64_t i = 0; i < ntri; i++) {
        for (int k = 0; k < 3; k++) tris[i].v[k] = ld3(verts + i * 9 + 3 * k), tris[i].n[k] = ld3(normals + i * 9 + 3 * k);
        tris[i].prim_id = (int)i;
    }
    return add_bvh_from(s, std::move(tris), material);
}
int orc_add_bvh_obj(OrcScene *s, const char *path, int material) {
    std::vector<Triangle> tris;
    if (!load_obj(path, &tris)) return -1;
    return add_bvh_from(s, std::move(tris), material);
}
int64_t orc_bvh_triangle_count(const OrcScene *s, int object_id) {
    const Object &o = s->objects[object_id];
    return o.is_bvh ? (int64_t)o.bvh->tris.size() : -1;
}
int orc_bvh_depth(const OrcScene *s, int object_id) {
    const Object &o = s->objects[object_id];
    return o.is_bvh ? o.bvh->depth : -1;
}
/* prim_id (index in the order the mesh was handed over) of the triangle at each leaf, in DFS order */
int64_t orc_bvh_leaf_order(const OrcScene *s, int object_id, int32_t *out) {
    const Object &o = s->objects[object_id];
    if (!o.is_bvh) return -1;
    for (size_t i = 0; i < o.bvh->tris.size(); i++) out[i] = (int32_t)o.bvh->tris[i].prim_id;
    return (int64_t)o.bvh->tris.size();
}
int64_t orc_load_obj(const char *path, double **verts, double **normals) {
    std::vector<Triangle> tris;
    if (!load_obj(path, &tris)) return -1;
    *verts = (double *)std::malloc(sizeof(double) * 9 * std::max<size_t>(1, tris.size()));
    *normals = (double *)std::malloc(sizeof(double) * 9 * std::max<size_t>(1, tris.size()));
    for (size_t i = 0; i < tris.size(); i++)
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++) {
                (*verts)[i * 9 + 3 * k + c] = tris[i].v[k][c];
                (*normals)[i * 9 + 3 * k + c] = tris[i].n[k][c];
            }
    return (int64_t)tris.size();
}

void orc_trace_rays(const OrcScene *s, int64_t n, const double *origins, const double *dirs, int mode, int32_t *object_id,
                    int32_t *prim_id, double *t, OrcTraceCounters *counters) {
    uint64_t nv = 0, nt = 0, nh = 0;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(all_cores()) reduction(+ : nv, nt, nh)
    for (int64_t i = 0; i < n; i++) {
        Ray r = ray_new(ld3(origins + 3 * i), ld3(dirs + 3 * i));
        Hit h{};
        Counters c;
        bool hit = scene_sample(*s, r, mode, &h, &c);
        object_id[i] = hit ? h.object_id : -1;
        prim_id[i] = hit ? h.prim_id : -1;
        t[i] = hit ? h.distance : kInf;
        nv += c.node_visits, nt += c.tri_tests, nh += hit ? 1 : 0;
    }
    if (counters) counters->rays = (uint64_t)n, counters->node_visits = nv, counters->tri_tests = nt, counters->hits = nh;
}
void orc_trace_rays_edge_distance(const OrcScene *s, int64_t n, const double *origins, const double *dirs, double *min_bary) {
#pragma omp parallel for schedule(dynamic, 1024) num_threads(all_cores())
    for (int64_t i = 0; i < n; i++) {
        Ray r = ray_new(ld3(origins + 3 * i), ld3(dirs + 3 * i));
        Hit h{};
        Counters c;
        min_bary[i] = scene_sample(*s, r, ORC_TRAVERSE_REFERENCE, &h, &c) ? h.min_bary : 2.0;
    }
}

/* camera.rs:95-130, generalised to spp samples per pixel (sample s of a pixel is applied to the
 * pixel's accumulator in order s = offset .. offset+spp-1; spp = 1 is the reference call). */
void orc_render_tile(const OrcScene *s, const uint64_t tile[4], uint64_t height, uint64_t width, const OrcRenderParams *p,
                     double *colour_sum, double *colour_bias, double *weight, double *weight_bias, double *colour,
                     double *photons, OrcRenderStats *stats) {
    uint64_t tw = tile[1] - tile[0], th = tile[3] - tile[2];
    uint64_t npix = tw * th;
    ImageSampler cam(width, height, s->camera);
    uint64_t primary = 0, bounce = 0, shadow = 0, nv = 0, nt = 0, missed = 0, escaped = 0, limited = 0;
#ifdef _OPENMP
    int nthreads = p->threads ? (int)p->threads : omp_get_num_procs(); /* all host cores, whatever OMP_NUM_THREADS says */
#else
    int nthreads = 1;
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) reduction(+ : primary, bounce, shadow, nv, nt, missed, escaped, limited)
    for (int64_t pi = 0; pi < (int64_t)npix; pi++) {
        uint64_t row = (uint64_t)pi / tw, col = (uint64_t)pi % tw;
        uint64_t grow = tile[2] + row, gcol = tile[0] + col;
        PixelAccum acc;
        RenderCtx cx;
        cx.sc = s, cx.p = p;
        for (uint32_t k = 0; k < p->spp; k++) {
            Rng rng{p->seed, (uint32_t)(grow * width + gcol), p->sample_offset + k, 0};
            double ux = rng.f64(), uy = rng.f64();
            Ray ray = cam.ray_for_pixel(grow, gcol, ux, uy);
            Hit h{};
            primary++;
            Photon photon;
            if (!scene_sample(*s, ray, p->traverse, &h, &cx.counters)) {
                missed++;
                photon = Photon{0.0, 0.0}; /* camera.rs:110-113 */
            } else {
                Photon start{380.0 + (740.0 - 380.0) * rng.f64(), 0.0}; /* photon.rs:18-24 */
                rng.ordinal = 4; /* ordinal 3 is left unused: material draws start on a Philox block boundary */
                photon = p->integrator == ORC_INTEGRATOR_WHITTED ? integrate_whitted(cx, h, start, rng, p->max_depth)
                                                                  : integrate_simple(cx, h, start, rng, p->max_depth);
            }
            photon.intensity = photon.intensity * (740.0 - 380.0); /* camera.rs:124, photon.rs:26-28 */
            if (photons) photons[(k * npix + pi) * 2] = photon.wavelength, photons[(k * npix + pi) * 2 + 1] = photon.intensity;
            acc.update(photon, 1.0);
        }
        if (colour_sum) colour_sum[pi * 3] = acc.sum.x, colour_sum[pi * 3 + 1] = acc.sum.y, colour_sum[pi * 3 + 2] = acc.sum.z;
        if (colour_bias) colour_bias[pi * 3] = acc.bias.x, colour_bias[pi * 3 + 1] = acc.bias.y, colour_bias[pi * 3 + 2] = acc.bias.z;
        if (colour) colour[pi * 3] = acc.colour.x, colour[pi * 3 + 1] = acc.colour.y, colour[pi * 3 + 2] = acc.colour.z;
        if (weight) weight[pi] = acc.weight;
        if (weight_bias) weight_bias[pi] = acc.weight_bias;
        bounce += cx.bounce_rays, shadow += cx.shadow_rays, nv += cx.counters.node_visits, nt += cx.counters.tri_tests;
        escaped += cx.escaped, limited += cx.depth_limited;
    }
    if (stats) {
        stats->primary_rays = primary, stats->bounce_rays = bounce, stats->shadow_rays = shadow;
        stats->node_visits = nv, stats->tri_tests = nt;
        stats->paths_missed = missed, stats->paths_escaped = escaped, stats->paths_depth_limited = limited;
    }
}

/* ---- unit-level entry points ---- */
int orc_triangle_intersect(const double *v9, const double *n9, const double *origin, const double *dir, double *out16) {
    Triangle t{};
    for (int i = 0; i < 3; i++) t.v[i] = ld3(v9 + 3 * i), t.n[i] = ld3(n9 + 3 * i);
    Hit h{};
    if (!triangle_intersect(t, ray_new(ld3(origin), ld3(dir)), &h)) return 0;
    fill_hit16(h, out16);
    return 1;
}
int orc_sphere_intersect(const double *centre, double radius, const double *origin, const double *dir, double *out16) {
    Hit h{};
    if (!sphere_intersect(Sphere{ld3(centre), radius, 0}, ray_new(ld3(origin), ld3(dir)), &h)) return 0;
    fill_hit16(h, out16);
    return 1;
}
int orc_plane_intersect(const double *normal, double dist, const double *origin, const double *dir, double *out16) {
    Hit h{};
    if (!plane_intersect(plane_new(ld3(normal), dist, 0), ray_new(ld3(origin), ld3(dir)), &h)) return 0;
    fill_hit16(h, out16);
    return 1;
}
int orc_aabb_intersect(const double *lo, const double *hi, const double *origin, const double *dir) {
    /* BoundingBox::from_corners (util/axis_aligned_bounding_box.rs:12-22) sorts each axis */
    Box b = {{interval_new(lo[0], hi[0]), interval_new(lo[1], hi[1]), interval_new(lo[2], hi[2])}};
    return box_intersect(b, ray_new(ld3(origin), ld3(dir))) ? 1 : 0;
}
void orc_triangle_helpers(const double *dir, int *perm3, double *shear2) {
    V3 d = ld3(dir);
    permutation_largest_last(d, perm3);
    V3 pd = permute(d, perm3);
    shear2[0] = -pd.x / pd.z, shear2[1] = -pd.y / pd.z;
}
double orc_spectrum_intensity(double lo, double hi, int n, const double *samples, double wavelength) {
    Spectrum s;
    s.shortest = lo, s.longest = hi;
    s.samples.assign(samples, samples + n);
    return s.intensity_at(wavelength);
}
void orc_rgb_to_spectrum(double r, double g, double b, double *samples32) {
    Spectrum s = reflection_from_linear_rgb(r, g, b);
    std::memcpy(samples32, s.samples.data(), sizeof(double) * 32);
}
void orc_cmf_xyz(double wavelength, double *xyz) {
    V3 c = cmf(wavelength);
    xyz[0] = c.x, xyz[1] = c.y, xyz[2] = c.z;
}
/* colour_xyz.rs:48-67 */
void orc_xyz_to_linear_rgb(const double *xyz, double *rgb) {
    M3 m = from_rows(v3(3.24096994, -1.53738318, -0.49861076), v3(-0.96924364, 1.87596750, 0.04155506),
                     v3(0.05563008, -0.20397696, 1.05697151));
    V3 r = mul(m, ld3(xyz));
    rgb[0] = r.x, rgb[1] = r.y, rgb[2] = r.z;
}
void orc_linear_rgb_to_xyz(const double *rgb, double *xyz) {
    M3 m = from_rows(v3(0.41239080, 0.35758434, 0.18048079), v3(0.21263901, 0.71516868, 0.07219232),
                     v3(0.01933082, 0.11919478, 0.95053215));
    V3 r = mul(m, ld3(rgb));
    xyz[0] = r.x, xyz[1] = r.y, xyz[2] = r.z;
}
/* colour_xyz.rs:78-84 (constants as written there) */
double orc_srgb_gamma(double u) { return u <= 0.0031308 ? 12.98 * u : 1.005 * std::pow(u, 1.0 / 2.4) - 0.055; }

/* image.rs:130-187 ClampingToneMapper; image.rs:120-123 normalized_to_byte (truncating, saturating cast; NaN -> 0).
 * source 0: ColourXyz::to_srgb first (colour_xyz.rs:48-84); source 1: linear RGB as is. */
void orc_tone_map(int source, const double *colour, int64_t n, uint8_t *rgb8) {
    for (int64_t i = 0; i < n; i++) {
        double c[3] = {colour[3 * i], colour[3 * i + 1], colour[3 * i + 2]};
        if (source == 0) {
            double lin[3];
            orc_xyz_to_linear_rgb(c, lin);
            for (int k = 0; k < 3; k++) c[k] = orc_srgb_gamma(lin[k]);
        }
        for (int k = 0; k < 3; k++) {
            double v = c[k] < 0.0 ? 0.0 : (c[k] > 1.0 ? 1.0 : c[k]); /* f64::clamp keeps NaN */
            double b = v * 255.0;
            rgb8[3 * i + k] = b != b ? 0 : (uint8_t)(int)b;
        }
    }
}

void orc_accum_update(double *st, double wavelength, double intensity, double weight) {
    PixelAccum a;
    a.colour = ld3(st), a.sum = ld3(st + 3), a.bias = ld3(st + 6), a.weight = st[9], a.weight_bias = st[10];
    a.update(Photon{wavelength, intensity}, weight);
    const V3 *vs[3] = {&a.colour, &a.sum, &a.bias};
    for (int i = 0; i < 3; i++) st[3 * i] = vs[i]->x, st[3 * i + 1] = vs[i]->y, st[3 * i + 2] = vs[i]->z;
    st[9] = a.weight, st[10] = a.weight_bias;
}
/* accumulation_buffer.rs:81-85 */
void orc_accum_blend(const double *c1, double w1, const double *c2, double w2, double *out3) {
    V3 r = (ld3(c1) * w1 + ld3(c2) * w2) * (1.0 / (w1 + w2));
    out3[0] = r.x, out3[1] = r.y, out3[2] = r.z;
}
void orc_camera_ray(uint64_t width, uint64_t height, const double *cam, uint64_t row, uint64_t col, double ux, double uy,
                    double *origin3, double *dir3) {
    ImageSampler is(width, height, ld3(cam));
    Ray r = is.ray_for_pixel(row, col, ux, uy);
    origin3[0] = r.origin.x, origin3[1] = r.origin.y, origin3[2] = r.origin.z;
    dir3[0] = r.direction.x, dir3[1] = r.direction.y, dir3[2] = r.direction.z;
}
int orc_mat3_inverse(const double *m9, double *out9) {
    M3 m, inv;
    std::memcpy(m.e, m9, sizeof m.e);
    if (!try_inverse(m, &inv)) return 0;
    std::memcpy(out9, inv.e, sizeof inv.e);
    return 1;
}
double orc_mat3_determinant(const double *m9) {
    M3 m;
    std::memcpy(m.e, m9, sizeof m.e);
    return determinant(m);
}
int orc_largest_dimension(const double *lo, const double *hi) {
    Box b = {{{lo[0], hi[0]}, {lo[1], hi[1]}, {lo[2], hi[2]}}};
    return largest_dimension(b);
}
/* util/tile_iterator.rs:41-66 -- row-major walk; tiles[4*i..] = start_col, end_col, start_row, end_row */
int64_t orc_tile_iterator(uint64_t width, uint64_t height, uint64_t tile_size, uint64_t *tiles, int64_t cap) {
    int64_t n = 0;
    uint64_t col = 0, row = 0;
    while (row < height) {
        uint64_t sc = col, ec = std::min(width, sc + tile_size), sr = row, er = std::min(height, sr + tile_size);
        col += tile_size;
        if (col >= width) row += tile_size, col = 0;
        if (n < cap) tiles[4 * n] = sc, tiles[4 * n + 1] = ec, tiles[4 * n + 2] = sr, tiles[4 * n + 3] = er;
        n++;
    }
    return n;
}
void orc_material_sample(const OrcScene *s, int material, const double *w_i, double wavelength, uint64_t seed, uint32_t pixel,
                         uint64_t sample, uint32_t first_ordinal, double *dir3, double *pdf, uint32_t *draws_used) {
    Rng rng{seed, pixel, sample, first_ordinal};
    SampleResult r = material_sample(*s, s->materials[material], ld3(w_i), wavelength, rng);
    dir3[0] = r.direction.x, dir3[1] = r.direction.y, dir3[2] = r.direction.z;
    *pdf = r.pdf;
    if (draws_used) *draws_used = rng.ordinal - first_ordinal;
}
double orc_material_bsdf(const OrcScene *s, int material, const double *w_o, const double *w_i, double wavelength, double in) {
    return material_bsdf(*s, s->materials[material], ld3(w_o), ld3(w_i), wavelength, in);
}
double orc_sky(const double *w, double wavelength) { return sky(ld3(w), wavelength); }

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
double orc_rng_f64(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal) { return Rng{seed, pixel, sample, ordinal}.f64(); }
double orc_rng_open01(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal) {
    return Rng{seed, pixel, sample, ordinal}.open01();
}
int orc_rng_bool(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t ordinal) {
    return Rng{seed, pixel, sample, ordinal}.boolean() ? 1 : 0;
}

} // extern "C"
