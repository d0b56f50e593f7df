// vrj_device.cuh -- device-side arithmetic of the render loop, binary64, in the reference's
// operation order (compile with -fmad=false: rustc never contracts a*b+c).
// Reference paths are relative to /root/reference/src/.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

// Experiment knobs (code size vs call overhead): which helpers are real functions instead of inlined copies.
#ifndef VRJ_OUTLINE
#define VRJ_OUTLINE 1
#endif
#define VRJ_FN_NORM ((VRJ_OUTLINE) & 1)
#define VRJ_FN_PHILOX ((VRJ_OUTLINE) & 2)
#define VRJ_FN_PRIM ((VRJ_OUTLINE) & 4)
#if VRJ_FN_NORM
#define VRJ_INL_NORM __noinline__
#else
#define VRJ_INL_NORM __forceinline__
#endif
#if VRJ_FN_PHILOX
#define VRJ_INL_PHILOX __noinline__
#else
#define VRJ_INL_PHILOX __forceinline__
#endif
#if VRJ_FN_PRIM
#define VRJ_INL_PRIM __noinline__
#else
#define VRJ_INL_PRIM __forceinline__
#endif

namespace vrj {

// ------------------------------------------------------------------------------------------
// math/vec3.rs, math/mat3.rs, math/mat2.rs
struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
// vec3.rs:76-82: products summed from 0.0 in x, y, z order
__device__ __forceinline__ double dot(D3 a, D3 b) { return ((0.0 + a.x * b.x) + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double norm(D3 a) { return sqrt(dot(a, a)); }
// vec3.rs:103-110: multiply by 1/norm
__device__ VRJ_INL_NORM D3 normalize(D3 a) {
    double inv = 1.0 / norm(a);
    return d3(a.x * inv, a.y * inv, a.z * inv);
}

struct M3 {
    double e[3][3];
};
// mat3.rs:72-94 (first_minor with mat2.rs:13-15, cofactor sign), written out per element
__device__ __forceinline__ double minor2(double a, double b, double c, double d) { return a * d - b * c; }
__device__ __forceinline__ double first_minor(const M3 &m, int r, int c) {
    const int r0 = r == 0 ? 1 : 0, r1 = r == 2 ? 1 : 2;
    const int c0 = c == 0 ? 1 : 0, c1 = c == 2 ? 1 : 2;
    return minor2(m.e[r0][c0], m.e[r0][c1], m.e[r1][c0], m.e[r1][c1]);
}
__device__ __forceinline__ double cofactor(const M3 &m, int r, int c) {
    return (((r + c) & 1) ? -1.0 : 1.0) * first_minor(m, r, c);
}
// mat3.rs:106-109
__device__ __forceinline__ double determinant(const M3 &m) {
    return m.e[0][0] * first_minor(m, 0, 0) - m.e[0][1] * first_minor(m, 0, 1) + m.e[0][2] * first_minor(m, 0, 2);
}
// mat3.rs:111-118: transpose(cofactor matrix) * determinant (sic)
__device__ __forceinline__ bool try_inverse(const M3 &m, M3 &out) {
    double det = determinant(m);
    if (det == 0.0) return false;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) out.e[i][j] = cofactor(m, j, i) * det;
    return true;
}
// mat3.rs:147-157
__device__ __forceinline__ D3 mul(const M3 &m, D3 v) {
    return d3(dot(d3(m.e[0][0], m.e[0][1], m.e[0][2]), v), dot(d3(m.e[1][0], m.e[1][1], m.e[1][2]), v),
              dot(d3(m.e[2][0], m.e[2][1], m.e[2][2]), v));
}
// util/algebra_utils.rs:3-5 + mat3.rs:34-42
__device__ __forceinline__ M3 from_rows(D3 a, D3 b, D3 c) {
    M3 m;
    m.e[0][0] = a.x, m.e[0][1] = a.y, m.e[0][2] = a.z;
    m.e[1][0] = b.x, m.e[1][1] = b.y, m.e[1][2] = b.z;
    m.e[2][0] = c.x, m.e[2][1] = c.y, m.e[2][2] = c.z;
    return m;
}

// ------------------------------------------------------------------------------------------
// Counter-based RNG replacing rand 0.7: Philox-4x32-10,
// key = seed, counter = (draw ordinal / 2, pixel, sample lo, sample hi); each block yields two
// 64-bit draws.  The last block is cached so consecutive draws cost one Philox call per pair.
struct Rng {
    uint32_t k0, k1, pixel, s0, s1;
    uint32_t ordinal;
    uint32_t cached_block;
    uint32_t w[4];

    __device__ __forceinline__ void init(uint64_t seed, uint32_t pixel_, uint64_t sample, uint32_t first_ordinal) {
        k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        pixel = pixel_, s0 = (uint32_t)sample, s1 = (uint32_t)(sample >> 32);
        ordinal = first_ordinal;
        cached_block = 0xffffffffu;
    }
    __device__ VRJ_INL_PHILOX void block(uint32_t b) {
        uint32_t c0 = b, c1 = pixel, c2 = s0, c3 = s1, ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            if (r) ka += 0x9E3779B9u, kb += 0xBB67AE85u;
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
            c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
        }
        w[0] = c0, w[1] = c1, w[2] = c2, w[3] = c3;
        cached_block = b;
    }
    __device__ __forceinline__ uint64_t bits() {
        uint32_t b = ordinal >> 1;
        if (b != cached_block) block(b);
        uint32_t lo = (ordinal & 1) ? w[2] : w[0], hi = (ordinal & 1) ? w[3] : w[1];
        ordinal++;
        return ((uint64_t)hi << 32) | lo;
    }
    // rand 0.7 Standard f64: 53 bits -> [0,1)
    __device__ __forceinline__ double f64() { return (double)(bits() >> 11) * (1.0 / 9007199254740992.0); }
    // rand 0.7 Open01: 52 bits -> (0,1)
    __device__ __forceinline__ double open01() { return ((double)(bits() >> 12) + 0.5) * (1.0 / 4503599627370496.0); }
    // rand 0.7 Standard bool: sign bit
    __device__ __forceinline__ bool boolean() { return (bits() >> 63) != 0; }
};

// ------------------------------------------------------------------------------------------
// colour/spectrum.rs
struct SpectrumDev {
    double shortest, longest;
    uint32_t first, n;
};

__device__ const double g_rgb_basis[7][32] = {
#include "rgb_basis_tables.inc"
};
enum { B_WHITE = 0, B_CYAN, B_MAGENTA, B_YELLOW, B_RED, B_GREEN, B_BLUE };

// spectrum.rs:50-79, `sample(i)` supplies sample i
template <typename F>
__device__ __forceinline__ double spectrum_lookup(double shortest, double longest, uint32_t n, double wavelength, F sample) {
    if (wavelength < shortest || wavelength > longest) return 0.0;
    double range = longest - shortest;
    double nm1 = (double)(n - 1);
    double fidx = nm1 * ((wavelength - shortest) / range);
    uint32_t before = (fidx != fidx || fidx < 0.0) ? 0u : (uint32_t)fidx;
    double wl_before = (double)before / nm1 * range + shortest;
    if (before == n - 1) return sample(before);
    double wl_after = (double)(before + 1) / nm1 * range + shortest;
    double delta = wl_after - wl_before;
    double ratio = (wavelength - wl_before) / delta;
    return sample(before) * (1.0 - ratio) + sample(before + 1) * ratio;
}
__device__ __forceinline__ double spectrum_intensity(const SpectrumDev *__restrict__ spectra,
                                                     const double *__restrict__ samples, uint32_t id, double wavelength) {
    SpectrumDev s = spectra[id];
    const double *p = samples + s.first;
    return spectrum_lookup(s.shortest, s.longest, s.n, wavelength, [p](uint32_t i) { return __ldg(p + i); });
}
// spectrum.rs:81-165 evaluated lazily: only the (at most two) samples the lookup touches are formed
__device__ __forceinline__ double rgb_reflection_intensity(double r, double g, double b, double wavelength) {
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
    return spectrum_lookup(380.0, 720.0, 32u, wavelength, [=](uint32_t i) {
        return c0 * g_rgb_basis[B_WHITE][i] + c1 * g_rgb_basis[second][i] + c2 * g_rgb_basis[third][i];
    });
}
// integrators/simple_random_integrator.rs:57-65
__device__ __forceinline__ double sky(D3 w, double wavelength) { return rgb_reflection_intensity(w.y, w.y, 1.0, wavelength); }

// colour/colour_xyz.rs:86-103
__device__ __forceinline__ double gaussian(double w, double alpha, double mu, double s1, double s2) {
    double sigma = w < mu ? s1 : s2;
    double denominator = 2.0 * (sigma * sigma);
    return alpha * exp(-((w - mu) * (w - mu)) / denominator);
}
__device__ __forceinline__ D3 cmf(double w) {
    double x = gaussian(w, 1.056, 599.8, 37.9, 31.0) + gaussian(w, 0.362, 442.0, 16.0, 26.7) +
               gaussian(w, -0.065, 501.1, 20.4, 26.2);
    double y = gaussian(w, 0.821, 568.8, 46.9, 40.5) + gaussian(w, 0.286, 530.9, 16.3, 31.1);
    double z = gaussian(w, 1.217, 437.0, 11.8, 36.0) + gaussian(w, 0.681, 459.0, 26.0, 13.8);
    return d3(x, y, z);
}

// ------------------------------------------------------------------------------------------
// raycasting: exact primitive tests (binary64)
struct HitFrame {
    double distance;
    D3 location, normal, tangent, cotangent, retro;
    uint32_t material;
};

// Per-ray constants of Triangle::intersect: permutation (triangle.rs:108-122, SIGNED largest
// component last, cyclic permutations only) and shear (triangle.rs:133-135).
struct TriRay {
    D3 o;
    double sx, sy, pdz;
    int perm; // 0: (x,y,z)  1: (y,z,x)  2: (z,x,y)
};
__device__ __forceinline__ D3 permute(D3 v, int perm) {
    return perm == 0 ? v : (perm == 1 ? d3(v.y, v.z, v.x) : d3(v.z, v.x, v.y));
}
__device__ __forceinline__ TriRay tri_ray(D3 o, D3 d) {
    TriRay r;
    r.o = o;
    if (d.x > d.y) r.perm = (d.z > d.x) ? 0 : 1;
    else r.perm = (d.z > d.y) ? 0 : 2;
    D3 pd = permute(d, r.perm);
    r.sx = -pd.x / pd.z, r.sy = -pd.y / pd.z, r.pdz = pd.z;
    return r;
}
__device__ __forceinline__ double edge_fn(D3 a, D3 b) { return a.x * b.y - b.x * a.y; }

// triangle.rs:35-72: returns true and the barycentrics + distance when the ray hits
__device__ VRJ_INL_PRIM bool triangle_test(const TriRay &r, D3 v0, D3 v1, D3 v2, double &distance, double &b0,
                                              double &b1, double &b2, D3 &location) {
    D3 p0 = permute(v0 - r.o, r.perm), p1 = permute(v1 - r.o, r.perm), p2 = permute(v2 - r.o, r.perm);
    D3 t0 = d3(p0.x + r.sx * p0.z, p0.y + r.sy * p0.z, p0.z);
    D3 t1 = d3(p1.x + r.sx * p1.z, p1.y + r.sy * p1.z, p1.z);
    D3 t2 = d3(p2.x + r.sx * p2.z, p2.y + r.sy * p2.z, p2.z);
    double e0 = edge_fn(t1, t2), e1 = edge_fn(t2, t0), e2 = edge_fn(t0, t1);
    // sign BITS, so +-0 matter (triangle.rs:52-53)
    int neg = (__double2hiint(e0) < 0) + (__double2hiint(e1) < 0) + (__double2hiint(e2) < 0);
    if (neg != 0 && neg != 3) return false;
    double a0 = fabs(e0), a1 = fabs(e1), a2 = fabs(e2);
    double inv = 1.0 / (((0.0 + a0) + a1) + a2);
    b0 = a0 * inv, b1 = a1 * inv, b2 = a2 * inv;
    double tz = ((0.0 + t0.z * b0) + t1.z * b1) + t2.z * b2;
    if ((__double2hiint(tz) < 0) != (__double2hiint(r.pdz) < 0)) return false;
    location = ((d3(0.0, 0.0, 0.0) + v0 * b0) + v1 * b1) + v2 * b2;
    distance = norm(r.o - location);
    return true;
}

struct SphereDev {
    double cx, cy, cz, radius;
    uint32_t material, pad;
};
// sphere.rs:39-75 (distance only)
__device__ VRJ_INL_PRIM bool sphere_test(const SphereDev &s, D3 o, D3 d, double &distance) {
    D3 c = d3(s.cx, s.cy, s.cz);
    double a = ((0.0 + d.x * d.x) + d.y * d.y) + d.z * d.z;
    double b = ((0.0 + (o.x * d.x - c.x * d.x) * 2.0) + (o.y * d.y - c.y * d.y) * 2.0) + (o.z * d.z - c.z * d.z) * 2.0;
    double cc = (((0.0 + ((o.x * o.x + c.x * c.x) - c.x * o.x * 2.0)) + ((o.y * o.y + c.y * c.y) - c.y * o.y * 2.0)) +
                 ((o.z * o.z + c.z * c.z) - c.z * o.z * 2.0)) -
                s.radius * s.radius;
    double delta_squared = b * b - 4.0 * a * cc;
    if (delta_squared < 0.0) return false;
    double delta = sqrt(delta_squared);
    double one_over_2a = 1.0 / (2.0 * a);
    double t1 = (-b - delta) * one_over_2a;
    double t2 = (-b + delta) * one_over_2a;
    distance = (t1 < 0.0 || (t2 >= 0.0 && t1 >= t2)) ? t2 : t1;
    return !(distance <= 0.0);
}
// sphere.rs:76-90
__device__ __forceinline__ void sphere_frame(const SphereDev &s, D3 o, D3 d, double distance, HitFrame &h) {
    h.distance = distance;
    h.location = o + d * distance;
    h.normal = normalize(h.location - d3(s.cx, s.cy, s.cz));
    h.tangent = normalize(cross(h.normal, d3(0.0, 0.0, 1.0)));
    h.cotangent = cross(h.normal, h.tangent);
    h.retro = -d;
    h.material = s.material;
}

struct PlaneDev {
    double n[3], t[3], c[3];
    double distance;
    uint32_t material, pad;
};
// plane.rs:48-63 (distance only)
__device__ __forceinline__ bool plane_test(const PlaneDev &p, D3 o, D3 d, double &t) {
    D3 n = d3(p.n[0], p.n[1], p.n[2]);
    double d_dot_n = dot(d, n);
    D3 point_on_plane = n * p.distance;
    double num = dot(point_on_plane - o, n);
    if (d_dot_n == 0.0 && num != 0.0) return false;
    t = num / d_dot_n;
    return !(t < 0.0);
}
// plane.rs:64-73
__device__ __forceinline__ void plane_frame(const PlaneDev &p, D3 o, D3 d, double t, HitFrame &h) {
    h.distance = t;
    h.location = o + d * t;
    h.normal = d3(p.n[0], p.n[1], p.n[2]);
    h.tangent = d3(p.t[0], p.t[1], p.t[2]);
    h.cotangent = d3(p.c[0], p.c[1], p.c[2]);
    h.retro = -d;
    h.material = p.material;
}

// ------------------------------------------------------------------------------------------
// materials
struct MaterialDev {
    uint32_t kind, spectrum;
    double p0, p1, p2;
};
struct Fresnel {
    D3 reflection_direction, transmission_direction;
    double reflection_strength, transmission_strength;
};
// smooth_transparent_dialectric.rs:15-60
__device__ __forceinline__ Fresnel fresnel(D3 w_i, double eta1, double eta2) {
    D3 normal = w_i.z > 0.0 ? d3(0.0, 0.0, 1.0) : -d3(0.0, 0.0, 1.0);
    Fresnel f;
    f.reflection_direction = d3(-w_i.x, -w_i.y, w_i.z);
    double r = eta1 / eta2;
    double cos1 = dot(normal, w_i);
    double cos2sq = 1.0 - r * r * (1.0 - cos1 * cos1);
    if (cos2sq >= 0.0) {
        double cos2 = sqrt(cos2sq);
        double rpar = (eta1 * cos2 - eta2 * cos1) / (eta1 * cos2 + eta2 * cos1);
        double rperp = (eta1 * cos1 - eta2 * cos2) / (eta1 * cos1 + eta2 * cos2);
        f.reflection_strength = 0.5 * (rpar * rpar + rperp * rperp);
        f.transmission_direction = normalize((w_i * (-r)) + (normal * (r * cos1 - cos2)));
        f.transmission_strength = 1.0 - f.reflection_strength;
    } else {
        f.reflection_strength = 1.0;
        f.transmission_strength = 0.0;
        f.transmission_direction = d3(0.0, 0.0, 0.0);
    }
    if (w_i.z < 0.0) {
        f.reflection_direction.z *= -1.0;
        f.transmission_direction.z *= -1.0;
    }
    return f;
}

#define VRJ_PI 3.14159265358979323846264338327950288

// Material::sample: returns direction (BSDF space) and pdf, consuming draws from rng
__device__ __forceinline__ void material_sample(const MaterialDev &m, double eta_or_zero, D3 w_i, Rng &rng, D3 &dir, double &pdf) {
    if (m.kind == 0) { // lambertian_material.rs:36-59 (rejection in the unit disc)
        double x = 2.0 * rng.open01() - 1.0;
        double y = 2.0 * rng.open01() - 1.0;
        while (((0.0 + x * x) + y * y) + 0.0 * 0.0 > 1.0) {
            x = 2.0 * rng.open01() - 1.0;
            y = 2.0 * rng.open01() - 1.0;
        }
        double z = fmax(sqrt(1.0 - x * x - y * y), 0.0);
        double cos_theta = ((0.0 + x * 0.0) + y * 0.0) + z * 1.0;
        double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
        dir = normalize(d3(x, y, z));
        pdf = (cos_theta * sin_theta) / VRJ_PI;
    } else if (m.kind == 2) { // reflective_material.rs:42-47
        dir = d3(-w_i.x, -w_i.y, w_i.z);
        pdf = 1.0;
    } else if (m.kind == 3) { // smooth_transparent_dialectric.rs:91-114
        double eta1 = w_i.z >= 0.0 ? 1.0 : eta_or_zero, eta2 = w_i.z >= 0.0 ? eta_or_zero : 1.0;
        Fresnel f = fresnel(w_i, eta1, eta2);
        pdf = 0.5;
        if (f.transmission_strength <= 0.0000000001) dir = f.reflection_direction;
        else if (f.reflection_strength <= 0.0000000001 || rng.boolean()) dir = f.transmission_direction;
        else dir = f.reflection_direction;
    } else { // materials/mod.rs:28-33 -> cosine_weighted_hemisphere.rs:19-33, unit_disc.rs:27-44, uniform_square.rs:20-25
        double sx = -1.0 + rng.open01() * 2.0;
        double sy = -1.0 + rng.open01() * 2.0;
        double dx, dy;
        if (sx == 0.0 && sy == 0.0) {
            dx = sx, dy = sy;
        } else {
            double radius, angle;
            if (fabs(sx) > fabs(sy)) radius = sx, angle = (VRJ_PI / 4.0) * sy / sx;
            else radius = sy, angle = VRJ_PI / 2.0 - (VRJ_PI / 4.0) * sx / sy;
            dx = cos(angle) * radius, dy = sin(angle) * radius;
        }
        double z = sqrt(fmax(0.0, 1.0 - dx * dx - dy * dy));
        dir = d3(dx, dy, z);
        pdf = sqrt(dx * dx + dy * dy) / VRJ_PI;
    }
}

// Material::bsdf as an affine map of the incoming intensity: out = a * in + b.
// `s` is the material spectrum at the photon's wavelength (colour, or eta for the dielectric).
__device__ __forceinline__ void material_bsdf_affine(const MaterialDev &m, double s, D3 w_o, D3 w_i, double &a, double &b) {
    if (m.kind == 0) { // lambertian_material.rs:27-34
        a = s * m.p0, b = 0.0;
    } else if (m.kind == 1) { // phong_material.rs:16-36
        if (w_i.z < 0.0 || w_o.z < 0.0) {
            a = 0.0, b = 0.0;
        } else {
            D3 refl = d3(-w_i.x, -w_i.y, w_i.z);
            a = s * m.p0;
            b = pow(fabs(dot(w_o, refl)), m.p2) * (m.p1 / dot(w_i, d3(0.0, 0.0, 1.0)));
        }
    } else if (m.kind == 2) { // reflective_material.rs:15-40
        if (w_i.z <= 0.0 || w_o.z <= 0.0) {
            a = 0.0, b = 0.0;
        } else {
            D3 refl = d3(-w_o.x, -w_o.y, w_o.z);
            double c = dot(w_i, refl);
            c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
            double theta = acos(fabs(c));
            double sigma = 0.05, two = 2.0;
            double rf = m.p1 * exp(-(theta * theta) / (two * sigma * sigma));
            a = (s * m.p0) * (1.0 - rf), b = rf;
        }
    } else { // smooth_transparent_dialectric.rs:74-89
        double eta1 = w_i.z >= 0.0 ? 1.0 : s, eta2 = w_i.z >= 0.0 ? s : 1.0;
        Fresnel f = fresnel(w_i, eta1, eta2);
        D3 dr = w_o - f.reflection_direction, dt = w_o - f.transmission_direction;
        b = 0.0;
        if (dot(dr, dr) < 0.0000000001) a = f.reflection_strength;
        else if (dot(dt, dt) < 0.0000000001) a = f.transmission_strength;
        else a = 0.0;
    }
}

} // namespace vrj
