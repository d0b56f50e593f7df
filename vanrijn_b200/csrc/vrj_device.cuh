// vrj_device.cuh -- device-side arithmetic of the render loop in the reference's operation order (compile with
// -fmad=false: rustc never contracts a*b+c).  Everything is a template over the real type R:
//   R = double : the reference's type (realtype.rs is binary64 everywhere) -- the parity path, bit-faithful;
//   R = float  : VRJ_PRECISION_F32_FAST, the same formulas in binary32 (SURVEY 8b/8d "f32 fast"); it has no
//                counterpart in the reference and is reported separately, never as a parity result.
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
template <typename R>
struct V3 {
    R x, y, z;
};
typedef V3<double> D3;
typedef V3<float> F3;
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
template <typename R>
__device__ __forceinline__ V3<R> v3(R x, R y, R z) { return V3<R>{x, y, z}; }
template <typename R, typename S>
__device__ __forceinline__ V3<R> convert(V3<S> a) { return V3<R>{(R)a.x, (R)a.y, (R)a.z}; }
template <typename R>
__device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return V3<R>{a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename R>
__device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return V3<R>{a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename R>
__device__ __forceinline__ V3<R> operator-(V3<R> a) { return V3<R>{-a.x, -a.y, -a.z}; }
template <typename R>
__device__ __forceinline__ V3<R> operator*(V3<R> a, R s) { return V3<R>{a.x * s, a.y * s, a.z * s}; }
// vec3.rs:76-82: products summed from 0.0 in x, y, z order
template <typename R>
__device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return ((R(0) + a.x * b.x) + a.y * b.y) + a.z * b.z; }
template <typename R>
__device__ __forceinline__ V3<R> cross(V3<R> a, V3<R> b) {
    return V3<R>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename R>
__device__ __forceinline__ R norm(V3<R> a) { return sqrt(dot(a, a)); }
// vec3.rs:103-110: multiply by 1/norm
template <typename R>
__device__ VRJ_INL_NORM V3<R> normalize(V3<R> a) {
    R inv = R(1) / norm(a);
    return V3<R>{a.x * inv, a.y * inv, a.z * inv};
}

template <typename R>
struct M3T {
    R e[3][3];
};
typedef M3T<double> M3;
// mat3.rs:72-94 (first_minor with mat2.rs:13-15, cofactor sign), written out per element
template <typename R>
__device__ __forceinline__ R minor2(R a, R b, R c, R d) { return a * d - b * c; }
template <typename R>
__device__ __forceinline__ R first_minor(const M3T<R> &m, int r, int c) {
    const int r0 = r == 0 ? 1 : 0, r1 = r == 2 ? 1 : 2;
    const int c0 = c == 0 ? 1 : 0, c1 = c == 2 ? 1 : 2;
    return minor2(m.e[r0][c0], m.e[r0][c1], m.e[r1][c0], m.e[r1][c1]);
}
template <typename R>
__device__ __forceinline__ R cofactor(const M3T<R> &m, int r, int c) {
    return (((r + c) & 1) ? R(-1) : R(1)) * first_minor(m, r, c);
}
// mat3.rs:106-109
template <typename R>
__device__ __forceinline__ R determinant(const M3T<R> &m) {
    return m.e[0][0] * first_minor(m, 0, 0) - m.e[0][1] * first_minor(m, 0, 1) + m.e[0][2] * first_minor(m, 0, 2);
}
// mat3.rs:111-118: transpose(cofactor matrix) * determinant (sic)
template <typename R>
__device__ __forceinline__ bool try_inverse(const M3T<R> &m, M3T<R> &out) {
    R det = determinant(m);
    if (det == R(0)) return false;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) out.e[i][j] = cofactor(m, j, i) * det;
    return true;
}
// mat3.rs:147-157
template <typename R>
__device__ __forceinline__ V3<R> mul(const M3T<R> &m, V3<R> v) {
    return V3<R>{dot(V3<R>{m.e[0][0], m.e[0][1], m.e[0][2]}, v), dot(V3<R>{m.e[1][0], m.e[1][1], m.e[1][2]}, v),
                 dot(V3<R>{m.e[2][0], m.e[2][1], m.e[2][2]}, v)};
}
// util/algebra_utils.rs:3-5 + mat3.rs:34-42
template <typename R>
__device__ __forceinline__ M3T<R> from_rows(V3<R> a, V3<R> b, V3<R> c) {
    M3T<R> m;
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
    // the same draws in the working precision (float: rand 0.7's f32 rules, 24 / 23 bits, from the same 64-bit word,
    // so a fast-mode path follows its parity twin as long as no discrete decision flips)
    template <typename R>
    __device__ __forceinline__ R uniform();
    template <typename R>
    __device__ __forceinline__ R uniform_open();
    // two consecutive Open01 draws.  EVEN: the caller knows the ordinal is even (both draws come from one Philox block), so
    // the block is formed at ONE place in the code instead of behind each draw's "is it cached?" test -- the Philox rounds
    // are the largest inlined piece of the shading kernels.  Same ordinals, same words, same values as two uniform_open calls.
    template <typename R, bool EVEN>
    __device__ __forceinline__ void open_pair(R &a, R &b);
};
template <>
__device__ __forceinline__ double Rng::uniform<double>() { return f64(); }
template <>
__device__ __forceinline__ double Rng::uniform_open<double>() { return open01(); }
template <>
__device__ __forceinline__ float Rng::uniform<float>() { return (float)(bits() >> 40) * (1.0f / 16777216.0f); }
template <>
__device__ __forceinline__ float Rng::uniform_open<float>() { return ((float)(bits() >> 41) + 0.5f) * (1.0f / 8388608.0f); }
template <typename R, bool EVEN>
__device__ __forceinline__ void Rng::open_pair(R &a, R &b) {
    if (EVEN) {
        block(ordinal >> 1);
        ordinal += 2;
        const uint64_t lo = ((uint64_t)w[1] << 32) | w[0], hi = ((uint64_t)w[3] << 32) | w[2];
        if (sizeof(R) == 8) {
            a = (R)(((double)(lo >> 12) + 0.5) * (1.0 / 4503599627370496.0)), b = (R)(((double)(hi >> 12) + 0.5) * (1.0 / 4503599627370496.0));
        } else {
            a = (R)(((float)(lo >> 41) + 0.5f) * (1.0f / 8388608.0f)), b = (R)(((float)(hi >> 41) + 0.5f) * (1.0f / 8388608.0f));
        }
    } else {
        a = uniform_open<R>(), b = uniform_open<R>();
    }
}

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

// spectrum.rs:50-79, `sample(i)` supplies sample i and `grid(i)` the wavelength of sample i, i / (n-1) * range + shortest
// (spectrum.rs:62,67).  In binary64 the grid comes from a table written at scene upload with exactly those operations
// (division, multiplication, addition of the same binary64 values round the same way on the host and here), which takes
// two of the lookup's four divisions off the shading path; `compute_grid` forms it in place (binary32 fast mode, lights).
template <typename R>
struct ComputedGrid {
    R shortest, range, nm1;
    __device__ __forceinline__ R operator()(uint32_t i) const { return (R)i / nm1 * range + shortest; }
};
template <typename R, typename F, typename G>
__device__ __forceinline__ R spectrum_lookup_grid(R shortest, R longest, uint32_t n, R wavelength, F sample, G grid) {
    if (wavelength < shortest || wavelength > longest) return R(0);
    R range = longest - shortest;
    R nm1 = (R)(n - 1);
    R fidx = nm1 * ((wavelength - shortest) / range);
    uint32_t before = (fidx != fidx || fidx < R(0)) ? 0u : (uint32_t)fidx;
    if (before > n - 1) before = n - 1; // cannot happen in binary64; guards binary32 rounding at the upper end
    R wl_before = grid(before);
    if (before == n - 1) return sample(before);
    R wl_after = grid(before + 1);
    R delta = wl_after - wl_before;
    R ratio = (wavelength - wl_before) / delta;
    return sample(before) * (R(1) - ratio) + sample(before + 1) * ratio;
}
template <typename R, typename F>
__device__ __forceinline__ R spectrum_lookup(R shortest, R longest, uint32_t n, R wavelength, F sample) {
    return spectrum_lookup_grid<R>(shortest, longest, n, wavelength, sample, ComputedGrid<R>{shortest, longest - shortest, (R)(n - 1)});
}
// `grids`: the table parallel to `samples` (same indices), binary64 only
__device__ __forceinline__ double spectrum_intensity(const SpectrumDev *__restrict__ spectra, const double *__restrict__ samples,
                                                     const double *__restrict__ grids, uint32_t id, double wavelength) {
    SpectrumDev s = spectra[id];
    const double *p = samples + s.first, *g = grids + s.first;
    return spectrum_lookup_grid<double>(s.shortest, s.longest, s.n, wavelength, [p](uint32_t i) { return __ldg(p + i); },
                                        [g](uint32_t i) { return __ldg(g + i); });
}
__device__ __forceinline__ float spectrum_intensity(const SpectrumDev *__restrict__ spectra, const double *__restrict__ samples,
                                                    const double *__restrict__, uint32_t id, float wavelength) {
    SpectrumDev s = spectra[id];
    const double *p = samples + s.first;
    return spectrum_lookup<float>((float)s.shortest, (float)s.longest, s.n, wavelength, [p](uint32_t i) { return (float)__ldg(p + i); });
}
// the sample wavelengths of the 32-entry rgb basis spectra (spectrum.rs:182-419 span 380..720 nm): i / 31 * 340 + 380,
// folded by the host compiler in IEEE binary64 (round to nearest), i.e. the values the expression gives at run time
#define VRJ_RGB_GRID(i) ((double)(i) / 31.0 * 340.0 + 380.0)
__device__ const double g_rgb_grid[32] = {
    VRJ_RGB_GRID(0),  VRJ_RGB_GRID(1),  VRJ_RGB_GRID(2),  VRJ_RGB_GRID(3),  VRJ_RGB_GRID(4),  VRJ_RGB_GRID(5),  VRJ_RGB_GRID(6),  VRJ_RGB_GRID(7),
    VRJ_RGB_GRID(8),  VRJ_RGB_GRID(9),  VRJ_RGB_GRID(10), VRJ_RGB_GRID(11), VRJ_RGB_GRID(12), VRJ_RGB_GRID(13), VRJ_RGB_GRID(14), VRJ_RGB_GRID(15),
    VRJ_RGB_GRID(16), VRJ_RGB_GRID(17), VRJ_RGB_GRID(18), VRJ_RGB_GRID(19), VRJ_RGB_GRID(20), VRJ_RGB_GRID(21), VRJ_RGB_GRID(22), VRJ_RGB_GRID(23),
    VRJ_RGB_GRID(24), VRJ_RGB_GRID(25), VRJ_RGB_GRID(26), VRJ_RGB_GRID(27), VRJ_RGB_GRID(28), VRJ_RGB_GRID(29), VRJ_RGB_GRID(30), VRJ_RGB_GRID(31)};
// spectrum.rs:81-165 evaluated lazily: only the (at most two) samples the lookup touches are formed
template <typename R>
__device__ __forceinline__ void rgb_basis_split(R r, R g, R b, int &second, int &third, R &c0, R &c1, R &c2) {
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
}
__device__ __forceinline__ double rgb_reflection_intensity(double r, double g, double b, double wavelength) {
    int second, third;
    double c0, c1, c2;
    rgb_basis_split(r, g, b, second, third, c0, c1, c2);
    return spectrum_lookup_grid<double>(380.0, 720.0, 32u, wavelength,
                                        [=](uint32_t i) { return c0 * g_rgb_basis[B_WHITE][i] + c1 * g_rgb_basis[second][i] + c2 * g_rgb_basis[third][i]; },
                                        [](uint32_t i) { return g_rgb_grid[i]; });
}
__device__ __forceinline__ float rgb_reflection_intensity(float r, float g, float b, float wavelength) {
    int second, third;
    float c0, c1, c2;
    rgb_basis_split(r, g, b, second, third, c0, c1, c2);
    return spectrum_lookup<float>(380.f, 720.f, 32u, wavelength, [=](uint32_t i) {
        return c0 * (float)g_rgb_basis[B_WHITE][i] + c1 * (float)g_rgb_basis[second][i] + c2 * (float)g_rgb_basis[third][i];
    });
}

// colour/colour_xyz.rs:86-103
template <typename R>
__device__ __forceinline__ R gaussian(R w, R alpha, R mu, R s1, R s2) {
    R sigma = w < mu ? s1 : s2;
    R denominator = R(2) * (sigma * sigma);
    return alpha * exp(-((w - mu) * (w - mu)) / denominator);
}
template <typename R>
__device__ __forceinline__ V3<R> cmf(R w) {
    R x = gaussian<R>(w, R(1.056), R(599.8), R(37.9), R(31.0)) + gaussian<R>(w, R(0.362), R(442.0), R(16.0), R(26.7)) +
          gaussian<R>(w, R(-0.065), R(501.1), R(20.4), R(26.2));
    R y = gaussian<R>(w, R(0.821), R(568.8), R(46.9), R(40.5)) + gaussian<R>(w, R(0.286), R(530.9), R(16.3), R(31.1));
    R z = gaussian<R>(w, R(1.217), R(437.0), R(11.8), R(36.0)) + gaussian<R>(w, R(0.681), R(459.0), R(26.0), R(13.8));
    return V3<R>{x, y, z};
}

// ------------------------------------------------------------------------------------------
// raycasting: the primitives' own intersection arithmetic (exact in the reference's sense when R = double)
template <typename R>
struct HitFrameT {
    R distance;
    V3<R> location, normal, tangent, cotangent, retro;
    uint32_t material;
};
typedef HitFrameT<double> HitFrame;

// Per-ray constants of Triangle::intersect: permutation (triangle.rs:108-122, SIGNED largest
// component last, cyclic permutations only) and shear (triangle.rs:133-135).
template <typename R>
struct TriRayT {
    V3<R> o;
    R sx, sy, pdz;
    int perm; // 0: (x,y,z)  1: (y,z,x)  2: (z,x,y)
};
typedef TriRayT<double> TriRay;
template <typename R>
__device__ __forceinline__ V3<R> permute(V3<R> v, int perm) {
    return perm == 0 ? v : (perm == 1 ? V3<R>{v.y, v.z, v.x} : V3<R>{v.z, v.x, v.y});
}
template <typename R>
__device__ __forceinline__ TriRayT<R> tri_ray(V3<R> o, V3<R> d) {
    TriRayT<R> r;
    r.o = o;
    if (d.x > d.y) r.perm = (d.z > d.x) ? 0 : 1;
    else r.perm = (d.z > d.y) ? 0 : 2;
    V3<R> pd = permute(d, r.perm);
    r.sx = -pd.x / pd.z, r.sy = -pd.y / pd.z, r.pdz = pd.z;
    return r;
}
template <typename R>
__device__ __forceinline__ R edge_fn(V3<R> a, V3<R> b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ bool sign_bit(double v) { return __double2hiint(v) < 0; }
__device__ __forceinline__ bool sign_bit(float v) { return __float_as_int(v) < 0; }

// triangle.rs:35-72: returns true and the barycentrics + distance when the ray hits
template <typename R>
__device__ VRJ_INL_PRIM bool triangle_test(const TriRayT<R> &r, V3<R> v0, V3<R> v1, V3<R> v2, R &distance, R &b0, R &b1, R &b2,
                                           V3<R> &location) {
    V3<R> p0 = permute(v0 - r.o, r.perm), p1 = permute(v1 - r.o, r.perm), p2 = permute(v2 - r.o, r.perm);
    V3<R> t0 = V3<R>{p0.x + r.sx * p0.z, p0.y + r.sy * p0.z, p0.z};
    V3<R> t1 = V3<R>{p1.x + r.sx * p1.z, p1.y + r.sy * p1.z, p1.z};
    V3<R> t2 = V3<R>{p2.x + r.sx * p2.z, p2.y + r.sy * p2.z, p2.z};
    R e0 = edge_fn(t1, t2), e1 = edge_fn(t2, t0), e2 = edge_fn(t0, t1);
    // sign BITS, so +-0 matter (triangle.rs:52-53)
    int neg = (int)sign_bit(e0) + (int)sign_bit(e1) + (int)sign_bit(e2);
    if (neg != 0 && neg != 3) return false;
    R a0 = fabs(e0), a1 = fabs(e1), a2 = fabs(e2);
    R inv = R(1) / (((R(0) + a0) + a1) + a2);
    b0 = a0 * inv, b1 = a1 * inv, b2 = a2 * inv;
    R tz = ((R(0) + t0.z * b0) + t1.z * b1) + t2.z * b2;
    if (sign_bit(tz) != sign_bit(r.pdz)) return false;
    location = ((V3<R>{R(0), R(0), R(0)} + v0 * b0) + v1 * b1) + v2 * b2;
    distance = norm(r.o - location);
    return true;
}

struct SphereDev {
    double cx, cy, cz, radius;
    uint32_t material, pad;
};
// sphere.rs:39-75 (distance only)
template <typename R>
__device__ VRJ_INL_PRIM bool sphere_test(const SphereDev &s, V3<R> o, V3<R> d, R &distance) {
    V3<R> c = V3<R>{(R)s.cx, (R)s.cy, (R)s.cz};
    const R radius = (R)s.radius;
    R a = ((R(0) + d.x * d.x) + d.y * d.y) + d.z * d.z;
    R b = ((R(0) + (o.x * d.x - c.x * d.x) * R(2)) + (o.y * d.y - c.y * d.y) * R(2)) + (o.z * d.z - c.z * d.z) * R(2);
    R cc = (((R(0) + ((o.x * o.x + c.x * c.x) - c.x * o.x * R(2))) + ((o.y * o.y + c.y * c.y) - c.y * o.y * R(2))) +
            ((o.z * o.z + c.z * c.z) - c.z * o.z * R(2))) -
           radius * radius;
    R delta_squared = b * b - R(4) * a * cc;
    if (delta_squared < R(0)) return false;
    R delta = sqrt(delta_squared);
    R one_over_2a = R(1) / (R(2) * a);
    R t1 = (-b - delta) * one_over_2a;
    R t2 = (-b + delta) * one_over_2a;
    distance = (t1 < R(0) || (t2 >= R(0) && t1 >= t2)) ? t2 : t1;
    return !(distance <= R(0));
}
// sphere.rs:76-90
template <typename R>
__device__ __forceinline__ void sphere_frame(const SphereDev &s, V3<R> o, V3<R> d, R distance, HitFrameT<R> &h) {
    h.distance = distance;
    h.location = o + d * distance;
    h.normal = normalize(h.location - V3<R>{(R)s.cx, (R)s.cy, (R)s.cz});
    h.tangent = normalize(cross(h.normal, V3<R>{R(0), R(0), R(1)}));
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
template <typename R>
__device__ __forceinline__ bool plane_test(const PlaneDev &p, V3<R> o, V3<R> d, R &t) {
    V3<R> n = V3<R>{(R)p.n[0], (R)p.n[1], (R)p.n[2]};
    R d_dot_n = dot(d, n);
    V3<R> point_on_plane = n * (R)p.distance;
    R num = dot(point_on_plane - o, n);
    if (d_dot_n == R(0) && num != R(0)) return false;
    t = num / d_dot_n;
    return !(t < R(0));
}
// plane.rs:64-73
template <typename R>
__device__ __forceinline__ void plane_frame(const PlaneDev &p, V3<R> o, V3<R> d, R t, HitFrameT<R> &h) {
    h.distance = t;
    h.location = o + d * t;
    h.normal = V3<R>{(R)p.n[0], (R)p.n[1], (R)p.n[2]};
    h.tangent = V3<R>{(R)p.t[0], (R)p.t[1], (R)p.t[2]};
    h.cotangent = V3<R>{(R)p.c[0], (R)p.c[1], (R)p.c[2]};
    h.retro = -d;
    h.material = p.material;
}

// ------------------------------------------------------------------------------------------
// materials
struct MaterialDev {
    uint32_t kind, spectrum;
    double p0, p1, p2;
};
template <typename R>
struct FresnelT {
    V3<R> reflection_direction, transmission_direction;
    R reflection_strength, transmission_strength;
};
// smooth_transparent_dialectric.rs:15-60
template <typename R>
__device__ __forceinline__ FresnelT<R> fresnel(V3<R> w_i, R eta1, R eta2) {
    V3<R> normal = w_i.z > R(0) ? V3<R>{R(0), R(0), R(1)} : -V3<R>{R(0), R(0), R(1)};
    FresnelT<R> f;
    f.reflection_direction = V3<R>{-w_i.x, -w_i.y, w_i.z};
    R r = eta1 / eta2;
    R cos1 = dot(normal, w_i);
    R cos2sq = R(1) - r * r * (R(1) - cos1 * cos1);
    if (cos2sq >= R(0)) {
        R cos2 = sqrt(cos2sq);
        R rpar = (eta1 * cos2 - eta2 * cos1) / (eta1 * cos2 + eta2 * cos1);
        R rperp = (eta1 * cos1 - eta2 * cos2) / (eta1 * cos1 + eta2 * cos2);
        f.reflection_strength = R(0.5) * (rpar * rpar + rperp * rperp);
        f.transmission_direction = normalize((w_i * (-r)) + (normal * (r * cos1 - cos2)));
        f.transmission_strength = R(1) - f.reflection_strength;
    } else {
        f.reflection_strength = R(1);
        f.transmission_strength = R(0);
        f.transmission_direction = V3<R>{R(0), R(0), R(0)};
    }
    if (w_i.z < R(0)) {
        f.reflection_direction.z *= R(-1);
        f.transmission_direction.z *= R(-1);
    }
    return f;
}

#define VRJ_PI 3.14159265358979323846264338327950288

// MM: bit k set = material kind k may occur in the scene (VrjScene knows its materials).  A kernel instantiated for the kinds
// a scene really has carries no code for the others (Phong's pow, the reflective lobe's acos / exp, the dielectric's Fresnel
// terms, the cosine-hemisphere sampler's sin / cos): less to fetch through the 32 KB instruction cache of an SM.
#define VRJ_MM_ALL 15
template <int MM>
__device__ __forceinline__ bool is_kind(uint32_t kind, int k) { return (MM & (1 << k)) != 0 && (MM == (1 << k) || kind == (uint32_t)k); }

// Material::sample: returns direction (BSDF space) and pdf, consuming draws from rng
template <int MM, typename R>
__device__ __forceinline__ void material_sample(const MaterialDev &m, R eta_or_zero, V3<R> w_i, Rng &rng, V3<R> &dir, R &pdf) {
    const R pi = R(VRJ_PI);
    if (is_kind<MM>(m.kind, 0)) { // lambertian_material.rs:36-59 (rejection in the unit disc)
        // in a scene whose materials are all Lambertian every material draw comes in pairs, so the ordinal stays even
        R x, y;
        do {
            R u, v;
            rng.open_pair<R, MM == 1>(u, v);
            x = R(2) * u - R(1), y = R(2) * v - R(1);
        } while (((R(0) + x * x) + y * y) + R(0) * R(0) > R(1));
        R z = fmax(sqrt(R(1) - x * x - y * y), R(0));
        R cos_theta = ((R(0) + x * R(0)) + y * R(0)) + z * R(1);
        R sin_theta = sqrt(R(1) - cos_theta * cos_theta);
        dir = normalize(V3<R>{x, y, z});
        pdf = (cos_theta * sin_theta) / pi;
    } else if (is_kind<MM>(m.kind, 2)) { // reflective_material.rs:42-47
        dir = V3<R>{-w_i.x, -w_i.y, w_i.z};
        pdf = R(1);
    } else if (is_kind<MM>(m.kind, 3)) { // smooth_transparent_dialectric.rs:91-114
        R eta1 = w_i.z >= R(0) ? R(1) : eta_or_zero, eta2 = w_i.z >= R(0) ? eta_or_zero : R(1);
        FresnelT<R> f = fresnel(w_i, eta1, eta2);
        pdf = R(0.5);
        if (f.transmission_strength <= R(0.0000000001)) dir = f.reflection_direction;
        else if (f.reflection_strength <= R(0.0000000001) || rng.boolean()) dir = f.transmission_direction;
        else dir = f.reflection_direction;
    } else if (MM & 2) { // Phong: materials/mod.rs:28-33 -> cosine_weighted_hemisphere.rs:19-33, unit_disc.rs:27-44, uniform_square.rs:20-25
        R sx = R(-1) + rng.uniform_open<R>() * R(2);
        R sy = R(-1) + rng.uniform_open<R>() * R(2);
        R dx, dy;
        if (sx == R(0) && sy == R(0)) {
            dx = sx, dy = sy;
        } else {
            R radius, angle;
            if (fabs(sx) > fabs(sy)) radius = sx, angle = (pi / R(4)) * sy / sx;
            else radius = sy, angle = pi / R(2) - (pi / R(4)) * sx / sy;
            dx = cos(angle) * radius, dy = sin(angle) * radius;
        }
        R z = sqrt(fmax(R(0), R(1) - dx * dx - dy * dy));
        dir = V3<R>{dx, dy, z};
        pdf = sqrt(dx * dx + dy * dy) / pi;
    } else { // a kind the scene was said not to contain: unreachable (vrj_scene_create forms the mask from the materials)
        dir = V3<R>{R(0), R(0), R(1)}, pdf = R(0);
    }
}

// Material::bsdf as an affine map of the incoming intensity: out = a * in + b.
// `s` is the material spectrum at the photon's wavelength (colour, or eta for the dielectric).
template <int MM, typename R>
__device__ __forceinline__ void material_bsdf_affine(const MaterialDev &m, R s, V3<R> w_o, V3<R> w_i, R &a, R &b) {
    const R p0 = (R)m.p0, p1 = (R)m.p1, p2 = (R)m.p2;
    if (is_kind<MM>(m.kind, 0)) { // lambertian_material.rs:27-34
        a = s * p0, b = R(0);
    } else if (is_kind<MM>(m.kind, 1)) { // phong_material.rs:16-36
        if (w_i.z < R(0) || w_o.z < R(0)) {
            a = R(0), b = R(0);
        } else {
            V3<R> refl = V3<R>{-w_i.x, -w_i.y, w_i.z};
            a = s * p0;
            b = pow(fabs(dot(w_o, refl)), p2) * (p1 / dot(w_i, V3<R>{R(0), R(0), R(1)}));
        }
    } else if (is_kind<MM>(m.kind, 2)) { // reflective_material.rs:15-40
        if (w_i.z <= R(0) || w_o.z <= R(0)) {
            a = R(0), b = R(0);
        } else {
            V3<R> refl = V3<R>{-w_o.x, -w_o.y, w_o.z};
            R c = dot(w_i, refl);
            c = c < R(0) ? R(0) : (c > R(1) ? R(1) : c);
            R theta = acos(fabs(c));
            R sigma = R(0.05), two = R(2);
            R rf = p1 * exp(-(theta * theta) / (two * sigma * sigma));
            a = (s * p0) * (R(1) - rf), b = rf;
        }
    } else if (MM & 8) { // smooth_transparent_dialectric.rs:74-89
        R eta1 = w_i.z >= R(0) ? R(1) : s, eta2 = w_i.z >= R(0) ? s : R(1);
        FresnelT<R> f = fresnel(w_i, eta1, eta2);
        V3<R> dr = w_o - f.reflection_direction, dt = w_o - f.transmission_direction;
        b = R(0);
        if (dot(dr, dr) < R(0.0000000001)) a = f.reflection_strength;
        else if (dot(dt, dt) < R(0.0000000001)) a = f.transmission_strength;
        else a = R(0);
    } else {
        a = R(0), b = R(0);
    }
}

} // namespace vrj
