#!/usr/bin/env python
"""second_source.py -- an INDEPENDENT restatement of the reference's per-sample algorithm, written in plain Python
straight from the Rust lines it cites (paths relative to /root/reference/src/), to pin the C++ oracle's rows that the
reference itself holds no vectors for: both integrators, the four materials, Fresnel, BVH build + traversal with
ties, the OBJ loader and the camera.  It imports NOTHING from oracle/, tests/oraclelib.py or the product; the only
thing it shares with them is the specification of the counter-based generator that replaces `rand` (SURVEY 8a row
27: Philox-4x32-10 as published by Random123, rand 0.7's Standard / Open01 / bool conversions), restated here too.

    python tests/golden/second_source.py            writes tests/golden/second_source.json
    tests/test_second_source.py                     asserts that the oracle reproduces every vector

Python floats are IEEE binary64 and + - * / math.sqrt are correctly rounded, exactly like the Rust f64 code; acos /
exp / pow / sin / cos go through the platform libm, as rustc's do, so vectors involving them are compared at 1e-12.

The rgb basis spectra (spectrum.rs:182-419) are read from the reference source when this script runs (it is only ever
run in the build container, where /root/reference exists); the vectors it writes are self-contained.
"""
import json
import math
import os
import re
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
INF = float("inf")

# ------------------------------------------------------------------ math/vec3.rs, math/mat3.rs


def dot(a, b):  # vec3.rs:76-82: products summed from 0.0 in x, y, z order
    s = 0.0
    for x, y in zip(a, b):
        s = s + x * y
    return s


def cross(a, b):  # vec3.rs:84-89
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def neg(a):
    return (-a[0], -a[1], -a[2])


def scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def norm(a):  # vec3.rs:95-101
    return math.sqrt(dot(a, a))


def normalize(a):  # vec3.rs:103-110: multiply by 1 / norm (a zero vector gives NaNs, as in IEEE arithmetic: no exception)
    n = norm(a)
    inv = 1.0 / n if n != 0.0 else INF
    return tuple(x * inv if not (inv == INF and x == 0.0) else float("nan") for x in a)


def smallest_coord(v):  # vec3.rs:112-127
    x, y, z = abs(v[0]), abs(v[1]), abs(v[2])
    if x < y:
        return 0 if x < z else 2
    return 1 if y < z else 2


def mat_mul_vec(m, v):  # mat3.rs:147-157
    return (dot(m[0], v), dot(m[1], v), dot(m[2], v))


def first_minor(m, row, col):  # mat3.rs:72-90 with mat2.rs:13-15
    e = [[m[i][j] for j in range(3) if j != col] for i in range(3) if i != row]
    return e[0][0] * e[1][1] - e[0][1] * e[1][0]


def cofactor(m, row, col):  # mat3.rs:92-94
    return float((-1) ** (row + col)) * first_minor(m, row, col)


def determinant(m):  # mat3.rs:106-109
    return m[0][0] * first_minor(m, 0, 0) - m[0][1] * first_minor(m, 0, 1) + m[0][2] * first_minor(m, 0, 2)


def try_inverse(m):  # mat3.rs:111-118: transpose(cofactor matrix) * determinant (sic)
    det = determinant(m)
    if det == 0.0:
        return None
    cof = [[cofactor(m, i, j) for j in range(3)] for i in range(3)]
    return [tuple(cof[j][i] * det for j in range(3)) for i in range(3)]


# ------------------------------------------------------------------ the generator replacing `rand`
M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):  # Random123 philox.h: 10 rounds, multipliers D2511F53 / CD9E8D57, Weyl 9E3779B9 / BB67AE85
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for r in range(10):
        if r:
            k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & M32, p1 >> 32, p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return c0, c1, c2, c3


class Rng:
    """draw `ordinal` of (seed, pixel, sample): block = ordinal // 2 -> counter (block, pixel, sample lo, sample hi), key = seed;
    even ordinals take words 0,1 and odd ones words 2,3 as one 64-bit value (low word first)."""

    def __init__(self, seed, pixel, sample, ordinal):
        self.seed, self.pixel, self.sample, self.ordinal = seed, pixel, sample, ordinal

    def bits(self):
        w = philox4x32_10((self.ordinal >> 1, self.pixel, self.sample & M32, self.sample >> 32), (self.seed & M32, self.seed >> 32))
        lo, hi = (w[2], w[3]) if self.ordinal & 1 else (w[0], w[1])
        self.ordinal += 1
        return (hi << 32) | lo

    def f64(self):  # rand 0.7 Standard for f64: 53 random bits -> [0, 1)
        return float(self.bits() >> 11) * (1.0 / 9007199254740992.0)

    def open01(self):  # rand 0.7 Open01 for f64: 52 random bits -> (0, 1)
        return (float(self.bits() >> 12) + 0.5) * (1.0 / 4503599627370496.0)

    def boolean(self):  # rand 0.7 Standard for bool: the most significant bit
        return (self.bits() >> 63) != 0


# ------------------------------------------------------------------ colour/spectrum.rs, colour/photon.rs
def read_rgb_basis():
    text = open(os.path.join(REF, "colour", "spectrum.rs")).read()
    body = text[text.index("pub mod reflection"):]
    out = {}
    for name in ("WHITE", "CYAN", "MAGENTA", "YELLOW", "RED", "GREEN", "BLUE"):
        m = re.search(r"pub const %s: \[f64; (\d+)\] = \[(.*?)\];" % name, body, re.S)
        vals = [float(x) for x in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", m.group(2))]
        assert len(vals) == int(m.group(1)) == 32, (name, len(vals))
        out[name] = vals
    return out


class Spectrum:
    def __init__(self, shortest, longest, samples):
        self.shortest, self.longest, self.samples = shortest, longest, list(samples)

    def intensity_at_wavelength(self, w):  # spectrum.rs:50-79
        if w < self.shortest or w > self.longest:
            return 0.0
        n = len(self.samples)
        rng = self.longest - self.shortest
        f = float(n - 1) * ((w - self.shortest) / rng)
        before = 0 if (f != f or f < 0.0) else int(f)  # `as usize`: saturating, NaN -> 0
        wl_before = float(before) / float(n - 1) * rng + self.shortest
        if before == n - 1:
            return self.samples[before]
        wl_after = float(before + 1) / float(n - 1) * rng + self.shortest
        delta = wl_after - wl_before
        ratio = (w - wl_before) / delta
        return self.samples[before] * (1.0 - ratio) + self.samples[before + 1] * ratio


def grey(b):  # spectrum.rs:21-27
    return Spectrum(380.0, 740.0, [b, b])


def diamond():  # spectrum.rs:29-49
    return Spectrum(326.27, 774.9, [2.505813241, 2.487866556, 2.473323675, 2.464986815, 2.455051934, 2.441251728,
                                    2.431478974, 2.427076431, 2.420857286, 2.411429037, 2.406543164, 2.406202402])


def reflection_from_linear_rgb(basis, r, g, b):  # spectrum.rs:81-165
    W = basis["WHITE"]
    if r <= g and r <= b:
        if g <= b:
            s = [r * w + (g - r) * c + (b - g) * x for w, c, x in zip(W, basis["CYAN"], basis["BLUE"])]
        else:
            s = [r * w + (b - r) * c + (g - b) * x for w, c, x in zip(W, basis["CYAN"], basis["GREEN"])]
    elif g <= r and g < b:
        if r <= b:
            s = [g * w + (r - g) * c + (b - r) * x for w, c, x in zip(W, basis["MAGENTA"], basis["BLUE"])]
        else:
            s = [g * w + (b - g) * c + (r - b) * x for w, c, x in zip(W, basis["MAGENTA"], basis["RED"])]
    else:
        if r <= g:
            s = [b * w + (r - b) * c + (g - r) * x for w, c, x in zip(W, basis["YELLOW"], basis["GREEN"])]
        else:
            s = [b * w + (g - b) * c + (r - g) * x for w, c, x in zip(W, basis["YELLOW"], basis["RED"])]
    return Spectrum(380.0, 720.0, s)


# ------------------------------------------------------------------ materials/*.rs   (a photon is (wavelength, intensity))
PI = math.pi


def fresnel(w_i, eta1, eta2):  # smooth_transparent_dialectric.rs:15-60
    normal = (0.0, 0.0, 1.0) if w_i[2] > 0.0 else (-0.0, -0.0, -1.0)
    refl = (-w_i[0], -w_i[1], w_i[2])
    r = eta1 / eta2
    cos1 = dot(normal, w_i)
    cos2sq = 1.0 - r * r * (1.0 - cos1 * cos1)
    if cos2sq >= 0.0:
        cos2 = math.sqrt(cos2sq)
        rpar = (eta1 * cos2 - eta2 * cos1) / (eta1 * cos2 + eta2 * cos1)
        rperp = (eta1 * cos1 - eta2 * cos2) / (eta1 * cos1 + eta2 * cos2)
        rs = 0.5 * (rpar * rpar + rperp * rperp)
        td = normalize(add(scale(w_i, -r), scale(normal, r * cos1 - cos2)))
        ts = 1.0 - rs
    else:
        rs, ts, td = 1.0, 0.0, (0.0, 0.0, 0.0)
    if w_i[2] < 0.0:
        refl = (refl[0], refl[1], refl[2] * -1.0)
        td = (td[0], td[1], td[2] * -1.0)
    return refl, rs, td, ts


class Lambertian:  # lambertian_material.rs
    def __init__(self, colour, diffuse):
        self.colour, self.diffuse = colour, diffuse

    def bsdf(self, w_o, w_i, photon):  # :27-34
        return (photon[0], photon[1] * self.colour.intensity_at_wavelength(photon[0]) * self.diffuse)

    def sample(self, w_i, photon, rng):  # :36-59
        x, y = 2.0 * rng.open01() - 1.0, 2.0 * rng.open01() - 1.0
        while dot((x, y, 0.0), (x, y, 0.0)) > 1.0:
            x, y = 2.0 * rng.open01() - 1.0, 2.0 * rng.open01() - 1.0
        z = max(math.sqrt(1.0 - x * x - y * y), 0.0)
        cos_theta = dot((x, y, z), (0.0, 0.0, 1.0))
        sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
        return normalize((x, y, z)), (cos_theta * sin_theta) / PI


class Phong:  # phong_material.rs; sample = the trait default (materials/mod.rs:28-33)
    def __init__(self, colour, diffuse, specular, smoothness):
        self.colour, self.diffuse, self.specular, self.smoothness = colour, diffuse, specular, smoothness

    def bsdf(self, w_o, w_i, photon):  # :16-36
        if w_i[2] < 0.0 or w_o[2] < 0.0:
            return (photon[0], 0.0)
        refl = (-w_i[0], -w_i[1], w_i[2])
        inten = (photon[1] * self.colour.intensity_at_wavelength(photon[0])) * self.diffuse + \
            math.pow(abs(dot(w_o, refl)), self.smoothness) * (self.specular / dot(w_i, (0.0, 0.0, 1.0)))
        return (photon[0], inten)

    def sample(self, w_i, photon, rng):  # cosine_weighted_hemisphere.rs:19-33, unit_disc.rs:27-44, uniform_square.rs:20-25
        sx, sy = -1.0 + rng.open01() * 2.0, -1.0 + rng.open01() * 2.0
        if sx == 0.0 and sy == 0.0:
            dx, dy = sx, sy
        else:
            if abs(sx) > abs(sy):
                radius, angle = sx, (PI / 4.0) * sy / sx
            else:
                radius, angle = sy, PI / 2.0 - (PI / 4.0) * sx / sy
            dx, dy = math.cos(angle) * radius, math.sin(angle) * radius
        z = math.sqrt(max(0.0, 1.0 - dx * dx - dy * dy))
        return (dx, dy, z), math.sqrt(dx * dx + dy * dy) / PI


class Reflective:  # reflective_material.rs
    def __init__(self, colour, diffuse, reflection):
        self.colour, self.diffuse, self.reflection = colour, diffuse, reflection

    def bsdf(self, w_o, w_i, photon):  # :15-40
        if w_i[2] <= 0.0 or w_o[2] <= 0.0:
            return (photon[0], 0.0)
        refl = (-w_o[0], -w_o[1], w_o[2])
        out = photon[1] * self.colour.intensity_at_wavelength(photon[0])
        out *= self.diffuse
        sigma, two = 0.05, 2.0
        c = dot(w_i, refl)
        c = 0.0 if c < 0.0 else (1.0 if c > 1.0 else c)  # f64::clamp(0, 1)
        theta = math.acos(abs(c))
        rf = self.reflection * math.exp(-(math.pow(theta, two)) / (two * sigma * sigma))
        return (photon[0], out * (1.0 - rf) + rf)

    def sample(self, w_o, photon, rng):  # :42-47
        return (-w_o[0], -w_o[1], w_o[2]), 1.0


class Dielectric:  # smooth_transparent_dialectric.rs
    def __init__(self, eta):
        self.eta = eta

    def _etas(self, w_i, photon):
        e = self.eta.intensity_at_wavelength(photon[0])
        return (1.0, e) if w_i[2] >= 0.0 else (e, 1.0)

    def bsdf(self, w_o, w_i, photon):  # :74-89
        eta1, eta2 = self._etas(w_i, photon)
        refl, rs, td, ts = fresnel(w_i, eta1, eta2)
        d = sub(w_o, refl)
        if dot(d, d) < 0.0000000001:
            return (photon[0], photon[1] * rs)
        d = sub(w_o, td)
        if dot(d, d) < 0.0000000001:
            return (photon[0], photon[1] * ts)
        return (photon[0], 0.0)

    def sample(self, w_i, photon, rng):  # :91-114 (`||` short-circuits: the bool is drawn only when both strengths are non-zero)
        eta1, eta2 = self._etas(w_i, photon)
        refl, rs, td, ts = fresnel(w_i, eta1, eta2)
        if ts <= 0.0000000001:
            return refl, 0.5
        if rs <= 0.0000000001 or rng.boolean():
            return td, 0.5
        return refl, 0.5


# ------------------------------------------------------------------ raycasting/*.rs   (a hit is a dict like IntersectionInfo)
def ray_new(origin, direction):  # raycasting/mod.rs:41-46
    return (origin, normalize(direction))


def point_at(ray, t):
    return add(ray[0], scale(ray[1], t))


def ray_bias(ray, amount):  # raycasting/mod.rs:58-60
    return ray_new(add(ray[0], scale(ray[1], amount)), ray[1])


class Sphere:
    def __init__(self, centre, radius, material):
        self.centre, self.radius, self.material = centre, radius, material

    def intersect(self, ray):  # sphere.rs:39-90
        o, d, c = ray[0], ray[1], self.centre
        a = 0.0
        for k in range(3):
            a = a + d[k] * d[k]
        b = 0.0
        for k in range(3):
            b = b + (o[k] * d[k] - c[k] * d[k]) * 2.0
        cc = 0.0
        for k in range(3):
            cc = cc + ((o[k] * o[k] + c[k] * c[k]) - c[k] * o[k] * 2.0)
        cc = cc - self.radius * self.radius
        delta_squared = b * b - 4.0 * a * cc
        if delta_squared < 0.0:
            return None
        delta = math.sqrt(delta_squared)
        one_over_2a = 1.0 / (2.0 * a)
        t1, t2 = (-b - delta) * one_over_2a, (-b + delta) * one_over_2a
        distance = t2 if (t1 < 0.0 or (t2 >= 0.0 and t1 >= t2)) else t1
        if distance <= 0.0:
            return None
        location = point_at(ray, distance)
        normal = normalize(sub(location, c))
        tangent = normalize(cross(normal, (0.0, 0.0, 1.0)))
        return dict(distance=distance, location=location, normal=normal, tangent=tangent, cotangent=cross(normal, tangent),
                    retro=neg(d), material=self.material, what=("sphere",))


class Plane:
    def __init__(self, normal, distance, material):  # plane.rs:17-32
        self.normal = normalize(normal)
        axis = [0.0, 0.0, 0.0]
        axis[smallest_coord(self.normal)] = 1.0
        self.cotangent = normalize(cross(self.normal, tuple(axis)))
        self.tangent = cross(self.normal, self.cotangent)
        self.distance, self.material = distance, material

    def intersect(self, ray):  # plane.rs:48-75
        d_dot_n = dot(ray[1], self.normal)
        num = dot(sub(scale(self.normal, self.distance), ray[0]), self.normal)
        if d_dot_n == 0.0 and num != 0.0:
            return None
        t = num / d_dot_n if d_dot_n != 0.0 else (float("nan") if num == 0.0 else math.copysign(INF, num))
        if t < 0.0:
            return None
        return dict(distance=t, location=point_at(ray, t), normal=self.normal, tangent=self.tangent, cotangent=self.cotangent,
                    retro=neg(ray[1]), material=self.material, what=("plane",))


def sign_positive(x):  # f64::is_sign_positive: the sign BIT
    return math.copysign(1.0, x) > 0.0


class Triangle:
    def __init__(self, vertices, normals, material, prim_id=0):
        self.v, self.n, self.material, self.prim_id = [tuple(x) for x in vertices], [tuple(x) for x in normals], material, prim_id

    def box(self):  # BoundingBox::from_points (util/axis_aligned_bounding_box.rs:40-47)
        return tuple((min(p[k] for p in self.v), max(p[k] for p in self.v)) for k in range(3))

    def intersect(self, ray):  # triangle.rs:35-100
        o, d = ray
        if d[0] > d[1]:  # :108-122 signed largest component last, cyclic permutations only
            idx = (0, 1, 2) if d[2] > d[0] else (1, 2, 0)
        else:
            idx = (0, 1, 2) if d[2] > d[1] else (2, 0, 1)
        pd = (d[idx[0]], d[idx[1]], d[idx[2]])
        sx, sy = fdiv(-pd[0], pd[2]), fdiv(-pd[1], pd[2])  # :133-135
        tv = []
        for v in self.v:
            p = add(v, neg(o))
            p = (p[idx[0]], p[idx[1]], p[idx[2]])
            tv.append((p[0] + sx * p[2], p[1] + sy * p[2], p[2]))  # :137-139
        edge = lambda a, b: a[0] * b[1] - b[0] * a[1]  # :141-143
        e = (edge(tv[1], tv[2]), edge(tv[2], tv[0]), edge(tv[0], tv[1]))  # :145-158
        if not (all(sign_positive(x) for x in e) or all(not sign_positive(x) for x in e)):
            return None
        ae = (abs(e[0]), abs(e[1]), abs(e[2]))
        s = 0.0
        for x in ae:
            s = s + x
        inv = 1.0 / s
        bary = (ae[0] * inv, ae[1] * inv, ae[2] * inv)  # :160-162
        tz = 0.0
        for coord, vert in zip(bary, tv):
            tz = tz + vert[2] * coord
        if sign_positive(tz) != sign_positive(pd[2]):
            return None
        location = (0.0, 0.0, 0.0)
        for coord, vert in zip(bary, self.v):
            location = add(location, scale(vert, coord))
        distance = norm(sub(o, location))
        nrm = (0.0, 0.0, 0.0)
        for coord, vn in zip(bary, self.n):
            nrm = add(nrm, scale(vn, coord))
        normal = normalize(nrm)
        cotangent = normalize(cross(sub(self.v[0], self.v[1]), normal))
        tangent = normalize(cross(cotangent, normal))
        return dict(distance=distance, location=location, normal=normal, tangent=tangent, cotangent=cotangent,
                    retro=normalize(sub(o, location)), material=self.material, what=("triangle", self.prim_id))


def min_by_distance(hits):  # sampler.rs:14-19 / vec_aggregate.rs:13-21: Iterator::min_by keeps the FIRST minimum; NaN compares Less
    best = None
    for h in hits:
        if best is None:
            best = h
            continue
        a, b = best["distance"], h["distance"]
        if a != a or b != b:
            order = -1  # partial_cmp is None -> Ordering::Less: the accumulator stays
        else:
            order = -1 if a < b else (1 if a > b else 0)
        if order > 0:
            best = h
    return best


class PrimitiveList:  # Vec<Box<dyn Primitive>>
    def __init__(self, prims):
        self.prims = prims

    def intersect(self, ray):
        return min_by_distance([h for h in (p.intersect(ray) for p in self.prims) if h is not None])


def interval_union(a, b):  # util/interval.rs:62-73
    if a[0] > a[1]:
        return b
    if b[0] > b[1]:
        return a
    return (min(a[0], b[0]), max(a[1], b[1]))


EMPTY_BOX = ((INF, -INF),) * 3


def box_union(a, b):
    return tuple(interval_union(a[k], b[k]) for k in range(3))


def largest_dimension(box):  # util/axis_aligned_bounding_box.rs:71-93
    acc, acc_size = 0, 0.0
    for k, (lo, hi) in enumerate(box):
        size = -1.0 if lo == hi else hi - lo
        if size > acc_size:
            acc, acc_size = k, size
    return acc


def box_intersect(box, ray):  # raycasting/axis_aligned_bounding_box.rs:9-27: a LINE test (t may be negative)
    lo, hi = -INF, INF
    for k in range(3):
        o, d = ray[0][k], ray[1][k]
        a = fdiv(box[k][0] - o, d)
        b = fdiv(box[k][1] - o, d)
        imin, imax = (b, a) if a > b else (a, b)  # Interval::new
        lo, hi = fmax(lo, imin), fmin(hi, imax)
        if lo > hi:
            return False
    return True


def fdiv(a, b):  # IEEE division including the zero divisor
    if b != 0.0:
        return a / b
    if a != a or a == 0.0:
        return float("nan")
    return math.copysign(INF, a) * math.copysign(1.0, b)


def fmax(a, b):  # f64::max: ignores a NaN operand
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def fmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


class Bvh:  # bounding_volume_hierarchy.rs:18-119
    def __init__(self, tris):
        self.leaf_order = []
        self.root = self._build(list(tris))

    def _build(self, prims):  # :49-75, sort = stable (the reference's sort_unstable leaves ties unspecified; DESIGN deviation iii)
        bounds = EMPTY_BOX
        for p in prims:
            bounds = box_union(bounds, p.box())
        if len(prims) <= 1:
            self.leaf_order.extend(p.prim_id for p in prims)
            return ("leaf", bounds, prims)
        axis = largest_dimension(bounds)
        prims.sort(key=lambda p: (p.box()[axis][0] + p.box()[axis][1]) / 2.0)  # :30-36, :38-46; list.sort is stable
        pivot = len(prims) // 2
        left = self._build(prims[:pivot])
        right = self._build(prims[pivot:])
        return ("node", bounds, left, right)

    @staticmethod
    def _closest(a, b):  # :77-92: a only if strictly closer
        if a is None:
            return b
        if b is None:
            return a
        return a if a["distance"] < b["distance"] else b

    def _intersect(self, node, ray):  # :94-119
        if not box_intersect(node[1], ray):
            return None
        if node[0] == "node":
            return self._closest(self._intersect(node[2], ray), self._intersect(node[3], ray))
        acc = None
        for p in node[2]:
            acc = self._closest(acc, p.intersect(ray))
        return acc

    def intersect(self, ray):
        return self._intersect(self.root, ray)


def sample_scene(objects, ray):  # sampler.rs:9-20
    return min_by_distance([h for h in (o.intersect(ray) for o in objects) if h is not None])


# ------------------------------------------------------------------ integrators/*.rs
def sky(basis, w_o, wavelength):  # simple_random_integrator.rs:57-65
    return reflection_from_linear_rgb(basis, w_o[1], w_o[1], 1.0).intensity_at_wavelength(wavelength)


def simple_random_integrate(basis, objects, info, photon, limit, rng, counters):  # simple_random_integrator.rs:12-55
    if limit == 0:
        counters["limited"] += 1
        return (0.0, 0.0)
    w2b = [info["tangent"], info["cotangent"], info["normal"]]
    b2w = try_inverse(w2b)
    assert b2w is not None
    w_i = mat_mul_vec(w2b, info["retro"])
    w_o, pdf = info["material"].sample(w_i, photon, rng)
    world_w_o = mat_mul_vec(b2w, w_o)
    counters["bounce"] += 1
    hit = sample_scene(objects, ray_bias(ray_new(info["location"], world_w_o), 0.0000001))
    if hit is None:
        counters["escaped"] += 1
        incoming = (photon[0], sky(basis, world_w_o, photon[0]))
    else:
        incoming = simple_random_integrate(basis, objects, hit, photon, limit - 1, rng, counters)
    incoming = (incoming[0], incoming[1] * pdf)
    incoming = (incoming[0], incoming[1] * abs(dot(world_w_o, info["normal"])))
    return info["material"].bsdf(w_o, w_i, incoming)


def whitted_integrate(objects, lights, ambient, info, photon, limit, rng, counters):  # whitted_integrator.rs:20-86
    w2b = [info["tangent"], info["cotangent"], info["normal"]]
    b2w = try_inverse(w2b)
    assert b2w is not None
    terms = []
    for ldir, lspec in lights:
        counters["shadow"] += 1
        if sample_scene(objects, ray_bias(ray_new(info["location"], ldir), 0.0000001)) is not None:
            terms.append((photon[0], ambient.intensity_at_wavelength(photon[0])))
        else:
            emitted = (photon[0], lspec.intensity_at_wavelength(photon[0]) * abs(dot(ldir, info["normal"])))
            terms.append(info["material"].bsdf(mat_mul_vec(w2b, info["retro"]), mat_mul_vec(w2b, ldir), emitted))
    direction, _pdf = info["material"].sample(mat_mul_vec(w2b, info["retro"]), photon, rng)
    world_dir = mat_mul_vec(b2w, direction)
    counters["bounce"] += 1
    hit = sample_scene(objects, ray_bias(ray_new(info["location"], world_dir), 0.0000001))
    if hit is not None and limit > 0:
        inner = whitted_integrate(objects, lights, ambient, hit, photon, limit - 1, rng, counters)
        ph = info["material"].bsdf(mat_mul_vec(w2b, info["retro"]), direction, inner)
        terms.append((ph[0], ph[1] * abs(dot(world_dir, info["normal"]))))
    else:
        terms.append((photon[0], photon[1] * 0.0))
    total = photon[1]
    for t in terms:
        total += t[1]
    return (photon[0], total)


# ------------------------------------------------------------------ camera.rs
def camera_ray(cam, width, height, row, column, rng):  # camera.rs:24-66: two Standard draws, x then y
    w, h = float(width), float(height)
    film_w, film_h = (w / h, 1.0) if w > h else (1.0, w / h)

    def sc(i, n, l):  # :45-50
        pixel_size = l * (1.0 / float(n))
        return (float(i) + rng.f64()) * pixel_size
    x = sc(column, width, film_w) - film_w * 0.5
    y = sc(height - (row + 1), height, film_h) - film_h * 0.5
    return ray_new(cam, (x, y, 1.0))


def render_sample(basis, cam, objects, width, height, row, column, seed, sample, max_depth, integrator="simple", lights=(), ambient=None):
    """camera.rs:105-128 for ONE sample of one pixel, with the draw layout of SURVEY 8a row 27: ordinals 0,1 = camera x,y;
    2 = wavelength; 3 unused; 4... = material draws in call order.  Returns the photon passed to update_pixel."""
    pixel = row * width + column
    counters = dict(bounce=0, escaped=0, limited=0, shadow=0, missed=0)
    ray = camera_ray(cam, width, height, row, column, Rng(seed, pixel, sample, 0))
    hit = sample_scene(objects, ray)
    if hit is None:
        counters["missed"] += 1
        photon = (0.0, 0.0)
    else:
        wl = 380.0 + (740.0 - 380.0) * Rng(seed, pixel, sample, 2).f64()  # photon.rs:18-24
        rng = Rng(seed, pixel, sample, 4)
        if integrator == "simple":
            photon = simple_random_integrate(basis, objects, hit, (wl, 0.0), max_depth, rng, counters)
        else:
            photon = whitted_integrate(objects, lights, ambient, hit, (wl, 0.0), max_depth, rng, counters)
    return (photon[0], photon[1] * (740.0 - 380.0)), counters, hit  # camera.rs:124, photon.rs:26-28


# ------------------------------------------------------------------ mesh.rs + obj 0.9 (simple polygons)
def f32(x):
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def load_obj(text, material):
    """obj 0.9 semantics as mesh.rs uses them: `v`, `vn`, `f` with a, a/b, a//c, a/b/c (1-based, negative = relative to the
    end at that point); positions parsed as f32 then widened (mesh.rs:21-26); missing normal -> zeros (:37); fan triangulation
    around the polygon's first vertex (mesh.rs:42-72)."""
    pos, nrm, tris = [], [], []
    for line in text.splitlines():
        tok = line.split("#")[0].split()
        if not tok:
            continue
        if tok[0] == "v":
            pos.append(tuple(f32(t) for t in tok[1:4]))
        elif tok[0] == "vn":
            nrm.append(tuple(f32(t) for t in tok[1:4]))
        elif tok[0] == "f":
            poly = []
            for t in tok[1:]:
                parts = t.split("/")
                vi = int(parts[0])
                vi = vi - 1 if vi > 0 else len(pos) + vi
                ni = None
                if len(parts) == 3 and parts[2]:
                    ni = int(parts[2])
                    ni = ni - 1 if ni > 0 else len(nrm) + ni
                poly.append((pos[vi], nrm[ni] if ni is not None else (0.0, 0.0, 0.0)))
            for a, b in zip(poly[1:], poly[2:]):
                tris.append(Triangle([poly[0][0], a[0], b[0]], [poly[0][1], a[1], b[1]], material, len(tris)))
    return tris


# ------------------------------------------------------------------ the vectors
def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def vec_bits(v):
    return [bits(x) for x in v]


NAMED = {"Yellow": (1.0, 1.0, 0.0), "Green": (0.0, 0.5, 0.0), "Blue": (0.0, 0.0, 1.0), "Red": (1.0, 0.0, 0.0)}

OBJ_TEXT = """# seven triangles: a quad fan, a negative-index face, ties on the split axis
v 0 0 4
v 1 0 4
v 1 1 4
v 0 1 4
v 0.5 0.5 3.5
vn 0 0 -1
vn 0.1 0 -1
f 1//1 2//1 3//2 4//1
f -5 -4 -1
f 2/1/1 3/1/1 5/1/2
f 3 4 5
v 2 0 4
v 2 1 4
f 2//1 6//1 7//1 3//1
"""


def main():
    basis = read_rgb_basis()
    out = {"about": "second-source vectors: tests/golden/second_source.py (plain Python from the Rust lines); floats as IEEE bit patterns"}
    # --- RNG conversions
    out["rng"] = []
    for seed, pixel, sample, ordinal in ((1, 0, 0, 0), (1, 12345, 7, 4), (0xDEADBEEFCAFE, 2073599, (1 << 33) + 5, 9), (7, 1, 1, 3)):
        a, b, c = Rng(seed, pixel, sample, ordinal), Rng(seed, pixel, sample, ordinal), Rng(seed, pixel, sample, ordinal)
        out["rng"].append(dict(seed=seed, pixel=pixel, sample=sample, ordinal=ordinal, f64=bits(a.f64()), open01=bits(b.open01()), boolean=int(c.boolean())))
    # --- spectra
    out["spectrum"] = []
    for rgb in ((1.0, 1.0, 0.0), (0.0, 0.5, 0.0), (0.55, 0.27, 0.04), (0.3, 0.3, 1.0), (0.9, 0.1, 0.5), (0.2, 0.8, 0.8)):
        sp = reflection_from_linear_rgb(basis, *rgb)
        for w in (379.0, 380.0, 455.5, 600.25, 719.999, 720.0, 730.0):
            out["spectrum"].append(dict(rgb=rgb, wavelength=w, intensity=bits(sp.intensity_at_wavelength(w))))
    out["diamond"] = [dict(wavelength=w, eta=bits(diamond().intensity_at_wavelength(w))) for w in (380.0, 500.5, 633.0, 740.0)]
    out["sky"] = [dict(w=list(w), wavelength=wl, value=bits(sky(basis, w, wl))) for w, wl in
                  (((0.1, 0.7, 0.2), 450.0), ((0.0, -0.3, 1.0), 610.0), ((0.5, 1.4, 0.1), 700.0), ((0.2, 0.999, 0.0), 380.0))]
    # --- materials: sample + bsdf for fixed draws
    colour = reflection_from_linear_rgb(basis, 0.55, 0.27, 0.04)
    mats = {"lambertian": (Lambertian(colour, 0.1), dict(kind=0, rgb=(0.55, 0.27, 0.04), p=(0.1, 0.0, 0.0))),
            "phong": (Phong(colour, 0.3, 0.5, 20.0), dict(kind=1, rgb=(0.55, 0.27, 0.04), p=(0.3, 0.5, 20.0))),
            "reflective": (Reflective(colour, 0.05, 0.9), dict(kind=2, rgb=(0.55, 0.27, 0.04), p=(0.05, 0.9, 0.0))),
            "dielectric": (Dielectric(diamond()), dict(kind=3, rgb=None, p=(0.0, 0.0, 0.0)))}
    out["materials"] = []
    w_is = [normalize((0.3, -0.2, 0.9)), normalize((-0.5, 0.1, 0.4)), normalize((0.2, 0.3, -0.7)), normalize((0.01, 0.0, 1.0)),
            normalize((0.9, 0.1, 0.05)), normalize((0.6, 0.0, -0.1))]
    for name, (m, desc) in mats.items():
        for k, w_i in enumerate(w_is):
            for wl in (420.0, 555.5, 689.0):
                seed, pixel, sample, ordinal = 3, 100 + k, 5, 4 + (k % 2 if name == "dielectric" else 0)
                rng = Rng(seed, pixel, sample, ordinal)
                d, pdf = m.sample(w_i, (wl, 0.0), rng)
                b1 = m.bsdf(d, w_i, (wl, 0.75))   # SimpleRandom's argument order (sampled, retro)
                b2 = m.bsdf(w_i, d, (wl, 0.75))   # Whitted's (retro, incoming)
                out["materials"].append(dict(material=name, desc=desc, w_i=vec_bits(w_i), wavelength=wl, seed=seed, pixel=pixel, sample=sample,
                                             ordinal=ordinal, direction=vec_bits(d), pdf=bits(pdf), ordinal_after=rng.ordinal,
                                             bsdf_sampled_retro=bits(b1[1]), bsdf_retro_sampled=bits(b2[1])))
    out["fresnel"] = []
    for w_i in w_is:
        for eta1, eta2 in ((1.0, 2.42), (2.42, 1.0), (1.0, 1.5)):
            refl, rs, td, ts = fresnel(w_i, eta1, eta2)
            out["fresnel"].append(dict(w_i=vec_bits(w_i), eta1=eta1, eta2=eta2, refl=vec_bits(refl), rs=bits(rs), td=vec_bits(td), ts=bits(ts)))
    # --- OBJ loader + BVH build (leaf order) + traversal with ties
    yellow = Lambertian(reflection_from_linear_rgb(basis, *NAMED["Yellow"]), 0.05)
    tris = load_obj(OBJ_TEXT, yellow)
    out["obj"] = dict(text=OBJ_TEXT, triangles=[dict(v=[vec_bits(p) for p in t.v], n=[vec_bits(p) for p in t.n]) for t in tris])
    bvh = Bvh(tris)
    out["bvh"] = dict(leaf_order=bvh.leaf_order, rays=[])
    rays = [((0.5, 0.5, 0.0), (0.0, 0.0, 1.0)), ((0.25, 0.75, 0.0), (0.01, -0.02, 1.0)), ((1.5, 0.5, 1.0), (0.0, 0.0, 1.0)),
            ((0.9, 0.2, 0.0), (0.02, 0.01, 1.0)), ((3.0, 3.0, 0.0), (0.0, 0.0, 1.0)), ((0.5, 0.5, 8.0), (0.01, 0.02, -1.0)),
            ((0.75, 0.4, 0.0), (0.0, 0.0, 1.0)), ((-1.0, 0.5, 3.0), (1.0, 0.0, 0.3)), ((0.6, 0.6, 0.0), (-0.03, -0.04, 1.0))]
    for o, d in rays:
        h = bvh.intersect(ray_new(o, d))
        out["bvh"]["rays"].append(dict(origin=list(o), direction=list(d), prim=(h["what"][1] if h else -1),
                                       distance=(bits(h["distance"]) if h else None)))
    # --- whole samples: the main.rs scene (plane + three spheres) + the 7-triangle mesh, SimpleRandom and Whitted
    def scene(variant):
        ground = Lambertian(reflection_from_linear_rgb(basis, 0.55, 0.27, 0.04), 0.1)
        if variant == "lambertian":
            m1 = Lambertian(reflection_from_linear_rgb(basis, *NAMED["Green"]), 0.1)
            m2 = Lambertian(reflection_from_linear_rgb(basis, *NAMED["Blue"]), 0.1)
            m3 = Lambertian(reflection_from_linear_rgb(basis, *NAMED["Red"]), 0.05)
            mesh_mat = yellow
        else:
            m1 = Phong(reflection_from_linear_rgb(basis, *NAMED["Green"]), 0.3, 0.5, 20.0)
            m2 = Reflective(reflection_from_linear_rgb(basis, *NAMED["Blue"]), 0.01, 0.99)
            m3 = Dielectric(diamond())
            mesh_mat = Reflective(reflection_from_linear_rgb(basis, *NAMED["Yellow"]), 0.05, 0.9)
        prims = [Plane((0.0, 1.0, 0.0), -2.0, ground), Sphere((-6.25, -0.5, 1.0), 1.0, m1), Sphere((-4.25, -0.5, 2.0), 1.0, m2),
                 Sphere((-5.0, 1.5, 1.0), 1.0, m3)]
        # faces the OBJ gives no normals (mesh.rs:37: zeros) would shade with NaN frames; the render vectors use a usable normal there
        mesh = [Triangle([add(scale(p, 2.0), (-3.0, -1.5, -3.0)) for p in t.v], [n if n != (0.0, 0.0, 0.0) else (0.0, 0.0, -1.0) for n in t.n],
                         mesh_mat, t.prim_id) for t in tris]
        return [PrimitiveList(prims), Bvh(mesh)], mesh
    cam = (-2.0, 1.0, -5.0)
    out["samples"] = []
    W, H = 24, 16
    for variant, integrator, depth in (("lambertian", "simple", 8), ("mixed", "simple", 6), ("lambertian", "whitted", 2), ("mixed", "whitted", 1),
                                       ("lambertian", "simple", 0), ("lambertian", "whitted", 0)):
        objects, mesh = scene(variant)
        lights = [((1.0, 1.0, -1.0), grey(1.0))] if integrator == "whitted" else []
        ambient = grey(0.05)
        rows = []
        for row in range(H):
            for column in range(W):
                for sample in (0, 3):
                    ph, counters, hit = render_sample(basis, cam, objects, W, H, row, column, 11, sample, depth, integrator, lights, ambient)
                    rows.append([row, column, sample, bits(ph[0]), bits(ph[1]), counters["bounce"], counters["shadow"]])
        out["samples"].append(dict(variant=variant, integrator=integrator, max_depth=depth, width=W, height=H, seed=11, camera=cam,
                                   mesh=[dict(v=[list(p) for p in t.v], n=[list(p) for p in t.n]) for t in mesh], photons=rows))
    path = os.path.join(HERE, "second_source.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
