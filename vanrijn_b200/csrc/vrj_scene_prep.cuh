// vrj_scene_prep.cuh -- scene upload, device side: the caller's flattened SoA arrays (VrjSceneDesc) are copied
// to the device as they are and re-expressed there as the traversal layout of vrj_traverse.cuh:
//   k_pack_triangles : 6 vertex/normal arrays + material + prim id  ->  96-byte position and normal records,
//                      optionally through a permutation (triangles of a BVH built on the device move to leaf order)
//   k_mark_internal / exclusive scan / k_wide_nodes : reference-topology nodes (a box per node)  ->  "wide" nodes that
//                      carry the boxes of both children (f32 rounded outward, and f64), leaves folded into child refs
// Nothing here decides a hit: boxes only feed the conservative filter (vrj_traverse.cuh).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace vrj {

struct RawTriangles {
    const double *v0, *v1, *v2, *n0, *n1, *n2; // 4 doubles per entry
    const uint32_t *material, *prim_id;
};

__global__ void k_pack_triangles(uint32_t n, RawTriangles raw, const uint32_t *__restrict__ perm, double *__restrict__ tri_pos,
                                 double *__restrict__ tri_nrm, float *__restrict__ tri_pos32, float *__restrict__ tri_nrm32) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t src = perm ? perm[t] : t;
    const double *v[3] = {raw.v0 + 4 * src, raw.v1 + 4 * src, raw.v2 + 4 * src};
    const double *nr[3] = {raw.n0 + 4 * src, raw.n1 + 4 * src, raw.n2 + 4 * src};
    double *tp = tri_pos + (size_t)t * 12, *tn = tri_nrm + (size_t)t * 12;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double2 a = *reinterpret_cast<const double2 *>(v[k]), b = *reinterpret_cast<const double2 *>(nr[k]);
        tp[3 * k] = a.x, tp[3 * k + 1] = a.y, tp[3 * k + 2] = v[k][2];
        tn[3 * k] = b.x, tn[3 * k + 1] = b.y, tn[3 * k + 2] = nr[k][2];
    }
    const unsigned long long bits = ((unsigned long long)raw.prim_id[src] << 32) | raw.material[src];
    tp[9] = __longlong_as_double((long long)bits);
    tp[10] = tp[11] = 0.0, tn[9] = tn[10] = tn[11] = 0.0;
    // the binary32 copies of VRJ_PRECISION_F32_FAST: 48-byte records (round to nearest; OBJ vertices are f32 already)
    float *fp = tri_pos32 + (size_t)t * 12, *fn = tri_nrm32 + (size_t)t * 12;
#pragma unroll
    for (int k = 0; k < 9; k++) fp[k] = (float)tp[k], fn[k] = (float)tn[k];
    fp[9] = __uint_as_float(raw.material[src]), fp[10] = __uint_as_float(raw.prim_id[src]), fp[11] = 0.f;
    fn[9] = fn[10] = fn[11] = 0.f;
}

// input-order triangle index of every leaf position of a BVH built on the device
__global__ void k_offset_perm(uint32_t n, const uint32_t *__restrict__ order, uint32_t first, uint32_t *__restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[first + i] = first + order[i];
}
__global__ void k_identity_perm(uint32_t n, uint32_t *__restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}
// 3 x (4 doubles per vertex) -> 9 doubles per triangle, the builder's input
__global__ void k_gather_vertices(uint32_t n, const double *__restrict__ v0, const double *__restrict__ v1, const double *__restrict__ v2,
                                  double *__restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double *v[3] = {v0 + 4 * (size_t)t, v1 + 4 * (size_t)t, v2 + 4 * (size_t)t};
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) out[(size_t)t * 9 + 3 * k + c] = v[k][c];
}

// One BVH's nodes inside the device node arrays.  Built by the caller: children and leaf triangles are absolute indices
// (VrjSceneDesc); built on the device: they are local to the BVH and get the offsets below.
struct BvhNodes {
    const double *node_min, *node_max; // 4 doubles per node, indexed by absolute node index
    const int32_t *node_child;         // 2 per node
    uint32_t first_node, n_nodes;
    int32_t child_offset, triangle_offset;
    uint32_t wide_base;                // index of this BVH's first wide node
};

__global__ void k_mark_internal(BvhNodes b, uint32_t *__restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n_nodes) flags[i] = b.node_child[2 * (size_t)(b.first_node + i)] >= 0 ? 1u : 0u;
}

__device__ __forceinline__ void put_child(const BvhNodes &b, const uint32_t *__restrict__ rank, float *__restrict__ f, double *__restrict__ g,
                                          int which, int64_t node /* absolute, or -1 for "no child" */, int32_t &ref_out) {
    double lo[3], hi[3];
    int32_t ref = -1;
    bool empty = node < 0;
    if (!empty) {
        const int32_t l = b.node_child[2 * node], r = b.node_child[2 * node + 1];
        if (l >= 0) {
            ref = (int32_t)(b.wide_base + rank[node - b.first_node]);
        } else if (r == 0) {
            empty = true; // an empty leaf never reports a hit
        } else {
            ref = ~((~l) + b.triangle_offset);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = empty ? CUDART_INF : b.node_min[4 * node + k];
        hi[k] = empty ? -CUDART_INF : b.node_max[4 * node + k];
    }
    // f32 layout: [c0.lox c0.hix c0.loy c0.hiy][c1.lox c1.hix c1.loy c1.hiy][c0.loz c0.hiz c1.loz c1.hiz][l r 0 0]
    f[which * 4 + 0] = __double2float_rd(lo[0]), f[which * 4 + 1] = __double2float_ru(hi[0]);
    f[which * 4 + 2] = __double2float_rd(lo[1]), f[which * 4 + 3] = __double2float_ru(hi[1]);
    f[8 + which * 2 + 0] = __double2float_rd(lo[2]), f[8 + which * 2 + 1] = __double2float_ru(hi[2]);
    // f64 layout: c0 {lox hix loy hiy loz hiz} c1 {...} {bits(l, r), 0}
#pragma unroll
    for (int k = 0; k < 3; k++) g[which * 6 + 2 * k] = lo[k], g[which * 6 + 2 * k + 1] = hi[k];
    ref_out = ref;
}

// one thread per node of the BVH; internal nodes write their wide node.  `rank` = exclusive scan of k_mark_internal.
__global__ void k_wide_nodes(BvhNodes b, const uint32_t *__restrict__ rank, float *__restrict__ nodes32, double *__restrict__ nodes64) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_nodes) return;
    const int64_t me = (int64_t)b.first_node + i;
    const int32_t l = b.node_child[2 * me], r = b.node_child[2 * me + 1];
    int64_t c0, c1;
    uint32_t w;
    if (l >= 0) {
        w = b.wide_base + rank[i];
        c0 = (int64_t)l + b.child_offset, c1 = (int64_t)r + b.child_offset;
    } else if (i == 0) {
        w = b.wide_base; // the root is a leaf: one wide node whose first child is that leaf
        c0 = me, c1 = -1;
    } else {
        return;
    }
    float *f = nodes32 + (size_t)w * 16;
    double *g = nodes64 + (size_t)w * 14;
    int32_t r0, r1;
    put_child(b, rank, f, g, 0, c0, r0);
    put_child(b, rank, f, g, 1, c1, r1);
    f[12] = __int_as_float(r0), f[13] = __int_as_float(r1), f[14] = 0.f, f[15] = 0.f;
    g[12] = __longlong_as_double((long long)(((unsigned long long)(uint32_t)r1 << 32) | (uint32_t)r0));
    g[13] = 0.0;
}

// ---- 16-bit nodes (VRJ_FILTER_Q16): the 2-wide tree with both children's boxes quantised OUTWARD onto one 65536^3 grid over
// the scene's meshes -- 32-byte records, one 256-bit load per step instead of two: {lo.x, hi.x, lo.y, hi.y} {lo.z, hi.z, l, r},
// each coordinate word = child 0 in the low half, child 1 in the high half.  A plane sits at grid.lo + q * grid.cell.
struct QGrid {
    double lo[3], cell[3];
};
__device__ __forceinline__ uint32_t quantise_lo(double v, double glo, double cell) {
    double x = floor((v - glo) / cell);
    x = x < 0.0 ? 0.0 : (x > 65535.0 ? 65535.0 : x);
    uint32_t q = (uint32_t)x;
    if (q > 0 && glo + (double)q * cell > v) q--; // the division rounded up across a grid line
    return q;
}
__device__ __forceinline__ uint32_t quantise_hi(double v, double glo, double cell) {
    double x = ceil((v - glo) / cell);
    x = x < 0.0 ? 0.0 : (x > 65535.0 ? 65535.0 : x);
    uint32_t q = (uint32_t)x;
    if (q < 65535u && glo + (double)q * cell < v) q++;
    return q;
}
__device__ __forceinline__ void q16_child(const BvhNodes &b, const uint32_t *__restrict__ rank, const QGrid &g, int64_t node, uint32_t qlo[3],
                                          uint32_t qhi[3], int32_t &ref) {
    ref = -1;
    bool empty = node < 0;
    if (!empty) {
        const int32_t l = b.node_child[2 * node], r = b.node_child[2 * node + 1];
        if (l >= 0) ref = (int32_t)(b.wide_base + rank[node - b.first_node]);
        else if (r == 0) empty = true;
        else ref = ~((~l) + b.triangle_offset);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        qlo[k] = empty ? 65535u : quantise_lo(b.node_min[4 * node + k], g.lo[k], g.cell[k]);
        qhi[k] = empty ? 0u : quantise_hi(b.node_max[4 * node + k], g.lo[k], g.cell[k]);
    }
}
__global__ void k_q16_nodes(BvhNodes b, const uint32_t *__restrict__ rank, QGrid g, uint4 *__restrict__ nodesq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_nodes) return;
    const int64_t me = (int64_t)b.first_node + i;
    const int32_t l = b.node_child[2 * me], r = b.node_child[2 * me + 1];
    int64_t c0, c1;
    uint32_t w;
    if (l >= 0) {
        w = b.wide_base + rank[i];
        c0 = (int64_t)l + b.child_offset, c1 = (int64_t)r + b.child_offset;
    } else if (i == 0) {
        w = b.wide_base;
        c0 = me, c1 = -1;
    } else {
        return;
    }
    uint32_t lo0[3], hi0[3], lo1[3], hi1[3];
    int32_t r0, r1;
    q16_child(b, rank, g, c0, lo0, hi0, r0);
    q16_child(b, rank, g, c1, lo1, hi1, r1);
    nodesq[(size_t)w * 2] = make_uint4(lo0[0] | (lo1[0] << 16), hi0[0] | (hi1[0] << 16), lo0[1] | (lo1[1] << 16), hi0[1] | (hi1[1] << 16));
    nodesq[(size_t)w * 2 + 1] = make_uint4(lo0[2] | (lo1[2] << 16), hi0[2] | (hi1[2] << 16), (uint32_t)r0, (uint32_t)r1);
}

// ---- 4-wide nodes (VRJ_FILTER_F32X4): every second level of the reference tree is folded away, so a node carries
// the f32 boxes (rounded outward) of its up to four grandchildren and a ray makes half as many dependent fetches.
// 128-byte records: 24 floats [child][lo.x hi.x lo.y hi.y lo.z hi.z], 4 child refs (>= 0: 4-wide node, < 0: ~triangle,
// empty slots have an inverted box and are never entered), 16 bytes of padding.  Leaves stay <= 1 triangle; the
// leaf order (= tie-breaking order) is the reference's.
constexpr uint32_t kNoParent = 0xffffffffu;
__global__ void k_parents(BvhNodes b, uint32_t *__restrict__ parent) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_nodes) return;
    if (i == 0) parent[0] = kNoParent; // the root is the BVH's first node
    const int64_t me = (int64_t)b.first_node + i;
    const int32_t l = b.node_child[2 * me], r = b.node_child[2 * me + 1];
    if (l >= 0) {
        parent[(int64_t)l + b.child_offset - b.first_node] = i;
        parent[(int64_t)r + b.child_offset - b.first_node] = i;
    }
}
// flags[i] = 1 for internal nodes at even depth: they become 4-wide nodes
__global__ void k_quad_flags(BvhNodes b, const uint32_t *__restrict__ parent, uint32_t *__restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_nodes) return;
    uint32_t depth = 0;
    for (uint32_t p = parent[i]; p != kNoParent && depth < 64; p = parent[p]) depth++;
    const bool internal = b.node_child[2 * (size_t)(b.first_node + i)] >= 0;
    flags[i] = (internal && (depth & 1u) == 0u) ? 1u : 0u;
}
__device__ __forceinline__ void put_quad_slot(const BvhNodes &b, const uint32_t *__restrict__ rank4, float *__restrict__ f, int slot,
                                              int64_t node /* absolute, or -1 */) {
    double lo[3], hi[3];
    int32_t ref = -1;
    bool empty = node < 0;
    if (!empty) {
        const int32_t l = b.node_child[2 * node], r = b.node_child[2 * node + 1];
        if (l >= 0) ref = (int32_t)(b.wide_base + rank4[node - b.first_node]);
        else if (r == 0) empty = true;
        else ref = ~((~l) + b.triangle_offset);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = empty ? CUDART_INF : b.node_min[4 * node + k];
        hi[k] = empty ? -CUDART_INF : b.node_max[4 * node + k];
        f[slot * 6 + 2 * k] = __double2float_rd(lo[k]), f[slot * 6 + 2 * k + 1] = __double2float_ru(hi[k]);
    }
    f[24 + slot] = __int_as_float(ref);
}
// `b.wide_base` is the index of this BVH's first 4-wide node; rank4 = exclusive scan of k_quad_flags
__global__ void k_quad_nodes(BvhNodes b, const uint32_t *__restrict__ flags_scanned, const uint32_t *__restrict__ is_quad,
                             float *__restrict__ nodes4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_nodes) return;
    const int64_t me = (int64_t)b.first_node + i;
    const int32_t l = b.node_child[2 * me], r = b.node_child[2 * me + 1];
    int64_t slots[4] = {-1, -1, -1, -1};
    uint32_t q;
    if (l >= 0) {
        if (!is_quad[i]) return;
        q = b.wide_base + flags_scanned[i];
        const int64_t c[2] = {(int64_t)l + b.child_offset, (int64_t)r + b.child_offset};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int32_t cl = b.node_child[2 * c[k]], cr = b.node_child[2 * c[k] + 1];
            if (cl >= 0) slots[2 * k] = (int64_t)cl + b.child_offset, slots[2 * k + 1] = (int64_t)cr + b.child_offset;
            else slots[2 * k] = c[k]; // a leaf child keeps its own slot
        }
    } else if (i == 0) {
        q = b.wide_base; // the root is a leaf
        slots[0] = me;
    } else {
        return;
    }
    float *f = nodes4 + (size_t)q * 32;
#pragma unroll
    for (int s = 0; s < 4; s++) put_quad_slot(b, flags_scanned, f, s, slots[s]);
    f[28] = f[29] = f[30] = f[31] = 0.f;
}

} // namespace vrj
