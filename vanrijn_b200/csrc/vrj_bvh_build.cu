// vrj_bvh_build.cu -- BoundingVolumeHierarchy::build (bounding_volume_hierarchy.rs:38-75) on the device,
// producing EXACTLY the tree the host builder (csrc/host/vanrijn_host.cpp, BuildCtx::build) produces:
// same node boxes, same DFS pre-order node numbering, same leaf order -- so hit ids and tie-breaking
// are unchanged and the tree is interchangeable with the host-built one (tests compare them bit for bit).
//
// The reference recursion is: bounds = union of the primitives' boxes; axis = largest_dimension(bounds);
// stable-sort the slice by box centre on that axis; split at len/2; recurse.  The SHAPE of that tree
// depends only on n (every segment [begin, begin+len) and its node index are known up front: left = me+1,
// right = me + 2*(len/2)); only the CONTENT of the segments is data dependent.  So the build runs level by
// level over all segments of a level at once:
//   * levels whose segments are longer than SMALL:  one segmented reduction (boxes), then a segmented,
//     stable LSD radix sort of (orderable 64-bit centre key, triangle) pairs -- every CTA owns a chunk
//     that lies inside one segment, histograms are laid out [segment][digit][chunk] so ONE global
//     exclusive scan yields the scatter offsets of all segments; digits in which no key of the level
//     differs are skipped (centres of f32-parsed meshes have 27 constant low bits);
//   * the first level whose segments fit SMALL: one CTA per segment finishes the whole subtree in shared
//     memory (box reduction, axis, rank sort by (key, position) = stable, split, repeat).
// All arithmetic that reaches the output is min / max / one add / one halving of binary64 values, in
// the host builder's order, so the results are identical, not merely close.  (-0.0 and +0.0 compare
// equal in the sort, as with the host's operator<; a box bound that is a zero may differ in sign.)
#include "vrj_internal.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace vrj_build {

constexpr int SMALL = 1024;     // segments up to this long are finished inside one CTA (rank sort: work grows with SMALL)
constexpr int CHUNK = 2048;     // elements per CTA in the global passes
constexpr int CHUNK_THREADS = 256;
constexpr int SCAN_TILE = 2048; // histogram entries per CTA in the scan
constexpr unsigned long long SIGN = 0x8000000000000000ull;

struct Segment {
    uint32_t begin, len, me; // positions [begin, begin+len) of the order array; node index (DFS pre-order)
};
struct Chunk {
    uint32_t begin, count; // positions [begin, begin+count), inside one segment
    uint32_t seg;          // index into the level's segment table
    uint32_t hist_base;    // 256 * (index of the segment's first chunk)
    uint32_t nb, bi;       // chunks in the segment, index of this chunk among them
};

// monotone map binary64 -> uint64 (x + 0.0 folds -0.0 onto +0.0 so the two sort as equal, like operator<)
__device__ __forceinline__ unsigned long long sort_key(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x + 0.0);
    return (u & SIGN) ? ~u : (u | SIGN);
}
// the same map without the zero fold, for min / max of box bounds
__device__ __forceinline__ unsigned long long ordered(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & SIGN) ? ~u : (u | SIGN);
}
__device__ __forceinline__ double unordered(unsigned long long k) {
    return __longlong_as_double((long long)((k & SIGN) ? (k ^ SIGN) : ~k));
}
constexpr unsigned long long ORD_POS_INF = 0xfff0000000000000ull; // ordered(+inf)
constexpr unsigned long long ORD_NEG_INF = 0x000fffffffffffffull; // ordered(-inf)

// util/axis_aligned_bounding_box.rs:76-99: first strictly-largest extent; a degenerate extent counts as -1
__device__ __forceinline__ int largest_dimension(const double lo[3], const double hi[3]) {
    int dim = 0;
    double best = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double extent = (lo[k] == hi[k]) ? -1.0 : hi[k] - lo[k];
        if (extent > best) dim = k, best = extent;
    }
    return dim;
}

// ---- per-triangle boxes and centres (triangle.rs bounding box of the three vertices; bvh.rs:30-36) ----
__global__ void k_tri_boxes(uint32_t n, const double *__restrict__ v, double *__restrict__ lo, double *__restrict__ hi,
                            double *__restrict__ centre, uint32_t *__restrict__ order) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double *p = v + (size_t)t * 9;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double a = p[k], b = p[3 + k], c = p[6 + k];
        double mn = fmin(fmin(a, b), c), mx = fmax(fmax(a, b), c);
        lo[3 * (size_t)t + k] = mn, hi[3 * (size_t)t + k] = mx;
        centre[3 * (size_t)t + k] = (mn + mx) / 2.0;
    }
    order[t] = t;
}

// ---- global levels --------------------------------------------------------------------------------
__global__ void k_init_bounds(uint32_t n_seg, unsigned long long *__restrict__ bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg * 6) return;
    bounds[i] = (i % 6) < 3 ? ORD_POS_INF : ORD_NEG_INF; // lo.xyz, hi.xyz
}

__device__ __forceinline__ unsigned long long warp_min(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// union of the triangle boxes of every segment of the level: one CTA per chunk, one atomic per warp and bound
__global__ void __launch_bounds__(CHUNK_THREADS) k_seg_bounds(const Chunk *__restrict__ chunks, const uint32_t *__restrict__ order,
                                                               const double *__restrict__ lo, const double *__restrict__ hi,
                                                               unsigned long long *__restrict__ bounds) {
    const Chunk c = chunks[blockIdx.x];
    unsigned long long mn[3] = {ORD_POS_INF, ORD_POS_INF, ORD_POS_INF}, mx[3] = {ORD_NEG_INF, ORD_NEG_INF, ORD_NEG_INF};
    for (uint32_t i = threadIdx.x; i < c.count; i += CHUNK_THREADS) {
        uint32_t t = order[c.begin + i];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            unsigned long long a = ordered(lo[3 * (size_t)t + k]), b = ordered(hi[3 * (size_t)t + k]);
            mn[k] = a < mn[k] ? a : mn[k], mx[k] = b > mx[k] ? b : mx[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        unsigned long long a = warp_min(mn[k]), b = warp_max(mx[k]);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bounds[(size_t)c.seg * 6 + k], a);
            atomicMax(&bounds[(size_t)c.seg * 6 + 3 + k], b);
        }
    }
}

// node record + split axis of every segment of a global level (all of them have len >= 2)
__global__ void k_seg_nodes(uint32_t n_seg, const Segment *__restrict__ segs, const unsigned long long *__restrict__ bounds,
                            double *__restrict__ node_min, double *__restrict__ node_max, int32_t *__restrict__ node_child,
                            uint32_t *__restrict__ seg_axis) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const Segment sg = segs[s];
    double lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) lo[k] = unordered(bounds[(size_t)s * 6 + k]), hi[k] = unordered(bounds[(size_t)s * 6 + 3 + k]);
    double *mn = node_min + (size_t)sg.me * 4, *mx = node_max + (size_t)sg.me * 4;
    mn[0] = lo[0], mn[1] = lo[1], mn[2] = lo[2], mn[3] = 0.0;
    mx[0] = hi[0], mx[1] = hi[1], mx[2] = hi[2], mx[3] = 0.0;
    node_child[2 * (size_t)sg.me] = (int32_t)(sg.me + 1);
    node_child[2 * (size_t)sg.me + 1] = (int32_t)(sg.me + 2 * (sg.len / 2));
    seg_axis[s] = (uint32_t)largest_dimension(lo, hi);
}

// sort keys of a level + which key bits vary.  The map to unsigned keys inverts negative values, so their constant
// low bits (an f32-parsed mesh has 27 of them) read as ones where a positive value has zeros; a digit that is constant
// within each sign class is still redundant in an LSD sort (the top digit separates the classes afterwards), so the
// OR and AND of the keys are kept per sign class: varying[0..1] = OR / AND of keys >= 0, varying[2..3] of keys < 0.
__global__ void __launch_bounds__(CHUNK_THREADS) k_make_keys(const Chunk *__restrict__ chunks, const uint32_t *__restrict__ seg_axis,
                                                              const uint32_t *__restrict__ order, const double *__restrict__ centre,
                                                              unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals,
                                                              unsigned long long *__restrict__ varying) {
    const Chunk c = chunks[blockIdx.x];
    const uint32_t axis = seg_axis[c.seg];
    unsigned long long acc[4] = {0ull, ~0ull, 0ull, ~0ull};
    for (uint32_t i = threadIdx.x; i < c.count; i += CHUNK_THREADS) {
        uint32_t t = order[c.begin + i];
        unsigned long long k = sort_key(centre[3 * (size_t)t + axis]);
        keys[c.begin + i] = k, vals[c.begin + i] = t;
        const int cls = (k & SIGN) ? 0 : 2;
        acc[cls] |= k, acc[cls + 1] &= k;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xffffffffu, acc[j], o);
            acc[j] = (j & 1) ? (acc[j] & w) : (acc[j] | w);
        }
    }
    if ((threadIdx.x & 31) == 0) {
        if (acc[0]) atomicOr(varying + 0, acc[0]);
        if (~acc[1]) atomicAnd(varying + 1, acc[1]);
        if (acc[2]) atomicOr(varying + 2, acc[2]);
        if (~acc[3]) atomicAnd(varying + 3, acc[3]);
    }
}

__global__ void __launch_bounds__(CHUNK_THREADS) k_hist(const Chunk *__restrict__ chunks, const unsigned long long *__restrict__ keys,
                                                         int shift, uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[256];
    const Chunk c = chunks[blockIdx.x];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < c.count; i += CHUNK_THREADS) atomicAdd(&h[(keys[c.begin + i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)c.hist_base + (size_t)threadIdx.x * c.nb + c.bi] = h[threadIdx.x];
}

// exclusive scan of the histogram array, three kernels: tiles, tile totals, (the add happens in k_scatter)
__global__ void __launch_bounds__(256) k_scan_tiles(uint32_t n, uint32_t *__restrict__ data, uint32_t *__restrict__ totals) {
    __shared__ uint32_t warp_sums[8];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = base + k < n ? data[base + k] : 0u, sum += v[k];
    uint32_t incl = sum;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; w++) before += warp_sums[w];
    uint32_t run = before + incl - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 255) totals[blockIdx.x] = before + incl;
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t n, uint32_t *__restrict__ totals) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? totals[i] : 0u, incl = v;
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t before = carry;
        for (uint32_t w = 0; w < warp; w++) before += warp_sums[w];
        if (i < n) totals[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
}

__global__ void k_scan_add(uint32_t n, uint32_t *__restrict__ data, const uint32_t *__restrict__ tile_prefix) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += tile_prefix[i / SCAN_TILE];
}

// stable scatter of one digit: warp w of the CTA owns elements [w*256, w*256+256) of the chunk, eight rounds
// of 32 consecutive elements; __match_any_sync ranks equal digits inside a round
__global__ void __launch_bounds__(CHUNK_THREADS) k_scatter(const Chunk *__restrict__ chunks, const unsigned long long *__restrict__ keys_in,
                                                            const uint32_t *__restrict__ vals_in, unsigned long long *__restrict__ keys_out,
                                                            uint32_t *__restrict__ vals_out, int shift, const uint32_t *__restrict__ hist,
                                                            const uint32_t *__restrict__ tile_prefix) {
    __shared__ uint32_t counters[CHUNK_THREADS / 32][256];
    const Chunk c = chunks[blockIdx.x];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < CHUNK_THREADS / 32; w++) counters[w][threadIdx.x] = 0;
    __syncthreads();
    unsigned long long key[8];
    uint32_t val[8], local[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t e = warp * 256 + r * 32 + lane;
        const bool valid = e < c.count;
        key[r] = valid ? keys_in[c.begin + e] : 0ull;
        val[r] = valid ? vals_in[c.begin + e] : 0u;
        const uint32_t d = valid ? (uint32_t)((key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pre = 0;
        if (valid) pre = counters[warp][d];
        __syncwarp();
        if (valid && rank == 0) counters[warp][d] = pre + __popc(peers);
        __syncwarp();
        local[r] = pre + rank;
    }
    __syncthreads();
    {
        const size_t hi = (size_t)c.hist_base + (size_t)threadIdx.x * c.nb + c.bi;
        uint32_t run = hist[hi] + tile_prefix[hi / SCAN_TILE];
        for (int w = 0; w < CHUNK_THREADS / 32; w++) {
            uint32_t t = counters[w][threadIdx.x];
            counters[w][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t e = warp * 256 + r * 32 + lane;
        if (e < c.count) {
            const uint32_t d = (uint32_t)((key[r] >> shift) & 255u);
            const uint32_t pos = counters[warp][d] + local[r];
            keys_out[pos] = key[r], vals_out[pos] = val[r];
        }
    }
}

// ---- small subtrees: one CTA per segment of the first level whose segments fit SMALL -----------------
struct SmallSeg {
    uint16_t begin, len; // CTA-local positions
    uint32_t me;         // node index; len == 0 marks a dead entry
};

__global__ void __launch_bounds__(SMALL, 2048 / SMALL) k_small_subtrees(const Segment *__restrict__ segs, uint32_t *__restrict__ order,
                                                           const double *__restrict__ lo, const double *__restrict__ hi,
                                                           const double *__restrict__ centre, double *__restrict__ node_min,
                                                           double *__restrict__ node_max, int32_t *__restrict__ node_child) {
    extern __shared__ unsigned long long smem_raw[];
    // layout (bytes): bounds 6*8*SMALL/2... sized for the widest level that still splits (<= SMALL/2 segments... SMALL to be safe)
    unsigned long long *bounds = smem_raw;                              // [SMALL][6]
    double *key = reinterpret_cast<double *>(bounds + (size_t)SMALL * 6); // [SMALL]
    uint32_t *idx = reinterpret_cast<uint32_t *>(key + SMALL);          // [SMALL]
    uint32_t *idx2 = idx + SMALL;                                       // [SMALL]
    SmallSeg *cur = reinterpret_cast<SmallSeg *>(idx2 + SMALL);         // [2*SMALL]
    SmallSeg *nxt = cur + 2 * SMALL;                                    // [2*SMALL]
    uint16_t *segof = reinterpret_cast<uint16_t *>(nxt + 2 * SMALL);    // [SMALL]
    uint8_t *axis = reinterpret_cast<uint8_t *>(segof + SMALL);         // [2*SMALL]
    __shared__ uint32_t s_live;

    const Segment root = segs[blockIdx.x];
    const uint32_t p = threadIdx.x;
    const bool mine = p < root.len;
    if (mine) idx[p] = order[root.begin + p], segof[p] = 0;
    if (p == 0) cur[0] = SmallSeg{0, (uint16_t)root.len, root.me};
    uint32_t n_cur = 1;
    __syncthreads();
    while (true) {
        // 1. boxes of the live segments
        for (uint32_t i = p; i < n_cur * 6; i += SMALL) bounds[i] = (i % 6) < 3 ? ORD_POS_INF : ORD_NEG_INF;
        if (p == 0) s_live = 0;
        __syncthreads();
        {
            const uint32_t s = mine ? segof[p] : 0xffffu;
            const uint32_t s0 = __shfl_sync(0xffffffffu, s, 0);
            const bool uniform = __all_sync(0xffffffffu, s == s0) && s0 != 0xffffu;
            unsigned long long mn[3], mx[3];
            if (mine) {
                const uint32_t t = idx[p];
#pragma unroll
                for (int k = 0; k < 3; k++) mn[k] = ordered(lo[3 * (size_t)t + k]), mx[k] = ordered(hi[3 * (size_t)t + k]);
            } else {
#pragma unroll
                for (int k = 0; k < 3; k++) mn[k] = ORD_POS_INF, mx[k] = ORD_NEG_INF;
            }
            if (uniform) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    unsigned long long a = warp_min(mn[k]), b = warp_max(mx[k]);
                    if ((p & 31) == 0) atomicMin(&bounds[s0 * 6 + k], a), atomicMax(&bounds[s0 * 6 + 3 + k], b);
                }
            } else if (mine) {
#pragma unroll
                for (int k = 0; k < 3; k++) atomicMin(&bounds[s * 6 + k], mn[k]), atomicMax(&bounds[s * 6 + 3 + k], mx[k]);
            }
        }
        __syncthreads();
        // 2. node records; children of the segments that split
        for (uint32_t s = p; s < n_cur; s += SMALL) {
            const SmallSeg sg = cur[s];
            SmallSeg l{0, 0, 0}, r{0, 0, 0};
            if (sg.len) {
                double blo[3], bhi[3];
#pragma unroll
                for (int k = 0; k < 3; k++) blo[k] = unordered(bounds[s * 6 + k]), bhi[k] = unordered(bounds[s * 6 + 3 + k]);
                double *mn = node_min + (size_t)sg.me * 4, *mx = node_max + (size_t)sg.me * 4;
                mn[0] = blo[0], mn[1] = blo[1], mn[2] = blo[2], mn[3] = 0.0;
                mx[0] = bhi[0], mx[1] = bhi[1], mx[2] = bhi[2], mx[3] = 0.0;
                if (sg.len <= 1) {
                    node_child[2 * (size_t)sg.me] = ~(int32_t)(root.begin + sg.begin); // leaf: ~first triangle (leaf order)
                    node_child[2 * (size_t)sg.me + 1] = (int32_t)sg.len;
                } else {
                    const uint32_t half = sg.len / 2;
                    node_child[2 * (size_t)sg.me] = (int32_t)(sg.me + 1);
                    node_child[2 * (size_t)sg.me + 1] = (int32_t)(sg.me + 2 * half);
                    axis[s] = (uint8_t)largest_dimension(blo, bhi);
                    l = SmallSeg{sg.begin, (uint16_t)half, sg.me + 1};
                    r = SmallSeg{(uint16_t)(sg.begin + half), (uint16_t)(sg.len - half), sg.me + 2 * half};
                    atomicAdd(&s_live, 1u);
                }
            }
            nxt[2 * s] = l, nxt[2 * s + 1] = r;
        }
        __syncthreads();
        if (s_live == 0) break;
        // 3. keys on each segment's axis; 4. stable rank sort inside the segment
        const SmallSeg sg = mine ? cur[segof[p]] : SmallSeg{0, 0, 0};
        const bool sorting = mine && sg.len >= 2;
        if (sorting) key[p] = centre[3 * (size_t)idx[p] + axis[segof[p]]];
        __syncthreads();
        if (sorting) {
            const double kp = key[p];
            uint32_t rank = 0;
            const uint32_t b = sg.begin, e = sg.begin + sg.len;
            for (uint32_t q = b; q < e; q++) {
                const double kq = key[q];
                rank += (kq < kp || (kq == kp && q < p)) ? 1u : 0u;
            }
            idx2[b + rank] = idx[p];
        } else if (mine) {
            idx2[p] = idx[p];
        }
        __syncthreads();
        if (mine) {
            idx[p] = idx2[p];
            if (sorting) segof[p] = (uint16_t)(2 * segof[p] + ((p - sg.begin) >= (uint32_t)(sg.len / 2) ? 1 : 0));
            else segof[p] = (uint16_t)(2 * segof[p]); // a finished leaf: its (dead) left child keeps it out of every live segment
        }
        SmallSeg *t = cur;
        cur = nxt, nxt = t;
        n_cur *= 2;
        __syncthreads();
    }
    if (mine) order[root.begin + p] = idx[p];
}

constexpr size_t SMALL_SMEM = (size_t)SMALL * 6 * 8 + (size_t)SMALL * 8 + (size_t)SMALL * 4 * 2 + (size_t)2 * SMALL * 8 * 2 +
                              (size_t)SMALL * 2 + (size_t)2 * SMALL;

} // namespace vrj_build

// =====================================================================================================

namespace vrj_build {

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) vrj_pool_free(p);
    }
    cudaError_t alloc(size_t bytes) { return vrj_pool_alloc(&p, bytes); }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};

#define VRJB(expr)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            vrj_set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                       \
            return e_ == cudaErrorMemoryAllocation ? VRJ_ERR_OUT_OF_MEMORY : VRJ_ERR_CUDA;           \
        }                                                                                            \
    } while (0)

struct Slice { // a piece of one device block
    size_t offset = 0;
    void *p = nullptr;
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};

uint32_t tree_depth(uint64_t n) { // BuildCtx::build's return value: a leaf at level d reports d + 1
    uint32_t d = 1;
    while (n > 1) n = n - n / 2, d++;
    return d;
}

VrjStatus exclusive_scan_u32(uint32_t *d_data, uint32_t n, uint32_t *d_scratch, cudaStream_t stream) {
    if (n == 0) return VRJ_OK;
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<n_tiles, 256, 0, stream>>>(n, d_data, d_scratch);
    k_scan_totals<<<1, 1024, 0, stream>>>(n_tiles, d_scratch);
    k_scan_add<<<(n + 255) / 256, 256, 0, stream>>>(n, d_data, d_scratch);
    VRJB(cudaGetLastError());
    return VRJ_OK;
}

// Build on the current device.  d_vertices: 9 doubles per triangle (input order).  Outputs (device):
// d_order[n], d_node_min / d_node_max [4 * (2n-1)], d_node_child [2 * (2n-1)].  n >= 1.
VrjStatus build_device(uint32_t n, const double *d_vertices, uint32_t *d_order, double *d_node_min, double *d_node_max,
                       int32_t *d_node_child, cudaStream_t stream, VrjBvhBuildStats *stats) {
    // every temporary lives in ONE device block (a dozen cudaMalloc / cudaFree pairs cost more than a small build)
    Slice lo, hi, centre, keys[2], vals[2], d_segs, d_chunks, bounds, seg_axis, hist, totals, varying;
    size_t block_bytes = 0;
    auto carve = [&block_bytes](Slice &sl, size_t bytes) {
        sl.offset = block_bytes;
        block_bytes += (bytes + 255) & ~size_t(255);
    };
    carve(lo, (size_t)n * 24), carve(hi, (size_t)n * 24), carve(centre, (size_t)n * 24);

    // ---- the shape of the tree: segments of every global level, then the small-subtree roots ----
    std::vector<std::vector<Segment>> levels;
    std::vector<Segment> level{Segment{0, n, 0}};
    while (true) {
        uint32_t longest = 0;
        for (const Segment &s : level) longest = std::max(longest, s.len);
        if (longest <= (uint32_t)SMALL) break;
        levels.push_back(level);
        std::vector<Segment> next;
        next.reserve(level.size() * 2);
        for (const Segment &s : level) {
            const uint32_t half = s.len / 2;
            next.push_back(Segment{s.begin, half, s.me + 1});
            next.push_back(Segment{s.begin + half, s.len - half, s.me + 2 * half});
        }
        level.swap(next);
    }
    const std::vector<Segment> &roots = level; // every entry has 1 <= len <= SMALL (n >= 1 and sizes differ by at most one)
    // chunk tables of the global levels, all in one upload
    std::vector<Segment> all_segs;
    std::vector<Chunk> all_chunks;
    struct LevelInfo {
        size_t seg_off, n_seg, chunk_off, n_chunk;
    };
    std::vector<LevelInfo> info;
    size_t max_chunks = 0, max_segs = 1;
    for (const auto &lv : levels) {
        LevelInfo li{all_segs.size(), lv.size(), all_chunks.size(), 0};
        uint32_t first_chunk = 0;
        for (uint32_t s = 0; s < lv.size(); s++) {
            const uint32_t nb = (lv[s].len + CHUNK - 1) / CHUNK;
            for (uint32_t b = 0; b < nb; b++)
                all_chunks.push_back(Chunk{lv[s].begin + b * CHUNK, std::min<uint32_t>(CHUNK, lv[s].len - b * CHUNK), s, 256u * first_chunk, nb, b});
            first_chunk += nb;
        }
        li.n_chunk = all_chunks.size() - li.chunk_off;
        all_segs.insert(all_segs.end(), lv.begin(), lv.end());
        info.push_back(li);
        max_chunks = std::max(max_chunks, li.n_chunk), max_segs = std::max(max_segs, li.n_seg);
    }
    const size_t roots_off = all_segs.size();
    all_segs.insert(all_segs.end(), roots.begin(), roots.end());
    carve(d_segs, all_segs.size() * sizeof(Segment));
    if (!all_chunks.empty()) {
        carve(d_chunks, all_chunks.size() * sizeof(Chunk));
        for (int i = 0; i < 2; i++) carve(keys[i], (size_t)n * 8), carve(vals[i], (size_t)n * 4);
        carve(bounds, max_segs * 6 * 8), carve(seg_axis, max_segs * 4);
        carve(hist, max_chunks * 256 * 4), carve(totals, ((max_chunks * 256 + SCAN_TILE - 1) / SCAN_TILE + 1) * 4);
        carve(varying, 32);
    }
    DevBuf block;
    VRJB(block.alloc(block_bytes));
    for (Slice *sl : {&lo, &hi, &centre, &keys[0], &keys[1], &vals[0], &vals[1], &d_segs, &d_chunks, &bounds, &seg_axis, &hist, &totals, &varying})
        sl->p = block.as<char>() + sl->offset;
    k_tri_boxes<<<(n + 255) / 256, 256, 0, stream>>>(n, d_vertices, lo.as<double>(), hi.as<double>(), centre.as<double>(), d_order);
    VRJB(cudaMemcpyAsync(d_segs.p, all_segs.data(), all_segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, stream));
    if (!all_chunks.empty())
        VRJB(cudaMemcpyAsync(d_chunks.p, all_chunks.data(), all_chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice, stream));
    uint32_t passes_run = 0;
    for (size_t L = 0; L < info.size(); L++) {
        const LevelInfo &li = info[L];
        const Segment *segs = d_segs.as<Segment>() + li.seg_off;
        const Chunk *chunks = d_chunks.as<Chunk>() + li.chunk_off;
        const uint32_t n_seg = (uint32_t)li.n_seg, n_chunk = (uint32_t)li.n_chunk;
        k_init_bounds<<<(n_seg * 6 + 255) / 256, 256, 0, stream>>>(n_seg, bounds.as<unsigned long long>());
        k_seg_bounds<<<n_chunk, CHUNK_THREADS, 0, stream>>>(chunks, d_order, lo.as<double>(), hi.as<double>(), bounds.as<unsigned long long>());
        k_seg_nodes<<<(n_seg + 127) / 128, 128, 0, stream>>>(n_seg, segs, bounds.as<unsigned long long>(), d_node_min, d_node_max, d_node_child,
                                                             seg_axis.as<uint32_t>());
        const unsigned long long varying_init[4] = {0ull, ~0ull, 0ull, ~0ull};
        VRJB(cudaMemcpyAsync(varying.p, varying_init, 32, cudaMemcpyHostToDevice, stream));
        k_make_keys<<<n_chunk, CHUNK_THREADS, 0, stream>>>(chunks, seg_axis.as<uint32_t>(), d_order, centre.as<double>(),
                                                           keys[0].as<unsigned long long>(), vals[0].as<uint32_t>(), varying.as<unsigned long long>());
        unsigned long long var[4];
        VRJB(cudaMemcpyAsync(var, varying.p, 32, cudaMemcpyDeviceToHost, stream));
        VRJB(cudaStreamSynchronize(stream));
        const bool has_pos = var[0] != 0, has_neg = var[2] != 0 || var[3] != ~0ull; // keys >= 0 carry the top bit
        unsigned long long mask = (has_pos ? var[0] ^ var[1] : 0ull) | (has_neg ? var[2] ^ var[3] : 0ull);
        if (has_pos && has_neg) mask |= 0xffull << 56;
        int src = 0;
        const uint32_t n_hist = n_chunk * 256, n_tiles = (n_hist + SCAN_TILE - 1) / SCAN_TILE;
        for (int shift = 0; shift < 64; shift += 8) {
            if (((mask >> shift) & 255ull) == 0) continue; // no key of this level differs in this digit
            k_hist<<<n_chunk, CHUNK_THREADS, 0, stream>>>(chunks, keys[src].as<unsigned long long>(), shift, hist.as<uint32_t>());
            k_scan_tiles<<<n_tiles, 256, 0, stream>>>(n_hist, hist.as<uint32_t>(), totals.as<uint32_t>());
            k_scan_totals<<<1, 1024, 0, stream>>>(n_tiles, totals.as<uint32_t>());
            k_scatter<<<n_chunk, CHUNK_THREADS, 0, stream>>>(chunks, keys[src].as<unsigned long long>(), vals[src].as<uint32_t>(),
                                                             keys[src ^ 1].as<unsigned long long>(), vals[src ^ 1].as<uint32_t>(), shift,
                                                             hist.as<uint32_t>(), totals.as<uint32_t>());
            src ^= 1;
            passes_run++;
        }
        VRJB(cudaMemcpyAsync(d_order, vals[src].p, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    }
    VRJB(cudaFuncSetAttribute(k_small_subtrees, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM));
    k_small_subtrees<<<(unsigned)roots.size(), SMALL, SMALL_SMEM, stream>>>(d_segs.as<Segment>() + roots_off, d_order, lo.as<double>(), hi.as<double>(),
                                                                            centre.as<double>(), d_node_min, d_node_max, d_node_child);
    VRJB(cudaGetLastError());
    VRJB(cudaStreamSynchronize(stream));
    if (stats) {
        stats->global_levels = (uint32_t)info.size();
        stats->radix_passes = passes_run;
        stats->small_subtrees = (uint32_t)roots.size();
    }
    return VRJ_OK;
}

} // namespace vrj_build

extern "C" VRJ_API VrjStatus vrj_bvh_build(int32_t device, uint64_t n_triangles, const double *vertices, uint32_t *order,
                                           double *node_min, double *node_max, int32_t *node_child, uint32_t *depth,
                                           VrjBvhBuildStats *stats) {
    using namespace vrj_build;
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (n_triangles && (!vertices || !order)) {
        vrj_set_error("vrj_bvh_build: NULL argument");
        return VRJ_ERR_INVALID_ARGUMENT;
    }
    if (!node_min || !node_max || !node_child) {
        vrj_set_error("vrj_bvh_build: NULL argument");
        return VRJ_ERR_INVALID_ARGUMENT;
    }
    if (n_triangles > 0x3ffffff0ull) {
        vrj_set_error("vrj_bvh_build: too many triangles for 31-bit node indices");
        return VRJ_ERR_UNSUPPORTED;
    }
    if (depth) *depth = tree_depth(n_triangles);
    DeviceGuard device_guard;
    VRJB(cudaSetDevice(device)); // no device: VRJ_ERR_CUDA -- there is no CPU fallback behind this entry point
    if (n_triangles == 0) { // bounding_volume_hierarchy.rs:53-61: a single empty leaf with BoundingBox::empty()
        const double inf = std::numeric_limits<double>::infinity();
        for (int k = 0; k < 3; k++) node_min[k] = inf, node_max[k] = -inf;
        node_min[3] = node_max[3] = 0.0;
        node_child[0] = ~0, node_child[1] = 0;
        return VRJ_OK;
    }
    const uint32_t n = (uint32_t)n_triangles;
    const size_t n_nodes = 2 * (size_t)n - 1;
    cudaEvent_t e0, e1;
    VRJB(cudaEventCreate(&e0));
    VRJB(cudaEventCreate(&e1));
    cudaStream_t stream;
    VRJB(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct Cleanup {
        cudaEvent_t a, b;
        cudaStream_t s;
        ~Cleanup() { cudaEventDestroy(a), cudaEventDestroy(b), cudaStreamDestroy(s); }
    } cleanup{e0, e1, stream};
    DevBuf d_v, d_order, d_min, d_max, d_child;
    VRJB(d_v.alloc((size_t)n * 72));
    VRJB(d_order.alloc((size_t)n * 4));
    VRJB(d_min.alloc(n_nodes * 32));
    VRJB(d_max.alloc(n_nodes * 32));
    VRJB(d_child.alloc(n_nodes * 8));
    VRJB(cudaMemcpyAsync(d_v.p, vertices, (size_t)n * 72, cudaMemcpyHostToDevice, stream));
    VRJB(cudaEventRecord(e0, stream));
    VrjStatus st = build_device(n, d_v.as<double>(), d_order.as<uint32_t>(), d_min.as<double>(), d_max.as<double>(), d_child.as<int32_t>(),
                                stream, stats);
    if (st != VRJ_OK) return st;
    VRJB(cudaEventRecord(e1, stream));
    VRJB(cudaMemcpyAsync(order, d_order.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    VRJB(cudaMemcpyAsync(node_min, d_min.p, n_nodes * 32, cudaMemcpyDeviceToHost, stream));
    VRJB(cudaMemcpyAsync(node_max, d_max.p, n_nodes * 32, cudaMemcpyDeviceToHost, stream));
    VRJB(cudaMemcpyAsync(node_child, d_child.p, n_nodes * 8, cudaMemcpyDeviceToHost, stream));
    VRJB(cudaStreamSynchronize(stream));
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        stats->device_ms = ms;
    }
    return VRJ_OK;
}
