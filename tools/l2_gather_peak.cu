// l2_gather_peak.cu -- what bounds k_trace: the rate at which an SM array can fetch 64-byte records at data-dependent
// addresses from a table that lives in L2 (the bench scene's 5.2 MB of BVH nodes) or in HBM (config C4's 3.6 GB).
//
// Every lane follows its own chain: it loads a 64-byte record with two 256-bit read-only loads (ld.global.nc.v8, exactly
// what load_node issues) and the record holds the index of the next one -- like a traversal step, the next address is not
// known before the data arrives.  CHAINS independent chains per lane model the instruction-level parallelism a smarter
// walk could expose.  The grid is k_trace's (148 SMs x RES CTAs of 128 threads).
//
//   build/l2_gather_peak            prints one JSON object: GB/s and records/s per table size x chains x resident CTAs
//
// Used by bench.py as the denominator of roofline.k_trace (profiles/r02_l2_gather_peak.json holds the committed result).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}

template <int CHAINS, int REC> // REC: record size in bytes (32, 64 or 128), read with REC / 32 256-bit loads
__global__ void __launch_bounds__(128) k_chase(const uint4 *__restrict__ table, uint32_t n_records, int steps, uint32_t *sink) {
    uint32_t idx[CHAINS], acc = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) idx[c] = (uint32_t)(((uint64_t)(tid * CHAINS + c) * 2654435761u) % n_records);
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            const uint4 *p = table + (size_t)idx[c] * (REC / 16);
            uint32_t next = 0;
#pragma unroll
            for (int q = 0; q < REC / 32; q++) {
                uint4 a, b;
                ldg256(p + 2 * q, a, b);
                if (q == 0) next = a.x;                    // the next record: known only once the data is here
                acc += a.y ^ b.z ^ b.w;                    // every part of the record is consumed
            }
            idx[c] = next;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc; // keeps the loads alive
}

// The access pattern of a traversal rather than of a uniform gather: every lane walks from the root of a median-split binary
// tree (records in DFS pre-order, like the product's wide nodes: n - 1 records of 64 bytes for n leaves) to a leaf, choosing a
// child at random once the record has arrived, and starts again at the root.  The top levels are shared by all lanes and stay
// in L1, the bottom levels come from L2 (or HBM for config C4's tree) -- the mix k_trace sees, without any of its arithmetic.
__global__ void __launch_bounds__(128) k_tree_walk(const uint4 *__restrict__ table, int steps, uint32_t *sink) {
    uint32_t acc = 0, node = 0;
    uint32_t rng = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int s = 0; s < steps; s++) {
        uint4 a, b;
        ldg256(table + (size_t)node * 4, a, b);
        uint4 c, d;
        ldg256(table + (size_t)node * 4 + 2, c, d);
        acc += a.z ^ b.w ^ c.x ^ d.y;
        rng = rng * 1664525u + 1013904223u;
        const uint32_t child = (rng >> 16) & 1u ? a.y : a.x; // known only once the record is here
        node = child == 0xffffffffu ? 0u : child;             // a leaf: back to the root
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
// records of the tree over `n` leaves, DFS pre-order: word 0 / 1 = left / right child record, ~0 = that child is a leaf
static uint32_t build_tree(std::vector<uint32_t> &host, uint32_t &next, uint32_t n) {
    if (n <= 1) return 0xffffffffu;
    const uint32_t me = next++;
    const uint32_t l = build_tree(host, next, n / 2), r = build_tree(host, next, n - n / 2);
    host[(size_t)me * 16] = l, host[(size_t)me * 16 + 1] = r;
    return me;
}
static int run_tree(uint32_t leaves, int grid, int steps, uint32_t *d_sink, double *gbs, double *grecs) {
    std::vector<uint32_t> host((size_t)(leaves - 1) * 16, 0x9e3779b9u);
    uint32_t next = 0;
    build_tree(host, next, leaves);
    uint4 *d_table;
    CK(cudaMalloc(&d_table, host.size() * 4));
    CK(cudaMemcpy(d_table, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_tree_walk<<<grid, 128>>>(d_table, steps / 4, d_sink);
    CK(cudaEventRecord(e0));
    k_tree_walk<<<grid, 128>>>(d_table, steps, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *grecs = (double)grid * 128.0 * steps / (ms * 1e-3) / 1e9;
    *gbs = *grecs * 64;
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    CK(cudaFree(d_table));
    return 0;
}

template <int CHAINS, int REC>
static int run(const uint4 *d_table, uint32_t n_records, int grid, int steps, uint32_t *d_sink, double *gbs, double *grecs) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_chase<CHAINS, REC><<<grid, 128>>>(d_table, n_records, steps / 4, d_sink); // warm-up: pulls the table into L2
    CK(cudaEventRecord(e0));
    k_chase<CHAINS, REC><<<grid, 128>>>(d_table, n_records, steps, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double records = (double)grid * 128.0 * CHAINS * steps;
    *grecs = records / (ms * 1e-3) / 1e9;
    *gbs = *grecs * REC;
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    return 0;
}

// a random cyclic permutation over n records of `rec` bytes: the first word of record i names the record that follows it
static std::vector<uint32_t> make_table(uint32_t n, int rec, uint64_t seed) {
    std::vector<uint32_t> perm(n);
    for (uint32_t i = 0; i < n; i++) perm[i] = i;
    std::mt19937_64 rng(seed);
    for (uint32_t i = n - 1; i > 0; i--) std::swap(perm[i], perm[rng() % (i + 1)]);
    std::vector<uint32_t> host((size_t)n * (rec / 4), 0x9e3779b9u);
    for (uint32_t i = 0; i < n; i++) host[(size_t)perm[i] * (rec / 4)] = perm[(i + 1) % n];
    return host;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t sizes_mb[] = {5, 32, 96, 1024, 3648};
    uint32_t *d_sink;
    CK(cudaMalloc(&d_sink, 4));
    std::printf("{\"what\": \"dependent gathers of whole records (256-bit ld.global.nc.v8 loads), one chain per lane unless stated; "
                "record_bytes 64 = a BVH node as k_trace reads it\", \"sms\": %d, \"results\": [", sms);
    bool first = true;
    for (size_t mb : sizes_mb) {
        const uint32_t n = (uint32_t)(mb * 1024 * 1024 / 64);
        std::vector<uint32_t> host = make_table(n, 64, 12345 + mb);
        uint4 *d_table;
        CK(cudaMalloc(&d_table, (size_t)n * 64));
        CK(cudaMemcpy(d_table, host.data(), (size_t)n * 64, cudaMemcpyHostToDevice));
        for (int res : {6, 8, 16}) {
            const int grid = sms * res;
            const int steps = mb >= 1024 ? 256 : 1024;
            double gbs[3], gr[3];
            if (run<1, 64>(d_table, n, grid, steps, d_sink, &gbs[0], &gr[0])) return 1;
            if (run<2, 64>(d_table, n, grid, steps, d_sink, &gbs[1], &gr[1])) return 1;
            if (run<4, 64>(d_table, n, grid, steps / 2, d_sink, &gbs[2], &gr[2])) return 1;
            std::printf("%s\n {\"record_bytes\": 64, \"table_mb\": %zu, \"ctas_per_sm\": %d, \"GBps_1chain\": %.1f, \"GBps_2chains\": %.1f, \"GBps_4chains\": %.1f, "
                        "\"Grecords_per_s_1chain\": %.2f}",
                        first ? "" : ",", mb, res, gbs[0], gbs[1], gbs[2], gr[0]);
            first = false;
        }
        CK(cudaFree(d_table));
    }
    // is the L2-resident rate a byte rate or a request rate?  32-byte (the 16-bit nodes) and 128-byte (the 4-wide nodes) records
    for (int rec : {32, 128}) {
        const size_t mb = 32;
        const uint32_t n = (uint32_t)(mb * 1024 * 1024 / rec);
        std::vector<uint32_t> host = make_table(n, rec, 777 + rec);
        uint4 *d_table;
        CK(cudaMalloc(&d_table, (size_t)n * rec));
        CK(cudaMemcpy(d_table, host.data(), (size_t)n * rec, cudaMemcpyHostToDevice));
        double gbs = 0, gr = 0;
        if (rec == 32 ? run<1, 32>(d_table, n, sms * 6, 1024, d_sink, &gbs, &gr) : run<1, 128>(d_table, n, sms * 6, 1024, d_sink, &gbs, &gr)) return 1;
        std::printf(",\n {\"record_bytes\": %d, \"table_mb\": %zu, \"ctas_per_sm\": 6, \"GBps_1chain\": %.1f, \"Grecords_per_s_1chain\": %.2f}", rec, mb, gbs, gr);
        CK(cudaFree(d_table));
    }
    std::printf("\n],\n \"tree_walk\": {\"what\": \"root-to-leaf walks of a median-split tree in DFS pre-order, one 64-byte record per step, random child, "
                "one walk per lane, 148 x ctas_per_sm CTAs of 128 threads\", \"results\": [");
    first = true;
    for (uint32_t leaves : {81920u, 9912320u}) { // the bench mesh (5.2 MB of records) and config C4's (634 MB)
        for (int res : {6, 8}) {
            double gbs = 0, gr = 0;
            if (run_tree(leaves, sms * res, leaves > 1000000u ? 512 : 2048, d_sink, &gbs, &gr)) return 1;
            std::printf("%s\n  {\"leaves\": %u, \"table_mb\": %.1f, \"ctas_per_sm\": %d, \"GBps\": %.1f, \"Grecords_per_s\": %.2f}", first ? "" : ",", leaves,
                        (leaves - 1) * 64.0 / 1048576.0, res, gbs, gr);
            first = false;
        }
    }
    std::printf("\n ]}\n}\n");
    return 0;
}
