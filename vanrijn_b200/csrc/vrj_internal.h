// vrj_internal.h -- what the translation units of libvanrijn_cuda.so share (not part of the C ABI).
#pragma once
#include "../../include/vanrijn_cuda.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <string>

// the message vrj_last_error() returns on this thread (defined in vanrijn_cuda.cu)
void vrj_set_error(const std::string &msg);

// pooled device memory (vanrijn_cuda.cu): cudaMalloc / cudaFree are too slow to sit on the per-scene path
cudaError_t vrj_pool_alloc(void **out, size_t bytes);
void vrj_pool_free(void *p);
void vrj_pool_trim();

// every exported function that switches devices puts the caller's current device back (a torch caller would otherwise
// find itself on another GPU after vrj_render_sharded)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1, cudaGetLastError();
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

namespace vrj_build {
// BoundingVolumeHierarchy::build on the current device (vrj_bvh_build.cu).  d_vertices: 9 doubles per triangle in
// input order.  Outputs (device memory): d_order[n] (leaf position -> input index), d_node_min / d_node_max
// [4 * (2n-1)], d_node_child [2 * (2n-1)] with BVH-local indices.  n >= 1.  Synchronises the stream.
VrjStatus build_device(uint32_t n, const double *d_vertices, uint32_t *d_order, double *d_node_min, double *d_node_max,
                       int32_t *d_node_child, cudaStream_t stream, VrjBvhBuildStats *stats);
// levels of the median-split tree over n primitives (VrjBvh.depth)
uint32_t tree_depth(uint64_t n);
// in-place exclusive prefix sum of n uint32 values; `scratch` holds at least n / 2048 + 2 values
VrjStatus exclusive_scan_u32(uint32_t *d_data, uint32_t n, uint32_t *d_scratch, cudaStream_t stream);
} // namespace vrj_build
