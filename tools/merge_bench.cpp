// merge_tile / merge_tiles on a 1080p frame: ms per merged buffer for 1, 2, 4, 8 buffers per pass (VRJ_MERGE_THREADS sets the
// pool size).  Build: make build/merge_bench
#include "vanrijn.hpp"
#include <chrono>
#include <cstdio>
#include <memory>
using namespace vanrijn;
int main() {
    const size_t W = 1920, H = 1080;
    AccumulationBuffer dst(W, H);
    std::vector<std::unique_ptr<AccumulationBuffer>> src;
    for (int k = 0; k < 8; k++) {
        src.emplace_back(new AccumulationBuffer(W, H));
        src[k]->weight.clear(), src[k]->uniform_weight = 1.0;
        for (size_t i = 0; i < 3 * W * H; i++) src[k]->colour[i] = 0.25 * (double)(i % 977) + k;
    }
    for (size_t i = 0; i < W * H; i++) dst.weight[i] = 3;
    Tile t{0, W, 0, H};
    for (size_t n : {1, 2, 4, 8}) {
        std::vector<const AccumulationBuffer *> p;
        for (size_t k = 0; k < n; k++) p.push_back(src[k].get());
        dst.merge_tiles(t, p);
        const int reps = 40;
        auto a = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; i++) dst.merge_tiles(t, p);
        auto b = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(b - a).count() / reps;
        std::printf("%zu buffers per pass: %.3f ms per pass, %.3f ms per buffer, %.0f GB/s\n", n, ms, ms / n, (133.0 + 50.0 * n) * 1e-3 / (ms * 1e-3));
    }
}
