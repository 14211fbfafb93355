// Dev probe: host memory write bandwidth with N threads (plain and non-temporal stores), pageable and pinned.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
#include <cuda_runtime.h>
static double run(float* buf, size_t n, int nt, int mode) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=] {
        size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        lo &= ~(size_t)15; hi &= ~(size_t)15;
        if (mode == 0) { for (size_t i = lo; i < hi; ++i) buf[i] = 1.0f; }
        else if (mode == 1) { __m128 v = _mm_set1_ps(1.0f); for (size_t i = lo; i < hi; i += 4) _mm_stream_ps(buf + i, v); _mm_sfence(); }
        else { memset(buf + lo, 0, (hi - lo) * 4); }
    });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int main() {
    const size_t n = (size_t)8500 * 3 * 32768;   // 3.34 GB
    float* pageable = (float*)aligned_alloc(64, n * 4);
    float* pinned = nullptr;
    cudaHostAlloc((void**)&pinned, n * 4, cudaHostAllocDefault);
    int hw = std::thread::hardware_concurrency();
    printf("hardware threads %d\n", hw);
    for (float* buf : {pageable, pinned}) {
        run(buf, n, hw, 2);
        for (int nt : {1, 4, 8, 16, 32}) {
            if (nt > hw) continue;
            for (int mode = 0; mode < 3; ++mode) {
                double best = 1e9;
                for (int r = 0; r < 3; ++r) best = std::min(best, run(buf, n, nt, mode));
                printf("%s threads %2d %-8s %.1f GB/s\n", buf == pinned ? "pinned  " : "pageable", nt, mode == 0 ? "plain" : mode == 1 ? "stream" : "memset", n * 4 / best / 1e9);
            }
        }
    }
    return 0;
}
