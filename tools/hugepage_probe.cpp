// Dev probe (host only): non-temporal store bandwidth of N threads into 4 KB pages, into transparent huge pages
// (mmap + MADV_HUGEPAGE) and into CUDA-pinned memory is not covered here - does the page size of the destination
// matter for the expansion of packed volumes?   g++ -O2 -std=c++17 -pthread -mavx512f tools/hugepage_probe.cpp
#include <immintrin.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double run(char* buf, size_t bytes, int nt) {
    double best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back([=] {
            const size_t a = bytes / nt * t, b = bytes / nt * (t + 1);
            const __m512i v = _mm512_set1_epi32(0x3f800000);
            for (size_t i = a; i + 64 <= b; i += 64) _mm512_stream_si512((__m512i*)(buf + i), v);
            _mm_sfence();
        });
        for (auto& x : th) x.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    return bytes / best / 1e9;
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)3 << 30;
    const int nt = argc > 1 ? atoi(argv[1]) : 8;
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    char line[128] = "?";
    if (f) { if (!fgets(line, sizeof line, f)) line[0] = 0; fclose(f); }
    printf("transparent_hugepage: %s", line);
    char* small = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(small, bytes, MADV_NOHUGEPAGE);
    memset(small, 1, bytes);
    printf("4 KB pages      : %.1f GB/s with %d threads\n", run(small, bytes, nt), nt);
    munmap(small, bytes);
    char* huge = (char*)mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    char* al = (char*)(((uintptr_t)huge + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    madvise(al, bytes, MADV_HUGEPAGE);
    memset(al, 1, bytes);
    printf("huge pages (THP): %.1f GB/s with %d threads\n", run(al, bytes, nt), nt);
    return 0;
}
